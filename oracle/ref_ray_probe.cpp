// ref_ray_probe.cpp — the REFERENCE's ray-generation program (programs/raygen.cu: camera ray, the color() path
// loop with its scatter-event state machine, Russian roulette and depth cut-off, de_nan, accumulation, display
// pixel) and miss programs (programs/miss.cu), compiled UNMODIFIED as host code against
// ref_shim/optix_device.h.  TEST INFRASTRUCTURE ONLY (tests/test_oracle_ray_golden.py).
//
// rtTrace(world, ray, prd) of a RADIANCE ray is handed to a hook the caller installs: it must do what one
// trace does in OptiX - find the closest hit and run that material's closest-hit program, or the miss program -
// on the payload.  The test installs the oracle's single-trace entry point there, so what is compared is the
// reference's own loop around it.
#include "ref_shim/optix_device.h"

struct uchar4 {
  unsigned char x, y, z, w;
};
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return {x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline float4 make_float4(float s) { return {s, s, s, s}; }
static inline void operator+=(float4& a, const float4& b) { a = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

static float3 ld(const float* p) { return make_float3(p[0], p[1], p[2]); }
static void st(float* p, const float3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

#if defined(PROBE_RAYGEN)
#include "prd.cuh"
// payload layout of the hook = orc_closest_hit_probe's prd16 (oracle.h)
typedef void (*radiance_hook_t)(const float* ray8, float* prd16);
static radiance_hook_t g_hook;
static inline void shim_trace(const Ray& r, PerRayData& p) {  // preferred over the shadow-ray template
  float ray8[8] = {r.origin.x, r.origin.y, r.origin.z, r.tmin, r.direction.x, r.direction.y, r.direction.z, r.tmax};
  float q[16];
  __builtin_memcpy(q, &p.seed, 4);
  q[1] = p.time;
  st(q + 2, p.throughput);
  st(q + 5, p.radiance);
  st(q + 8, p.origin);
  st(q + 11, p.direction);
  q[14] = (float)(int)p.scatterEvent;
  q[15] = p.isSpecular ? 1.f : 0.f;
  g_hook(ray8, q);
  __builtin_memcpy(&p.seed, q, 4);
  p.time = q[1];
  p.throughput = ld(q + 2);
  p.radiance = ld(q + 5);
  p.origin = ld(q + 8);
  p.direction = ld(q + 11);
  p.scatterEvent = (ScatterEvent)(int)q[14];
  p.isSpecular = q[15] != 0.f;
}
#include "raygen.cu"

// cam: origin, lower_left_corner, horizontal, vertical, u, v (3 floats each), lens_radius, time0, time1
// renders frames [first, first + n) of a W x H image with the reference's renderPixel; acc: W*H float4,
// disp: W*H uchar4 (both row 0 = top, raygen.cu:153)
extern "C" void refray_render_rows(const float* cam, int W, int H, int y0, int y1, int first, int n, int max_depth,
                                   int russian_, radiance_hook_t hook, float* acc, unsigned char* disp) {
  camera_origin = ld(cam);
  camera_lower_left_corner = ld(cam + 3);
  camera_horizontal = ld(cam + 6);
  camera_vertical = ld(cam + 9);
  camera_u = ld(cam + 12);
  camera_v = ld(cam + 15);
  camera_lens_radius = cam[18];
  time0 = cam[19];
  time1 = cam[20];
  launchDim = make_uint2((unsigned)W, (unsigned)H);
  maxDepth = max_depth;
  russian = russian_;
  samples = first + n;
  g_hook = hook;
  acc_buffer.w = display_buffer.w = W;
  // only the buffer rows that hold pixel rows [y0, y1) travel in and out: a band costs its own pixels, not W*H
  const size_t n_px = (size_t)W * H, r0 = (size_t)(H - y1) * W, r1 = (size_t)(H - y0) * W;
  if (acc_buffer.v.size() != n_px) {
    acc_buffer.v.assign(n_px, make_float4(0.f));
    display_buffer.v.assign(n_px, make_uchar4(0, 0, 0, 0));
  }
  for (size_t i = r0; i < r1; i++) {
    acc_buffer.v[i] = {acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]};
    display_buffer.v[i] = make_uchar4(0, 0, 0, 0);
  }
  for (int f = first; f < first + n; f++) {
    frame = f;
    for (int y = y0; y < y1; y++)  // pixelID.y counts from the bottom row (raygen.cu:153)
      for (int x = 0; x < W; x++) {
        pixelID = make_uint2((unsigned)x, (unsigned)y);
        renderPixel();
      }
  }
  for (size_t i = r0; i < r1; i++) {
    const float4 a = acc_buffer.v[i];
    acc[4 * i] = a.x; acc[4 * i + 1] = a.y; acc[4 * i + 2] = a.z; acc[4 * i + 3] = a.w;
    const uchar4 d = display_buffer.v[i];
    disp[4 * i] = d.x; disp[4 * i + 1] = d.y; disp[4 * i + 2] = d.z; disp[4 * i + 3] = d.w;
  }
}
extern "C" void refray_render(const float* cam, int W, int H, int first, int n, int max_depth, int russian_,
                              radiance_hook_t hook, float* acc, unsigned char* disp) {
  refray_render_rows(cam, W, H, 0, H, first, n, max_depth, russian_, hook, acc, disp);
}
#endif

#if defined(PROBE_MISS)
#include "miss.cu"
// type: 0 gradient, 1 constant, 2 image background, 3 HDR environment (spherical flag); colors: color1, color2;
// texel: the value the image sampler returns; returns throughput after the program and the (u, v) the sampler got
extern "C" void refmiss(int type, int spherical, const float* color1, const float* color2, const float* texel,
                        const float* dir, const float* thr_in, float* thr_out, float* uv_out, int* event_out) {
  static float g_uv[2];
  g_uv[0] = g_uv[1] = 0.f;
  sample_color1.f = [=](float, float, float3, int) { return ld(color1); };
  sample_color2.f = [=](float, float, float3, int) { return ld(color2); };
  if (type == 1) sample_texture.f = [=](float, float, float3, int) { return ld(color1); };
  else sample_texture.f = [=](float u, float v, float3, int) { g_uv[0] = u; g_uv[1] = v; return ld(texel); };
  isSpherical = spherical;
  ray = make_Ray(make_float3(0.f), ld(dir), 0, 1e-3f, RT_DEFAULT_MAX);
  prd.throughput = ld(thr_in);
  prd.scatterEvent = rayGotBounced;
  if (type == 0) gradient_color();
  else if (type == 1) constant_color();
  else if (type == 2) image_background();
  else environmental_mapping();
  st(thr_out, prd.throughput);
  uv_out[0] = g_uv[0];
  uv_out[1] = g_uv[1];
  *event_out = (int)prd.scatterEvent;
}
#endif
