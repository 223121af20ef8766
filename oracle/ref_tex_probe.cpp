// ref_tex_probe.cpp — the REFERENCE's texture programs (programs/textures/{constant,checkered,gradient,noise}_texture.cu,
// vector.cu) compiled UNMODIFIED as host code against ref_shim/optix_device.h, one per translation unit.
// TEST INFRASTRUCTURE ONLY (tests/test_oracle_tex_golden.py).
//   void reftex_<name>(const float* prm, int n, float u, float v, const float* p, int index, float* rgb)
#include "ref_shim/optix_device.h"

static float3 ld(const float* p) { return make_float3(p[0], p[1], p[2]); }
static void st(float* p, const float3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

#if defined(PROBE_TEX_CONSTANT)
#include "textures/constant_texture.cu"
#define NAME constant
static void set_params(const float* q, int) { color = ld(q); }  // prm: colour
#elif defined(PROBE_TEX_CHECKER)
#include "textures/checkered_texture.cu"
#define NAME checker
static void set_params(const float* q, int) {  // prm: odd colour, even colour (constant sub-textures)
  odd.f = [=](float, float, float3, int) { return ld(q); };
  even.f = [=](float, float, float3, int) { return ld(q + 3); };
}
#elif defined(PROBE_TEX_GRADIENT)
#include "textures/gradient_texture.cu"
#define NAME gradient
static void set_params(const float* q, int) { colorA = ld(q); colorB = ld(q + 3); colorC = ld(q + 6); }
#elif defined(PROBE_TEX_VECTOR)
#include "textures/vector.cu"
#define NAME vector
static void set_params(const float* q, int n) {  // prm: n colours (constant sub-textures)
  size = n;
  texture_vector.v.resize(n);
  for (int i = 0; i < n; i++) texture_vector.v[i].f = [=](float, float, float3, int) { return ld(q + 3 * i); };
}
#elif defined(PROBE_TEX_NOISE)
#include "textures/noise_texture.cu"
#define NAME noise
static void set_params(const float* q, int) {  // prm: scale, axis, 768 gradient floats, 3 x 256 permutation entries
  scale = q[0];
  ax = (int)q[1];
  ranvec.v.resize(256);
  perm_x.v.resize(256);
  perm_y.v.resize(256);
  perm_z.v.resize(256);
  for (int i = 0; i < 256; i++) {
    ranvec.v[i] = ld(q + 2 + 3 * i);
    perm_x.v[i] = (int)q[2 + 768 + i];
    perm_y.v[i] = (int)q[2 + 1024 + i];
    perm_z.v[i] = (int)q[2 + 1280 + i];
  }
}
#endif

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
extern "C" void CAT(reftex_, NAME)(const float* prm, int n, float u, float v, const float* p, int index, float* rgb) {
  set_params(prm, n);
  st(rgb, sample_texture(u, v, ld(p), index));
}
