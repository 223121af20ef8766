// ref_cu_probe.cpp — compiles the REFERENCE's intersection / bounds / hit-record / light programs
// (programs/hitables/*.cu, programs/pdfs/*.cu: included from /root/reference by -I, never copied)
// UNMODIFIED as host code against ref_shim/optix_device.h and exports plain-C probes.
// One reference file per translation unit (-DPROBE_<X>, see oracle/Makefile); the objects are linked
// into oracle/_ref/libref_cu.so.  TEST INFRASTRUCTURE ONLY: pins oracle.cpp's restatement of those
// files (tests/test_oracle_golden.py, tests/golden/make_golden_cu.py).
//
// Every probe has the same shape:
//   int refcu_<kind>(const float* prm, const float* ray8, float time, unsigned seed,
//                    float* hit3, float* rec14, float* bounds6)
// prm: the primitive's parameters (below); ray8: origin.xyz, tmin, dir.xyz, tmax; time/seed: prd.time and
// prd.seed for the programs that read the payload; returns 1 when the program reported a hit and then
//   hit3   = closest reported t, barycentric u, v
//   rec14  = HitRecord: P.xyz, geometric_normal.xyz, shading_normal.xyz, u, v, Wo.xyz
//   bounds6 = the bounding-box program's result (always written)
#include "ref_shim/optix_device.h"

static float3 ld(const float* p) { return make_float3(p[0], p[1], p[2]); }
static void st(float* p, const float3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

#if defined(PROBE_SPHERE)
#include "hitables/sphere.cu"
#define KIND sphere
static void set_params(const float* p) { center = ld(p); radius = p[3]; }
static void run_hit() { hit_sphere(0); }
static void run_bounds(float* b) { get_bounds(0, b); }
#define HAS_GET_HITRECORD 1
#elif defined(PROBE_RECT)
#include "hitables/aarect.cu"
#define KIND rect
static void set_params(const float* p) { a0 = p[0]; a1 = p[1]; b0 = p[2]; b1 = p[3]; k = p[4]; axis = (int)p[5]; flip = (int)p[6]; }
static void run_hit() { Hit_Rect(0); }
static void run_bounds(float* b) { Get_Bounds(0, b); }
#define HAS_GET_HITRECORD 1
#elif defined(PROBE_BOX)
#include "hitables/box.cu"
#define KIND box
static void set_params(const float* p) { boxmin = ld(p); boxmax = ld(p + 3); }
static void run_hit() { Intersect(0); }
static void run_bounds(float* b) { Get_Bounds(0, b); }
#define HAS_GET_HITRECORD 1
#elif defined(PROBE_TRIANGLE)
#include "hitables/triangle.cu"
#define KIND triangle
// prm: 9 vertex floats, flag_n, 9 normal floats, flag_uv, 6 uv floats
static void set_params(const float* p) {
  vertex_buffer.v = {ld(p), ld(p + 3), ld(p + 6)};
  index_buffer.v = {int3{0, 1, 2}};
  material_buffer.v = {0};
  normal_buffer.v.clear();
  texcoord_buffer.v.clear();
  if (p[9] != 0.f) normal_buffer.v = {ld(p + 10), ld(p + 13), ld(p + 16)};
  if (p[19] != 0.f) texcoord_buffer.v = {make_float2(p[20], p[21]), make_float2(p[22], p[23]), make_float2(p[24], p[25])};
}
static void run_hit() { Intersect(0); }
static void run_bounds(float* b) { Get_Bounds(0, b); }
#define HAS_GET_HITRECORD 1
#elif defined(PROBE_MOVING_SPHERE)
#include "hitables/moving_sphere.cu"
#define KIND moving_sphere
static void set_params(const float* p) { center0 = ld(p); center1 = ld(p + 3); radius = p[6]; time0 = p[7]; time1 = p[8]; index = 0; }
static void run_hit() { hit_sphere(0); }
static void run_bounds(float* b) { get_bounds(0, b); }
#define HAS_GET_HITRECORD 0
#elif defined(PROBE_VOLUME_SPHERE)
#include "hitables/volume_sphere.cu"
#define KIND volume_sphere
static void set_params(const float* p) { center = ld(p); radius = p[3]; density = p[4]; index = 0; }
static void run_hit() { hit_sphere(0); }
static void run_bounds(float* b) { get_bounds(0, b); }
#define HAS_GET_HITRECORD 0
#elif defined(PROBE_VOLUME_BOX)
#include "hitables/volume_box.cu"
#define KIND volume_box
static void set_params(const float* p) { boxmin = ld(p); boxmax = ld(p + 3); density = p[6]; index = 0; }
static void run_hit() { hit_volume(0); }
static void run_bounds(float* b) { get_bounds(0, b); }
#define HAS_GET_HITRECORD 0
#endif

#ifdef KIND
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

static void put_record(const HitRecord& rec, float* o) {
  st(o, rec.P);
  st(o + 3, rec.geometric_normal);
  st(o + 6, rec.shading_normal);
  o[9] = rec.u;
  o[10] = rec.v;
  st(o + 11, rec.Wo);
}

extern "C" int CAT(refcu_, KIND)(const float* prm, const float* ray8, float time, unsigned seed, float* hit3,
                                 float* rec14, float* bounds6) {
  set_params(prm);
  ray = make_Ray(ld(ray8), ld(ray8 + 4), 0, ray8[3], ray8[7]);
#if !HAS_GET_HITRECORD
  prd.time = time;
  prd.seed = seed;
#else
  (void)time;
  (void)seed;
  bc = make_float2(0.f);
#endif
  run_bounds(bounds6);
  shim_begin(ray8[3], ray8[7]);
  run_hit();
  if (g_hits.n == 0) return 0;
  hit3[0] = g_hits.tmax;  // the closest accepted hit
#if HAS_GET_HITRECORD
  hit3[1] = bc.x;
  hit3[2] = bc.y;
  // closest-hit programs call Get_HitRecord(geo_index, ray, t_hit, bc) (e.g. materials/lambertian.cu:51)
  put_record(Get_HitRecord(geo_index, ray, hit3[0], bc), rec14);
#else
  hit3[1] = hit3[2] = 0.f;
  put_record(hit_rec, rec14);  // old-style attribute written by the intersection program itself
#endif
  return 1;
}
#endif  // KIND

#if defined(PROBE_RECT_PDF)
#include "pdfs/rect_pdf.cu"
static void set_rect(const float* p) { a0 = p[0]; a1 = p[1]; b0 = p[2]; b1 = p[3]; k = p[4]; }
// prm: a0, a1, b0, b1, k ; axis 0/1/2
extern "C" float refcu_rect_pdf(const float* prm, int axis, const float* P, const float* Wi) {
  set_rect(prm);
  const float3 z = make_float3(0.f);
  return axis == 0 ? PDF_X(ld(P), z, ld(Wi), z) : axis == 1 ? PDF_Y(ld(P), z, ld(Wi), z) : PDF_Z(ld(P), z, ld(Wi), z);
}
extern "C" void refcu_rect_sample(const float* prm, int axis, const float* P, unsigned* seed, float* Wi) {
  set_rect(prm);
  const float3 z = make_float3(0.f);
  st(Wi, axis == 0 ? Sample_X(ld(P), z, z, *seed) : axis == 1 ? Sample_Y(ld(P), z, z, *seed) : Sample_Z(ld(P), z, z, *seed));
}
#endif

#if defined(PROBE_SPHERE_PDF)
#include "pdfs/sphere_pdf.cu"
// prm: center.xyz, radius
extern "C" float refcu_spherelight_pdf(const float* prm, const float* P, const float* Wi) {
  center = ld(prm);
  radius = prm[3];
  const float3 z = make_float3(0.f);
  return PDF(ld(P), z, ld(Wi), z);
}
extern "C" void refcu_spherelight_sample(const float* prm, const float* P, unsigned* seed, float* Wi) {
  center = ld(prm);
  radius = prm[3];
  const float3 z = make_float3(0.f);
  st(Wi, Sample(ld(P), z, z, *seed));
}
#endif

#if defined(PROBE_ANY_HIT)
#include "hit.cu"
// the per-hit rule of the shadow any-hit program (hit.cu:16-24): returns prd_shadow.inShadow after any_hit() ran on a
// hit of a material with `is_light` = light, for shading normal N and shadow-ray direction d.  WHICH hit OptiX hands
// to it is traversal-order dependent and stays the oracle's pin (the closest one).
extern "C" int refcu_any_hit(int light, const float* N, const float* d) {
  is_light = light;
  prd_shadow.normal = ld(N);
  prd_shadow.inShadow = false;
  ray = make_Ray(make_float3(0.f), ld(d), 1, 1e-3f, RT_DEFAULT_MAX);
  any_hit();
  return prd_shadow.inShadow ? 1 : 0;
}
#endif
