// ref_probe.cpp — compiles the REFERENCE's own header-only device code
// (included from /root/reference by -I, never copied) against ref_shim/ and
// exports plain-C probes.  Output goes to oracle/_ref/libref_probe.so.
// TEST INFRASTRUCTURE ONLY: used to pin oracle.cpp's restatement and to
// generate tests/golden/*.json (tests/golden/make_golden.py).
//
// NOTE on draw order: g++ evaluates `f(rnd(seed), rnd(seed))` right-to-left
// while nvcc device code (the reference's real target) goes left-to-right
// (SURVEY.md §7 hard part 1).  ref_arg_order() reports what THIS build does so
// tests can map draws onto the oracle's explicit (u1,u2) entry points.
#include <optix.h>

#include "random.cuh"
#include "sampling.cuh"
#include "vec.hpp"
#include "materials/material.cuh"
#include "materials/microfacets.cuh"
#include "materials/lambertian.cuh"
#include "materials/oren_nayar.cuh"
#include "materials/torrance_sparrow.cuh"
#include "materials/ashikhmin_shirley.cuh"
#include "materials/diffuse_light.cuh"
#include "materials/isotropic.cuh"

static float3 ld(const float* p) { return make_float3(p[0], p[1], p[2]); }
static void st(float* p, const float3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

extern "C" {

int ref_arg_order() {  // 1 = left-to-right, 0 = right-to-left
  uint seed = 12345u, s2 = 12345u;
  float first = rnd(s2);
  float2 r = make_float2(rnd(seed), rnd(seed));
  return r.x == first ? 1 : 0;
}
unsigned ref_tea64(unsigned a, unsigned b) { return tea<64>(a, b); }
float ref_rnd(unsigned* seed) { return rnd(*seed); }
void ref_random_in_unit_disk(unsigned* seed, float* o) { st(o, random_in_unit_disk(*seed)); }
void ref_random_in_unit_sphere(unsigned* seed, float* o) { st(o, random_in_unit_sphere(*seed)); }
void ref_random_on_unit_sphere(unsigned* seed, float* o) { st(o, random_on_unit_sphere(*seed)); }

// mat ids follow include/ptb200.h PT_MAT_*
void ref_bsdf_sample(int mat, const float* color, const float* color2, float p0, float p1,
                     const float* P, const float* Wo, const float* N, unsigned* seed, float* out) {
  float3 wi = make_float3(0.f);
  switch (mat) {
    case 0: { Lambertian_Parameters s; s.color = ld(color); wi = Sample(s, ld(P), ld(Wo), ld(N), *seed); break; }
    case 4: { Oren_Nayar_Parameters s; s.color = ld(color); s.rA = p0; s.rB = p1; wi = Sample(s, ld(P), ld(Wo), ld(N), *seed); break; }
    case 5: { Torrance_Sparrow_Parameters s; s.color = ld(color); s.nu = p0; s.nv = p1; wi = Sample(s, ld(P), ld(Wo), ld(N), *seed); break; }
    case 6: { Ashikhmin_Shirley_Parameters s; s.diffuse_color = ld(color); s.specular_color = ld(color2); s.nu = p0; s.nv = p1; wi = Sample(s, ld(P), ld(Wo), ld(N), *seed); break; }
    case 3: { Diffuse_Light_Parameters s; s.color = ld(color); wi = Sample(s, ld(P), ld(Wo), ld(N), *seed); break; }
    case 8: { Isotropic_Parameters s; s.color = ld(color); wi = Sample(s, ld(P), ld(Wo), ld(N), *seed); break; }
  }
  st(out, wi);
}
void ref_bsdf_eval(int mat, const float* color, const float* color2, float p0, float p1,
                   const float* P, const float* Wo, const float* Wi, const float* N, float* f,
                   float* pdf) {
  float3 v = make_float3(0.f);
  float pd = 0.f;  // same "unset reads as 0" pin as the oracle
  switch (mat) {
    case 0: { Lambertian_Parameters s; s.color = ld(color); v = Evaluate(s, ld(P), ld(Wo), ld(Wi), ld(N), pd); break; }
    case 4: { Oren_Nayar_Parameters s; s.color = ld(color); s.rA = p0; s.rB = p1; v = Evaluate(s, ld(P), ld(Wo), ld(Wi), ld(N), pd); break; }
    case 5: { Torrance_Sparrow_Parameters s; s.color = ld(color); s.nu = p0; s.nv = p1; v = Evaluate(s, ld(P), ld(Wo), ld(Wi), ld(N), pd); break; }
    case 6: { Ashikhmin_Shirley_Parameters s; s.diffuse_color = ld(color); s.specular_color = ld(color2); s.nu = p0; s.nv = p1; v = Evaluate(s, ld(P), ld(Wo), ld(Wi), ld(N), pd); break; }
    case 3: { Diffuse_Light_Parameters s; s.color = ld(color); v = Evaluate(s, ld(P), ld(Wo), ld(Wi), ld(N), pd); break; }
    case 8: { Isotropic_Parameters s; s.color = ld(color); v = Evaluate(s, ld(P), ld(Wo), ld(Wi), ld(N), pd); break; }
  }
  st(f, v);
  *pdf = pd;
}
float ref_schlick(float cosine, float ref_idx) { return schlick(cosine, ref_idx); }
int ref_refract(const float* v, const float* n, float ni_over_nt, float* out) {
  float3 r = make_float3(0.f);
  bool ok = Refract(ld(v), ld(n), ni_over_nt, r);
  st(out, r);
  return ok ? 1 : 0;
}
void ref_ggx_sample(const float* origin, float u1, float u2, float nu, float nv, float* out) {
  st(out, GGX_Sample(ld(origin), make_float2(u1, u2), nu, nv));
}
float ref_ggx_d(const float* H, float nu, float nv) { return GGX_D(ld(H), nu, nv); }
float ref_ggx_g(const float* Wo, const float* Wi, float nu, float nv) { return GGX_G(ld(Wo), ld(Wi), nu, nv); }
float ref_max_component(const float* v) { return max_component(ld(v)); }
}
