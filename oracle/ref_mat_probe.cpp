// ref_mat_probe.cpp — compiles the REFERENCE's closest-hit programs (programs/materials/*.cu, with
// light_sample.cuh's Direct_Light) UNMODIFIED as host code against ref_shim/optix_device.h, one material per
// translation unit (-DPROBE_MAT_<X>, see oracle/Makefile); linked into oracle/_ref/libref_mat.so together with
// the light programs of ref_cu_probe.cpp.  TEST INFRASTRUCTURE ONLY: pins oracle.cpp's restatement of the
// closest-hit programs, of Direct_Light and of their RNG consumption (tests/test_oracle_mat_golden.py).
//
//   void refmat_<name>(const float* rec14, const float* ray8, float t_hit, const float* tex6,
//                      const float* mp3, const float* lights, int in_shadow, float* prd16, float* shadow8)
// rec14: the hit record Get_HitRecord returns (P, Ng, Ns, u, v, Wo); tex6: the values of the material's one or
// two texture programs; mp3: its float parameters + flag; lights: count, then per light 10 floats
// (type, axis, a0, a1, b0, b1, k, emission.xyz - pt_light); in_shadow: what the shadow ray answers;
// prd16 / shadow8: as orc_closest_hit_probe (oracle.h).
#include "ref_shim/optix_device.h"

static float3 ld(const float* p) { return make_float3(p[0], p[1], p[2]); }
static void st(float* p, const float3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
// a texture callable: the given constant colour, or - in the emulated pipeline - texture program tex[0] of the scene
static std::function<float3(float, float, float3, int)> tex_fn(const float* tex) {
  return [=](float u, float v, float3 p, int idx) {
    if (refshim_texture_hook) {
      float o[3];
      refshim_texture_hook((int)tex[0], u, v, &p.x, idx, o);
      return ld(o);
    }
    return ld(tex);
  };
}

// light programs live in their own translation units (ref_cu_probe.cpp)
extern "C" float refcu_rect_pdf(const float* prm, int axis, const float* P, const float* Wi);
extern "C" void refcu_rect_sample(const float* prm, int axis, const float* P, unsigned* seed, float* Wi);
extern "C" float refcu_spherelight_pdf(const float* prm, const float* P, const float* Wi);
extern "C" void refcu_spherelight_sample(const float* prm, const float* P, unsigned* seed, float* Wi);

#if defined(PROBE_MAT_LAMBERTIAN)
#include "materials/lambertian.cu"
#define NAME lambertian
#define USES_LIGHTS 1
static void set_params(const float* tex, const float*) { sample_texture.f = tex_fn(tex); }
#elif defined(PROBE_MAT_METAL)
#include "materials/metal.cu"
#define NAME metal
#define USES_LIGHTS 0
static void set_params(const float* tex, const float* mp) {
  sample_texture.f = tex_fn(tex);
  fuzz = mp[0];
}
#elif defined(PROBE_MAT_DIELECTRIC)
#include "materials/dielectric.cu"
#define NAME dielectric
#define USES_LIGHTS 0
static void set_params(const float* tex, const float* mp) {
  base_texture.f = tex_fn(tex);
  extinction_texture.f = tex_fn(tex + 3);
  ref_idx = mp[0];
  density = mp[1];
}
#elif defined(PROBE_MAT_DIFFUSE_LIGHT)
#include "materials/diffuse_light.cu"
#define NAME diffuse_light
#define USES_LIGHTS 1
static void set_params(const float* tex, const float*) { sample_texture.f = tex_fn(tex); }
#elif defined(PROBE_MAT_OREN_NAYAR)
#include "materials/oren_nayar.cu"
#define NAME oren_nayar
#define USES_LIGHTS 1
static void set_params(const float* tex, const float* mp) {
  sample_texture.f = tex_fn(tex);
  rA = mp[0];
  rB = mp[1];
}
#elif defined(PROBE_MAT_TORRANCE_SPARROW)
#include "materials/torrance_sparrow.cu"
#define NAME torrance_sparrow
#define USES_LIGHTS 1
static void set_params(const float* tex, const float* mp) {
  sample_texture.f = tex_fn(tex);
  nu = mp[0];
  nv = mp[1];
}
#elif defined(PROBE_MAT_ASHIKHMIN_SHIRLEY)
#include "materials/ashikhmin_shirley.cu"
#define NAME ashikhmin_shirley
#define USES_LIGHTS 1
static void set_params(const float* tex, const float* mp) {
  diffuse_color.f = tex_fn(tex);
  specular_color.f = tex_fn(tex + 3);
  nu = mp[0];
  nv = mp[1];
}
#elif defined(PROBE_MAT_NORMAL_SHADER)
#include "materials/normal_shader.cu"
#define NAME normal_shader
#define USES_LIGHTS 0
static void set_params(const float*, const float* mp) { useShadingNormal = (int)mp[2]; }
#elif defined(PROBE_MAT_ISOTROPIC)
#include "materials/isotropic.cu"
#define NAME isotropic
#define USES_LIGHTS 1
static void set_params(const float* tex, const float*) { sample_texture.f = tex_fn(tex); }
#endif

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

extern "C" void CAT(refmat_, NAME)(const float* rec14, const float* ray8, float t, const float* tex6, const float* mp3,
                                   const float* lights, int in_shadow, float* prd16, float* shadow8) {
  set_params(tex6, mp3);
  HitRecord rec;
  rec.index = refshim_hit_index;
  rec.P = ld(rec14);
  rec.geometric_normal = ld(rec14 + 3);
  rec.shading_normal = ld(rec14 + 6);
  rec.u = rec14[9];
  rec.v = rec14[10];
  rec.Wo = ld(rec14 + 11);
  rec.t = t;
  Get_HitRecord.f = [=](int, Ray, float, float2) { return rec; };
  ray = make_Ray(ld(ray8), ld(ray8 + 4), 0, ray8[3], ray8[7]);
  t_hit = t;
  geo_index = 0;
  bc = make_float2(0.f);
#if USES_LIGHTS
  const int n = (int)lights[0];
  numLights = n;
  Light_Emissions.v.resize(n);
  Light_Sample.v.resize(n);
  Light_PDF.v.resize(n);
  for (int i = 0; i < n; i++) {
    const float* L = lights + 1 + 10 * i;
    Light_Emissions.v[i] = ld(L + 7);
    if ((int)L[0] == 0) {  // rectangle: a0 a1 b0 b1 k, axis
      const int axis = (int)L[1];
      Light_Sample.v[i].f = [=](const float3& P, const float3&, const float3&, uint& seed) {
        float w[3];
        refcu_rect_sample(L + 2, axis, &P.x, &seed, w);
        return ld(w);
      };
      Light_PDF.v[i].f = [=](const float3& P, const float3&, const float3& Wi, const float3&) {
        return refcu_rect_pdf(L + 2, axis, &P.x, &Wi.x);
      };
    } else {  // sphere: centre = (a0, a1, b0), radius = k
      Light_Sample.v[i].f = [=](const float3& P, const float3&, const float3&, uint& seed) {
        const float prm[4] = {L[2], L[3], L[4], L[6]};
        float w[3];
        refcu_spherelight_sample(prm, &P.x, &seed, w);
        return ld(w);
      };
      Light_PDF.v[i].f = [=](const float3& P, const float3&, const float3& Wi, const float3&) {
        const float prm[4] = {L[2], L[3], L[4], L[6]};
        return refcu_spherelight_pdf(prm, &P.x, &Wi.x);
      };
    }
  }
#else
  (void)lights;
#endif
  __builtin_memcpy(&prd.seed, prd16, 4);  // (<cstring> would bring in index(), a name the reference uses)
  prd.time = prd16[1];
  prd.throughput = ld(prd16 + 2);
  prd.radiance = ld(prd16 + 5);
  prd.origin = prd.direction = make_float3(0.f);
  prd.scatterEvent = rayGotBounced;
  prd.isSpecular = false;
  g_shadow.traced = 0;
  g_shadow.answer_in_shadow = in_shadow;
  g_shadow.ray = make_Ray(make_float3(0.f), make_float3(0.f), 1, 0.f, 0.f);

  closest_hit();

  __builtin_memcpy(prd16, &prd.seed, 4);
  prd16[1] = prd.time;
  st(prd16 + 2, prd.throughput);
  st(prd16 + 5, prd.radiance);
  st(prd16 + 8, prd.origin);
  st(prd16 + 11, prd.direction);
  prd16[14] = (float)(int)prd.scatterEvent;
  prd16[15] = prd.isSpecular ? 1.f : 0.f;
  shadow8[0] = (float)g_shadow.traced;
  st(shadow8 + 1, g_shadow.ray.origin);
  st(shadow8 + 4, g_shadow.ray.direction);
  shadow8[7] = (float)in_shadow;
}
