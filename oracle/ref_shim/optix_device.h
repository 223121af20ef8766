// optix_device.h — host emulation of the OptiX 6 *device-program* environment, so the reference's
// intersection / bounds / hit-record / light-pdf programs (programs/hitables/*.cu, programs/pdfs/*.cu)
// compile UNMODIFIED with g++ and can be called one at a time.  TEST INFRASTRUCTURE ONLY (used by
// oracle/ref_cu_probe.cpp to pin oracle.cpp's restatement of those files); nothing here is copied from
// the reference or from the OptiX SDK.
//
// One reference .cu file is compiled per translation unit (they all declare `ray`, `center`,
// `Get_HitRecord`, ... with the same names), so every program variable and program becomes `static`.
//
// What is emulated, and how:
//   rtDeclareVariable(T, name, semantic, annotation)   -> static T name   (the probe assigns them)
//   rtBuffer<T>                                         -> std::vector-backed buffer
//   RT_PROGRAM / RT_CALLABLE_PROGRAM                    -> static functions
//   rtPotentialIntersection(t) / rtReportIntersection   -> open interval (tmin, tmax) of the current
//        ray; an accepted hit shrinks tmax to its t, like OptiX does for closest-hit rays.  (The
//        interval convention is the oracle's pin, not something this emulation can confirm.)
//   rtTransformPoint / rtTransformNormal                -> identity: the probes cover untransformed
//        primitives (the reference mixes object and world space for transformed ones)
#pragma once
#include <optix.h>

#include <cfloat>
#include <vector>

using namespace optix;

#define RT_PROGRAM static
#define RT_CALLABLE_PROGRAM static
#define rtDeclareVariable(type, name, semantic, annotation) static type name

template <class T, int N = 1>
struct shim_buffer {
  std::vector<T> v;
  int w = 0;  // row length of a 2-D buffer
  T& operator[](size_t i) { return v[i]; }
  const T& operator[](size_t i) const { return v[i]; }
  T& operator[](const uint2& i) { return v[(size_t)i.y * w + i.x]; }
  size_t size() const { return v.size(); }
};
#define rtBuffer static shim_buffer  // `rtBuffer<float3> Light_Emissions;` -> a buffer private to this translation unit

enum RTtransformkind { RT_WORLD_TO_OBJECT = 0, RT_OBJECT_TO_WORLD = 1 };
static inline float3 rtTransformPoint(RTtransformkind, const float3& p) { return p; }
static inline float3 rtTransformVector(RTtransformkind, const float3& v) { return v; }
static inline float3 rtTransformNormal(RTtransformkind, const float3& n) { return n; }

struct Aabb {
  float3 m_min, m_max;
  void include(const Aabb& o) {
    m_min = make_float3(fminf(m_min.x, o.m_min.x), fminf(m_min.y, o.m_min.y), fminf(m_min.z, o.m_min.z));
    m_max = make_float3(fmaxf(m_max.x, o.m_max.x), fmaxf(m_max.y, o.m_max.y), fmaxf(m_max.z, o.m_max.z));
  }
};

// intersection reporting of the program being probed
struct ShimHits {
  float tmin, tmax;  // current valid interval
  float pending;
  int n;             // accepted hits, in report order
  float t[8];
};
static ShimHits g_hits;
static inline void shim_begin(float tmin, float tmax) {
  g_hits.tmin = tmin;
  g_hits.tmax = tmax;
  g_hits.n = 0;
}
static inline bool rtPotentialIntersection(float t) {
  if (t > g_hits.tmin && t < g_hits.tmax) {
    g_hits.pending = t;
    return true;
  }
  return false;
}
static inline bool rtReportIntersection(unsigned) {
  g_hits.tmax = g_hits.pending;
  if (g_hits.n < 8) g_hits.t[g_hits.n++] = g_hits.pending;
  return true;
}
static inline int rtGetPrimitiveIndex() { return 0; }
static inline float2 rtGetTriangleBarycentrics() { return make_float2(0.f); }

// ---- closest-hit programs (programs/materials/*.cu) ------------------------------------------------
//   rtCallableProgramId<R(A...)> / rtCallableProgramX<R(A...)>  -> std::function the probe fills in
//   rtTrace(world, shadowRay, prdShadow)                         -> records the ray and answers with the
//        visibility the probe was GIVEN (the shadow test itself is the oracle's pinned rule; what is
//        compared is everything the program does before and after it)
#include <functional>
template <class R, class... A>
struct rtCallableProgramId<R(A...)> {
  std::function<R(A...)> f;
  R operator()(A... a) const { return f(a...); }
};
template <class R, class... A>
struct rtCallableProgramX<R(A...)> {
  std::function<R(A...)> f;
  R operator()(A... a) const { return f(a...); }
};
typedef int rtObject;
struct ShimShadow {
  int answer_in_shadow;  // set by the probe's caller
  int traced;
  Ray ray;
};
static ShimShadow g_shadow;
// Hooks of the emulated pipeline (oracle/ref_render.cpp); null in the unit probes.  Weak, so that every translation
// unit of one library shares them.
extern "C" {
typedef int (*refshim_shadow_hook_t)(const float* ray8, const float* normal3);
typedef void (*refshim_texture_hook_t)(int tex, float u, float v, const float* p3, int index, float* rgb3);
__attribute__((weak)) refshim_shadow_hook_t refshim_shadow_hook = nullptr;
__attribute__((weak)) refshim_texture_hook_t refshim_texture_hook = nullptr;
// HitRecord::index of the hit being shaded: 0 for analytic primitives (their intersection programs report
// geo_index = 0), material_buffer[primitive] for mesh triangles (hitables/triangle.cu:146).  The float record
// that travels between the probes has no slot for it, so the emulated pipeline hands it over here.
__attribute__((weak)) int refshim_hit_index = 0;
}
template <class PRD>
static inline void shim_trace(const Ray& r, PRD& prd) {
  g_shadow.traced++;
  g_shadow.ray = r;
  if (refshim_shadow_hook) {
    const float ray8[8] = {r.origin.x, r.origin.y, r.origin.z, r.tmin, r.direction.x, r.direction.y, r.direction.z, r.tmax};
    prd.inShadow = refshim_shadow_hook(ray8, &prd.normal.x) != 0;
    return;
  }
  prd.inShadow = g_shadow.answer_in_shadow != 0;
}
// a macro, so that the `world` argument never reaches the compiler: light_sample.cuh's Direct_Light template
// names it before any .cu declares it (accepted by nvcc, rejected by g++'s two-phase lookup)
#define rtTrace(world_, ray_, prd_) shim_trace(ray_, prd_)
static inline void rtTerminateRay() {}
