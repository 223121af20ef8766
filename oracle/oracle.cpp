// oracle.cpp — CPU oracle: host C++/OpenMP restatement of the reference's
// device programs.  TEST INFRASTRUCTURE ONLY (see oracle.h).
// Pinned bit for bit against the reference's own source compiled as host code
// (oracle/ref_*probe.cpp, tests/test_oracle_*golden.py): headers, intersection /
// hit-record / light / closest-hit / ray-generation / miss / texture programs.
//
// Citations use P/ = /root/reference/OptiX-Path-Tracer/programs/ and
// H/ = /root/reference/OptiX-Path-Tracer/host_includes/.
//
// Build: g++ -O2 -fopenmp -ffp-contract=off (no -ffast-math): every float
// expression below is evaluated exactly as written, left to right, in
// binary32, so the CUDA kernels (which use non-contracted arithmetic for the
// intersection tests) can match hit ids bit-exactly.
//
// Pinned semantics where the reference depends on OptiX runtime behaviour
// (SURVEY.md §7 hard part 3, §8a):
//  * uniform draws inside one argument list are taken in DEVICE order
//    (left to right), e.g. cosine_sample_hemisphere(rnd, rnd, ..) uses the
//    first draw as u1 (P/materials/lambertian.cuh:13).
//  * a ray's valid interval is the open interval (tmin, tmax); equal-t hits
//    are resolved towards the lowest primitive id.
//  * hit records of transformed primitives are computed in object space and
//    then moved to world space (the intended rtTransform* semantics; the
//    reference's Get_HitRecord mixes spaces, P/hitables/box.cu:66-71).
//  * shadow rays: visible <=> the CLOSEST hit in (1e-3, inf) is a
//    Diffuse_Light and dot(N, Wi) > 0; a miss is "in shadow"
//    (P/hit.cu:16-24 leaves it undefined).
//  * `pdf` left unset by an early-out Evaluate reads as 0.
//  * image textures: software bilinear, repeat addressing, texel centres at
//    +0.5 (the reference uses the hardware sampler's 8-bit weights).
//  * volumes (P/hitables/volume_sphere.cu, volume_box.cu) draw rnd(prd.seed)
//    inside the intersection program, i.e. once per program invocation in
//    OptiX's traversal order and under OptiX's culling - unknowable.  Pinned:
//    the uniform of (ray, volume primitive) is rnd(tea<4>(ray seed, prim id)),
//    where the ray seed is prd.seed when the ray is traced; the path's stream
//    is not advanced by traversal.  Shadow rays (whose payload has no seed in
//    the reference) use the seed at the time Direct_Light traces them.
//  * moving spheres are bounded over the camera's shutter interval as well as
//    their own [time0,time1] (the reference bounds only the latter).
#include "oracle.h"

#include <omp.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// Two uniforms drawn inside ONE argument list (cosine_sample_hemisphere(rnd(seed), rnd(seed), ..),
// make_float3(k, a0 + rnd(seed) * .., b0 + rnd(seed) * ..)): device code evaluates left to right - the
// pinned order.  -DORC_ARGS_R2L builds the variant in which they are taken right to left, which is what g++
// does: oracle/_ref/libref_mat.so is the reference's closest-hit programs compiled by g++, and
// liboracle_r2l.so is what must reproduce THEM bit for bit (tests/test_oracle_mat_golden.py).  The two oracle
// builds differ in nothing else.
#ifdef ORC_ARGS_R2L
#define ORC_DRAW2(seed, a, b) \
  float b = rnd(seed);        \
  float a = rnd(seed)
#else
#define ORC_DRAW2(seed, a, b) \
  float a = rnd(seed);        \
  float b = rnd(seed)
#endif

// -DORC_ULP_NOISE builds liboracle_ulp.so: the same oracle with every libm call (sinf, cosf, tanf, atanf, atan2f,
// asinf, acosf, logf, powf, expf - the functions whose CUDA implementations are allowed to differ from glibc's by
// 1-2 ulp; sqrtf and the arithmetic are IEEE-exact on both sides) nudged by -1 / 0 / +1 ulp, chosen by a hash of
// the result.  It answers ONE question (tests/test_gpu_parity.py::test_fixed_seed_radiance): how many pixels of a
// fixed-seed image change when nothing but last-bit libm rounding changes?  That measured fraction is the bar the
// GPU image is held to - the GPU's differing pixels are expected to be exactly this kind (a flipped branch after
// a 1-ulp difference), so it must not show more of them than ulp noise alone produces.
#ifdef ORC_ULP_NOISE
namespace ulp_noise {
inline float nudge(float v) {
  if (!(v == v) || std::isinf(v) || v == 0.f) return v;
  uint32_t b;
  memcpy(&b, &v, 4);
  uint32_t hsh = b * 2654435761u;
  hsh ^= hsh >> 15;
  hsh *= 2246822519u;
  hsh ^= hsh >> 13;
  const uint32_t k = hsh >> 30;  // 0,1: unchanged; 2: +1 ulp in magnitude; 3: -1 ulp
  if (k == 2) b += 1;
  else if (k == 3) b -= 1;
  memcpy(&v, &b, 4);
  return v;
}
inline float n_sinf(float x) { return nudge(::sinf(x)); }
inline float n_cosf(float x) { return nudge(::cosf(x)); }
inline float n_tanf(float x) { return nudge(::tanf(x)); }
inline float n_atanf(float x) { return nudge(::atanf(x)); }
inline float n_atan2f(float y, float x) { return nudge(::atan2f(y, x)); }
inline float n_asinf(float x) { return nudge(::asinf(x)); }
inline float n_acosf(float x) { return nudge(::acosf(x)); }
inline float n_logf(float x) { return nudge(::logf(x)); }
inline float n_powf(float x, float y) { return nudge(::powf(x, y)); }
inline float n_expf(float x) { return nudge(::expf(x)); }
}  // namespace ulp_noise
#define sinf ulp_noise::n_sinf
#define cosf ulp_noise::n_cosf
#define tanf ulp_noise::n_tanf
#define atanf ulp_noise::n_atanf
#define atan2f ulp_noise::n_atan2f
#define asinf ulp_noise::n_asinf
#define acosf ulp_noise::n_acosf
#define logf ulp_noise::n_logf
#define powf ulp_noise::n_powf
#define expf ulp_noise::n_expf
#endif

namespace {

// ---------------------------------------------------------------------------
// float3 with the OptiX SDK's optixu_math_namespace.h semantics (SDK 6.0.0,
// not vendored in the reference; restated from its public definitions):
// normalize(v) = v * (1/sqrtf(dot(v,v))), v/s = v * (1/s).
// ---------------------------------------------------------------------------
struct f3 {
  float x, y, z;
};
inline f3 mk(float x, float y, float z) { return {x, y, z}; }
inline f3 mk(float s) { return {s, s, s}; }
inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline f3 operator/(f3 a, f3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline f3 operator/(f3 a, float s) {
  float inv = 1.0f / s;
  return a * inv;
}
inline f3 operator+(f3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline f3 operator-(f3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline f3& operator+=(f3& a, f3 b) { return a = a + b; }
inline f3& operator*=(f3& a, f3 b) { return a = a * b; }
inline f3& operator*=(f3& a, float s) { return a = a * s; }
inline f3& operator/=(f3& a, float s) { return a = a / s; }
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline f3 cross(f3 a, f3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(f3 v) { return sqrtf(dot(v, v)); }
inline f3 normalize(f3 v) {
  float invLen = 1.0f / sqrtf(dot(v, v));
  return v * invLen;
}
inline float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
inline f3 clamp3(f3 v, float a, float b) {
  return {clampf(v.x, a, b), clampf(v.y, a, b), clampf(v.z, a, b)};
}
inline f3 reflect(f3 i, f3 n) { return i - 2.0f * n * dot(n, i); }
inline f3 ld3(const float* p) { return {p[0], p[1], p[2]}; }
inline void st3(float* p, f3 v) {
  p[0] = v.x;
  p[1] = v.y;
  p[2] = v.z;
}

const float PI_F = 3.141592654f;                // P/common.hpp:18
const float OPTIX_PIf = 3.14159265358979323846f;    // optixu_math: OPTIX_PIf
const float OPTIX_1_PIf = 0.318309886183790671538f; // optixu_math: OPTIX_1_PIf
const float RT_DEFAULT_MAX = 1.e27f;            // optix: RT_DEFAULT_MAX

// P/vec.hpp:24-84
inline float squared_length(f3 a) { return length(a) * length(a); }
inline bool isNull(f3 a) { return (a.x == 0.f) && (a.y == 0.f) && (a.z == 0.f); }
inline f3 unit_vector(f3 v) { return v / length(v); }
inline float ffmin(float a, float b) { return a < b ? a : b; }
inline float ffmax(float a, float b) { return a > b ? a : b; }
inline f3 min_vec(f3 a, f3 b) { return {ffmin(a.x, b.x), ffmin(a.y, b.y), ffmin(a.z, b.z)}; }
inline f3 max_vec(f3 a, f3 b) { return {ffmax(a.x, b.x), ffmax(a.y, b.y), ffmax(a.z, b.z)}; }
inline float max_component(f3 a) { return ffmax(ffmax(a.x, a.y), a.z); }
inline float min_component(f3 a) { return ffmin(ffmin(a.x, a.y), a.z); }
inline float saturate(float x) { return clampf(x, 0.f, 1.f); }

// optixu_math: cosine_sample_hemisphere, Onb (SDK 6.0.0)
inline void cosine_sample_hemisphere(float u1, float u2, f3& p) {
  const float r = sqrtf(u1);
  const float phi = 2.0f * OPTIX_PIf * u2;
  p.x = r * cosf(phi);
  p.y = r * sinf(phi);
  p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
}
struct Onb {
  f3 m_tangent, m_binormal, m_normal;
  explicit Onb(f3 normal) {
    m_normal = normal;
    if (fabsf(m_normal.x) > fabsf(m_normal.z)) {
      m_binormal.x = -m_normal.y;
      m_binormal.y = m_normal.x;
      m_binormal.z = 0;
    } else {
      m_binormal.x = 0;
      m_binormal.y = -m_normal.z;
      m_binormal.z = m_normal.y;
    }
    m_binormal = normalize(m_binormal);
    m_tangent = cross(m_binormal, m_normal);
  }
  void inverse_transform(f3& p) const {
    p = p.x * m_tangent + p.y * m_binormal + p.z * m_normal;
  }
};

// ---------------------------------------------------------------------------
// RNG — P/random.cuh:33-65 (integer arithmetic, must be bit-identical)
// ---------------------------------------------------------------------------
template <unsigned N>
inline uint32_t tea(uint32_t val0, uint32_t val1) {
  uint32_t v0 = val0, v1 = val1, s0 = 0;
  for (unsigned n = 0; n < N; n++) {
    s0 += 0x9e3779b9;
    v0 += ((v1 << 4) + 0xa341316c) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4);
    v1 += ((v0 << 4) + 0xad90777d) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761e);
  }
  return v0;
}
inline uint32_t lcg(uint32_t& prev) {
  prev = 1664525u * prev + 1013904223u;
  return prev & 0x00FFFFFF;
}
inline float rnd(uint32_t& prev) { return (float)lcg(prev) / (float)0x01000000; }

// ---------------------------------------------------------------------------
// sampling — P/sampling.cuh:22-59
// ---------------------------------------------------------------------------
inline f3 random_in_unit_disk(uint32_t& seed) {
  float a = rnd(seed) * 2.f * PI_F;
  f3 xy = mk(sinf(a), cosf(a), 0);
  xy *= sqrtf(rnd(seed));
  return xy;
}
inline f3 random_in_unit_sphere(uint32_t& seed) {
  float z = rnd(seed) * 2.f - 1.f;
  float t = rnd(seed) * 2.f * PI_F;
  float r = sqrtf((0.f > (1.f - z * z) ? 0.f : (1.f - z * z)));
  float x = r * cosf(t);
  float y = r * sinf(t);
  f3 res = mk(x, y, z);
  res *= powf(rnd(seed), 1.f / 3.f);
  return res;
}
inline f3 random_on_unit_sphere(uint32_t& seed) { return unit_vector(random_in_unit_sphere(seed)); }

// ---------------------------------------------------------------------------
// shading-frame trigonometry (world Y up) — P/math/trigonometric.cuh:5-43,124
// ---------------------------------------------------------------------------
inline float CosTheta(f3 w) { return w.y; }
inline float Cos2Theta(f3 w) { return w.y * w.y; }
inline float AbsCosTheta(f3 w) { return fabsf(w.y); }
inline float Sin2Theta(f3 w) { return fmaxf(0.f, 1.f - Cos2Theta(w)); }
inline float SinTheta(f3 w) { return sqrtf(Sin2Theta(w)); }
inline float TanTheta(f3 w) { return SinTheta(w) / CosTheta(w); }
inline float CosPhi(f3 w) {
  float sinTheta = SinTheta(w);
  return (sinTheta == 0) ? 1.f : clampf(w.x / sinTheta, -1.f, 1.f);
}
inline float SinPhi(f3 w) {
  float sinTheta = SinTheta(w);
  return (sinTheta == 0) ? 0.f : clampf(w.z / sinTheta, -1.f, 1.f);
}
inline float Cos2Phi(f3 w) {
  float c = CosPhi(w);
  return c * c;
}
inline float Sin2Phi(f3 w) {
  float s = SinPhi(w);
  return s * s;
}
inline f3 Spherical_Vector(float sintheta, float costheta, float phi) {
  return mk(sintheta * cosf(phi), costheta, sintheta * sinf(phi));
}

// ---------------------------------------------------------------------------
// P/materials/material.cuh:30-63
// ---------------------------------------------------------------------------
inline float schlick(float cosine, float ref_idx) {
  float r0 = (1.f - ref_idx) / (1.f + ref_idx);
  r0 = r0 * r0;
  return r0 + (1.f - r0) * powf((1.f - cosine), 5.f);
}
inline f3 schlick3(f3 r0, float cosine) {
  float exponential = powf(1.f - cosine, 5.f);
  return r0 + (mk(1.f) - r0) * exponential;
}
inline bool Refract(f3 v, f3 n, float ni_over_nt, f3& refracted) {
  f3 uv = normalize(v);
  float dt = dot(uv, n);
  float discriminant = 1.f - ni_over_nt * ni_over_nt * (1.f - dt * dt);
  if (discriminant > 0.f) {
    refracted = ni_over_nt * (uv - n * dt) - n * sqrtf(discriminant);
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------
// P/materials/microfacets.cuh:11-20, 85-201
// ---------------------------------------------------------------------------
inline void Sample_Quadrant(float nu, float nv, float randX, float randY, float& phi,
                            float& theta) {
  phi = atanf(sqrtf((nu + 1.f) / (nv + 1.f)) * tanf(2.f * PI_F * randX));
  float cos_phi = cosf(phi);
  float sin_phi = sinf(phi);
  theta = powf(randY, 1.f / (nu * cos_phi * cos_phi + nv * sin_phi * sin_phi + 1.f));
}
inline float GGX_D(f3 H, float nu, float nv) {
  const float CosTheta2 = Cos2Theta(H);
  if (CosTheta2 <= 0.0f) return 0.f;
  if (nu == nv) {
    float beta = CosTheta2 * ((nu * nu) - 1) + 1;
    return (nu * nu) / (PI_F * beta * beta);
  } else {
    const float Sin2PhiNV = Sin2Phi(H) / (nv * nv);
    const float Cos2PhiNU = (1.f - Sin2Phi(H)) / (nu * nu);
    float beta = (CosTheta2 + (1.f - CosTheta2) * (Cos2PhiNU + Sin2PhiNV));
    return 1.f / (PI_F * nu * nv * beta * beta);
  }
}
inline f3 GGX_Sample(f3 origin, float rx, float ry, float nu, float nv) {
  bool flip = origin.y < 0;
  f3 so = mk(origin.x * nu, origin.y, origin.z * nv);
  so = normalize(so);
  float slopeX, slopeY;
  float cosTheta = CosTheta(so);
  if (cosTheta > 0.9999f) {
    float r = sqrtf(rx / (1 - rx));
    float phi = (float)(6.28318530718 * (double)ry);  // double literal, :123
    slopeX = r * cosf(phi);
    slopeY = r * sinf(phi);
  } else {
    float sinTheta = sqrtf(fmaxf(0.f, 1.f - cosTheta * cosTheta));
    float tanTheta = sinTheta / cosTheta;
    float a = 1.f / tanTheta;
    float G1 = 2.f / (1.f + sqrtf(1.f + 1.f / (a * a)));
    float A = 2.f * rx / G1 - 1.f;
    float tmp = 1.f / (A * A - 1.f);
    if (tmp > 1e10f) tmp = 1e10f;
    float B = tanTheta;
    float D = sqrtf(fmaxf(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
    float slopeX1 = B * tmp - D;
    float slopeX2 = B * tmp + D;
    slopeX = (A < 0 || slopeX2 > 1.f / tanTheta) ? slopeX1 : slopeX2;
    float S;
    if (ry > 0.5f) {
      S = 1.f;
      ry = 2.f * (ry - .5f);
    } else {
      S = -1.f;
      ry = 2.f * (.5f - ry);
    }
    float z = (ry * (ry * (ry * 0.27385f - 0.73369f) + 0.46341f)) /
              (ry * (ry * (ry * 0.093073f + 0.309420f) - 1.f) + 0.597999f);
    slopeY = S * z * sqrtf(1.f + slopeX * slopeX);
  }
  float t = CosPhi(so) * slopeX - SinPhi(so) * slopeY;
  slopeY = SinPhi(so) * slopeX + CosPhi(so) * slopeY;
  slopeX = t;
  slopeX = nu * slopeX;
  slopeY = nv * slopeY;
  f3 H = normalize(mk(-slopeX, 1.f, -slopeY));
  if (flip) H *= -1.f;
  return H;
}
inline float GGX_Lambda(f3 V, float nu, float nv) {
  float absTanTheta = fabsf(TanTheta(V));
  if (std::isinf(absTanTheta)) return 0.f;
  float alpha = sqrtf(Cos2Phi(V) * nu * nu + Sin2Phi(V) * nv * nv);
  float alpha2Tan2Theta = (alpha * absTanTheta) * (alpha * absTanTheta);
  return (-1.f + sqrtf(1.f + alpha2Tan2Theta)) / 2.f;
}
inline float GGX_G(f3 Wo, f3 Wi, float nu, float nv) {
  return 1.f / (1.f + GGX_Lambda(Wo, nu, nv) + GGX_Lambda(Wi, nu, nv));
}
inline float GGX_PDF(f3 H, f3 /*origin*/, float nu, float nv) {
  return GGX_D(H, nu, nv) * AbsCosTheta(H);
}

// ---------------------------------------------------------------------------
// BSDF Sample / Evaluate — one struct for every material's *_Parameters
// ---------------------------------------------------------------------------
struct Surface {
  int type;
  f3 color;   // sample_texture / diffuse_color / base
  f3 color2;  // specular_color (Ashikhmin-Shirley)
  float p0, p1;
};

// P/materials/lambertian.cuh:7-19, oren_nayar.cuh:8-20 (identical bodies)
inline f3 Sample_cosine(f3 N, float u1, float u2) {
  f3 Wi;
  cosine_sample_hemisphere(u1, u2, Wi);
  Onb uvw(N);
  uvw.inverse_transform(Wi);
  return Wi;
}
// P/materials/torrance_sparrow.cuh:10-37
inline f3 Sample_TS(const Surface& s, f3 Wo, f3 N, float u1, float u2) {
  float nu = s.p0, nv = s.p1;
  f3 Nn = normalize(N);
  f3 T = normalize(cross(Nn, mk(0.f, 1.f, 0.f)));
  f3 B = cross(T, Nn);
  f3 H = normalize(GGX_Sample(Wo, u1, u2, nu, nv));
  H = H.x * B + H.y * Nn + H.z * T;
  float HdotI = dot(H, Wo);
  if (HdotI < 0.f) H = -H;
  return normalize(-Wo + 2.f * dot(Wo, H) * H);
}
// P/materials/ashikhmin_shirley.cuh:11-85
inline f3 Sample_AS(const Surface& s, f3 Wo, f3 Ns, float rx, float ry) {
  float nu = s.p0, nv = s.p1;
  f3 Wi;
  f3 N = normalize(Ns);
  f3 T = normalize(cross(N, mk(0.f, 1.f, 0.f)));
  f3 B = cross(T, N);
  if ((double)rx < 0.5) {
    rx = fminf(2 * rx, 1.f - 1e-6f);
    cosine_sample_hemisphere(rx, ry, Wi);
    Onb uvw(N);
    uvw.inverse_transform(Wi);
  } else {
    rx = fminf(2 * (rx - 0.5f), 1.f - 1e-6f);
    float phi, cos_theta;
    if (nu == nv) {
      phi = 2.f * PI_F * rx;
      cos_theta = powf(ry, 1.f / (nu + 1.f));
    } else {
      if (rx < 0.25f) {
        float r = 4.f * rx;
        Sample_Quadrant(nu, nv, r, ry, phi, cos_theta);
      } else if (rx < 0.5f) {
        float r = 4.f * (0.5f - rx);
        Sample_Quadrant(nu, nv, r, ry, phi, cos_theta);
        phi = PI_F - phi;
      } else if (rx < 0.75f) {
        float r = 4.f * (rx - 0.5f);
        Sample_Quadrant(nu, nv, r, ry, phi, cos_theta);
        phi = PI_F + phi;
      } else {
        float r = 4.f * (1.f - rx);
        Sample_Quadrant(nu, nv, r, ry, phi, cos_theta);
        phi = 2.f * PI_F - phi;
      }
    }
    float sin_theta = sqrtf(fmaxf(0.f, 1.f - cos_theta * cos_theta));
    f3 H = normalize(Spherical_Vector(sin_theta, cos_theta, phi));
    H = H.x * B + H.y * N + H.z * T;
    float HdotI = dot(H, Wo);
    if (HdotI < 0.f) H = -H;
    Wi = normalize(-Wo + 2.f * dot(Wo, H) * H);
  }
  return Wi;
}

// draws two uniforms in device order, then samples
inline f3 Sample(const Surface& s, f3 /*P*/, f3 Wo, f3 N, uint32_t& seed) {
  switch (s.type) {
    case PT_MAT_LAMBERTIAN:
    case PT_MAT_OREN_NAYAR: {
      ORC_DRAW2(seed, u1, u2);
      return Sample_cosine(N, u1, u2);
    }
    case PT_MAT_TORRANCE_SPARROW: {
      ORC_DRAW2(seed, u1, u2);
      return Sample_TS(s, Wo, N, u1, u2);
    }
    case PT_MAT_ASHIKHMIN_SHIRLEY: {
      ORC_DRAW2(seed, u1, u2);
      return Sample_AS(s, Wo, N, u1, u2);
    }
    case PT_MAT_DIFFUSE_LIGHT:  // P/materials/diffuse_light.cuh:7-13
    case PT_MAT_ISOTROPIC:      // P/materials/isotropic.cuh:7-13
      return random_on_unit_sphere(seed);
    default:
      return mk(0.f);
  }
}

// P/materials/lambertian.cuh:29-43 ; pdf unset on early-out reads as 0 (pinned)
inline f3 Evaluate_Lambertian(const Surface& s, f3 Wi, f3 N, float& pdf) {
  float cosine = dot(normalize(Wi), normalize(N));
  if (cosine < 0.f) return mk(0.f);
  pdf = dot(normalize(Wi), normalize(N)) / PI_F;
  return (cosine * s.color) / PI_F;
}
// P/materials/oren_nayar.cuh:30-65
inline f3 Evaluate_ON(const Surface& s, f3 Wo, f3 Wi, f3 N, float& pdf) {
  f3 WiN = normalize(Wi);
  float sinThetaI = SinTheta(WiN);
  float sinThetaO = SinTheta(Wo);
  float maxCos = 0;
  if ((double)sinThetaI > 1e-4 && (double)sinThetaO > 1e-4) {
    float sinPhiI = SinPhi(WiN), cosPhiI = CosPhi(WiN);
    float sinPhiO = SinPhi(Wo), cosPhiO = CosPhi(Wo);
    float dCos = cosPhiI * cosPhiO + sinPhiI * sinPhiO;
    maxCos = fmaxf(0.f, dCos);
  }
  float sinAlpha, tanBeta;
  if (AbsCosTheta(WiN) > AbsCosTheta(Wo)) {
    sinAlpha = sinThetaO;
    tanBeta = sinThetaI / AbsCosTheta(WiN);
  } else {
    sinAlpha = sinThetaI;
    tanBeta = sinThetaO / AbsCosTheta(Wo);
  }
  pdf = dot(normalize(Wi), normalize(N)) / PI_F;
  float rA = s.p0, rB = s.p1;
  return s.color * (1.f / PI_F) * (rA + rB * maxCos * sinAlpha * tanBeta);
}
// P/materials/torrance_sparrow.cuh:39-65
inline f3 Evaluate_TS(const Surface& s, f3 Wo, f3 Wi, float& pdf) {
  f3 Rs = s.color;
  float nu = s.p0, nv = s.p1;
  f3 Up = mk(0.f, 1.f, 0.f);
  float NdotI = fmaxf(dot(Up, Wi), 1e-6f), NdotO = fmaxf(dot(Up, Wo), 1e-6f);
  f3 H = Wo + Wi;
  if (isNull(H)) return mk(0.f);
  H = normalize(H);
  float HdotI = fabsf(dot(H, Wi));
  f3 F = schlick3(Rs, HdotI);
  float G = GGX_G(Wo, Wi, nu, nv);
  float D = GGX_D(H, nu, nv);
  pdf = GGX_PDF(H, Wo, nu, nv) / (4.f * dot(Wo, H));
  return Rs * D * G * F / (4.f * NdotI * NdotO);
}
// P/materials/ashikhmin_shirley.cuh:87-151
inline f3 Evaluate_AS(const Surface& s, f3 Wo, f3 Wi, f3 Ns, float& pdf) {
  f3 Rd = s.color, Rs = s.color2;
  float nu = s.p0, nv = s.p1;
  f3 Up = mk(0.f, 1.f, 0.f);
  f3 N = normalize(Ns);
  f3 T = normalize(cross(N, Up));
  f3 B = cross(T, N);
  float NdotI = fabsf(dot(Up, Wi)), NdotO = fabsf(dot(Up, Wo));
  f3 diffuse_component = Rd * (28.f / (23.f * PI_F));
  diffuse_component *= (mk(1.f) - Rs);
  diffuse_component *= (1.f - powf(1.f - NdotI * 0.5f, 5.f));
  diffuse_component *= (1.f - powf(1.f - NdotO * 0.5f, 5.f));
  float diffuse_PDF = dot(Wi, N) / PI_F;
  if (diffuse_PDF < 0.f) return mk(0.f);
  f3 H = Wo + Wi;
  if (isNull(H)) return mk(0.f);
  H = normalize(H);
  float HdotI = fabsf(dot(H, Wi));
  float HdotN = fabsf(dot(H, N)), HdotT = fabsf(dot(H, T)), HdotB = fabsf(dot(H, B));
  float norm, lobe;
  if (nu == nv) {
    norm = (nu + 1.f) / (8.f * PI_F);
    lobe = powf(HdotN, nu);
  } else {
    norm = sqrtf((nu + 1.f) * (nv + 1.f)) / (8.f * PI_F);
    if (HdotN < 1.f)
      lobe = powf(HdotN, (nu * HdotT * HdotT + nv * HdotB * HdotB) / (1.f - HdotN * HdotN));
    else
      lobe = 1.f;
  }
  f3 specular_component = schlick3(Rs, HdotI);
  specular_component *= norm;
  specular_component *= lobe / (HdotI * fmaxf(NdotI, NdotO));
  float specular_PDF = HdotI / (norm * lobe);
  pdf = (diffuse_PDF + specular_PDF) / 2.f;
  return (diffuse_component / diffuse_PDF) + (specular_component / specular_PDF);
}
inline f3 Evaluate(const Surface& s, f3 /*P*/, f3 Wo, f3 Wi, f3 N, float& pdf) {
  switch (s.type) {
    case PT_MAT_LAMBERTIAN:
      return Evaluate_Lambertian(s, Wi, N, pdf);
    case PT_MAT_OREN_NAYAR:
      return Evaluate_ON(s, Wo, Wi, N, pdf);
    case PT_MAT_TORRANCE_SPARROW:
      return Evaluate_TS(s, Wo, Wi, pdf);
    case PT_MAT_ASHIKHMIN_SHIRLEY:
      return Evaluate_AS(s, Wo, Wi, N, pdf);
    case PT_MAT_DIFFUSE_LIGHT:  // P/materials/diffuse_light.cuh:15-22
      pdf = 1.f;
      return s.color;
    case PT_MAT_ISOTROPIC:  // P/materials/isotropic.cuh:15-33
      pdf = 0.25f * PI_F;
      return (0.25f * PI_F) * s.color;
    default:
      return mk(0.f);
  }
}

// ---------------------------------------------------------------------------
// rect light — P/pdfs/rect_pdf.cu:11-146
// ---------------------------------------------------------------------------
inline f3 Rect_Sample(const pt_light& l, f3 P, float ua, float ub) {
  float a = l.a0 + ua * (l.a1 - l.a0);
  float b = l.b0 + ub * (l.b1 - l.b0);
  f3 random_point;
  switch (l.axis) {
    case PT_X_AXIS:
      random_point = mk(l.k, a, b);
      break;
    case PT_Y_AXIS:
      random_point = mk(a, l.k, b);
      break;
    default:
      random_point = mk(a, b, l.k);
      break;
  }
  return random_point - P;
}
inline float Rect_PDF(const pt_light& l, f3 P, f3 Wi) {
  float t, a, b;
  f3 rectNormal;
  switch (l.axis) {
    case PT_X_AXIS:
      t = (l.k - P.x) / Wi.x;
      a = P.y + t * Wi.y;
      b = P.z + t * Wi.z;
      rectNormal = mk(1.f, 0.f, 0.f);
      break;
    case PT_Y_AXIS:
      t = (l.k - P.y) / Wi.y;
      a = P.x + t * Wi.x;
      b = P.z + t * Wi.z;
      rectNormal = mk(0.f, 1.f, 0.f);
      break;
    default:
      t = (l.k - P.z) / Wi.z;
      a = P.x + t * Wi.x;
      b = P.y + t * Wi.y;
      rectNormal = mk(0.f, 0.f, 1.f);
      break;
  }
  if (a < l.a0 || a > l.a1 || b < l.b0 || b > l.b1) return 0.f;
  if (!(t < 3.402823466e+38f && t > 0.001f)) return 0.f;  // (0.001f, FLT_MAX)
  float distance_squared = t * t * squared_length(Wi);
  float cosine = fabsf(dot(Wi, rectNormal) / length(Wi));
  float area = (l.a1 - l.a0) * (l.b1 - l.b0);
  return distance_squared / (cosine * area);
}

// sphere light — P/pdfs/sphere_pdf.cu:8-84 (centre = (a0,a1,b0), radius = k).  As written: the
// sampled cone direction is rotated into a frame built around ITSELF (:79-80), not around the
// direction to the sphere.
inline f3 SphereL_Sample(const pt_light& l, f3 P, uint32_t& seed) {
  f3 center = mk(l.a0, l.a1, l.b0);
  float radius = l.k;
  float r1 = rnd(seed);
  float r2 = rnd(seed);
  float distance_squared = squared_length(center - P);
  float z = 1.f + r2 * (sqrtf(1.f - radius * radius / distance_squared) - 1.f);
  float phi = 2.f * PI_F * r1;
  float x = cosf(phi) * sqrtf(1.f - z * z);
  float y = sinf(phi) * sqrtf(1.f - z * z);
  f3 Wi = mk(x, y, z);
  Onb uvw(normalize(Wi));
  uvw.inverse_transform(Wi);
  return Wi;
}
inline float SphereL_PDF(const pt_light& l, f3 P, f3 Wi) {
  f3 center = mk(l.a0, l.a1, l.b0);
  float radius = l.k;
  const f3 oc = P - center;
  const float a = dot(Wi, Wi);
  const float b = dot(oc, Wi);
  const float c = dot(oc, oc) - radius * radius;
  const float discriminant = b * b - a * c;
  if (discriminant < 0.f) return 0.f;
  const float tmin = 0.001f, tmax = 3.402823466e+38f;
  bool hit = false;
  float temp = (-b - sqrtf(discriminant)) / a;
  if (temp < tmax && temp > tmin) hit = true;
  if (!hit) {
    temp = (-b + sqrtf(discriminant)) / a;
    if (temp < tmax && temp > tmin) hit = true;
  }
  if (!hit) return 0.f;
  float distance_squared = squared_length(center - P);
  float cos_theta_max = sqrtf(1.f - radius * radius / distance_squared);
  float solid_angle = 2.f * PI_F * (1.f - cos_theta_max);
  return 1.f / solid_angle;
}
inline f3 Light_Sample(const pt_light& l, f3 P, uint32_t& seed) {
  if (l.type == PT_LIGHT_SPHERE) return SphereL_Sample(l, P, seed);
  ORC_DRAW2(seed, ua, ub);
  return Rect_Sample(l, P, ua, ub);
}
inline float Light_PDF(const pt_light& l, f3 P, f3 Wi) {
  return l.type == PT_LIGHT_SPHERE ? SphereL_PDF(l, P, Wi) : Rect_PDF(l, P, Wi);
}

inline float PowerHeuristic(unsigned numf, float fPdf, unsigned numg, float gPdf) {
  float f = numf * fPdf;  // P/materials/light_sample.cuh:27-33
  float g = numg * gPdf;
  return (f * f) / (f * f + g * g);
}

// ---------------------------------------------------------------------------
// scene + intersection
// ---------------------------------------------------------------------------
struct Ray {
  f3 o, d;
  float tmin, tmax;
  float time;     // prd.time of the sample (P/raygen.cu:69), read by moving spheres
  uint32_t seed;  // prd.seed when the ray is traced, read by volumes (pinned, see header)
};
struct Box3 {
  f3 lo, hi;
};

inline f3 xf_point(const float* m, f3 p) {
  return mk(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
            m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
inline f3 xf_vector(const float* m, f3 v) {
  return mk(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
            m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
// rtTransformNormal(RT_OBJECT_TO_WORLD, n) = transpose(world->object) * n
inline f3 xf_normal(const float* w2o, f3 n) {
  return mk(w2o[0] * n.x + w2o[4] * n.y + w2o[8] * n.z, w2o[1] * n.x + w2o[5] * n.y + w2o[9] * n.z,
            w2o[2] * n.x + w2o[6] * n.y + w2o[10] * n.z);
}

struct HitRecord {  // P/prd.cuh:32-41
  int index;
  f3 P;
  float u, v;
  f3 geometric_normal, shading_normal, Wo;
};

}  // namespace

struct orc_scene {
  pt_scene_desc d;
  int mode;
  // BVH2 (flat): node i children at left[i], left[i]+1 ; leaf when count>0
  struct Node {
    Box3 box;
    int left, first, count;
  };
  std::vector<Node> nodes;
  std::vector<int> order;
  std::vector<Box3> pbox;
};

namespace {

// --- per-primitive candidate t in (tmin,tmax) ------------------------------
// P/hitables/sphere.cu:32-62
inline bool hit_sphere(const pt_sphere& s, const Ray& r, float& t) {
  const f3 center = ld3(s.center);
  const f3 oc = r.o - center;
  const float a = dot(r.d, r.d);
  const float b = dot(oc, r.d);
  const float c = dot(oc, oc) - s.radius * s.radius;
  const float discriminant = b * b - a * c;
  if (discriminant < 0.f) return false;
  float t0 = (-b - sqrtf(discriminant)) / a;
  if (t0 > r.tmin && t0 < r.tmax) {
    t = t0;
    return true;
  }
  float t1 = (-b + sqrtf(discriminant)) / a;
  if (t1 > r.tmin && t1 < r.tmax) {
    t = t1;
    return true;
  }
  return false;
}
// P/hitables/aarect.cu:20-69
inline bool hit_rect(const pt_rect& q, const Ray& r, float& t) {
  float a, b;
  switch (q.axis) {
    case PT_X_AXIS:
      t = (q.k - r.o.x) / r.d.x;
      a = r.o.y + t * r.d.y;
      b = r.o.z + t * r.d.z;
      break;
    case PT_Y_AXIS:
      t = (q.k - r.o.y) / r.d.y;
      a = r.o.x + t * r.d.x;
      b = r.o.z + t * r.d.z;
      break;
    default:
      t = (q.k - r.o.z) / r.d.z;
      a = r.o.x + t * r.d.x;
      b = r.o.y + t * r.d.y;
      break;
  }
  if (a < q.a0 || a > q.a1 || b < q.b0 || b > q.b1) return false;
  return t > r.tmin && t < r.tmax;
}
// P/hitables/box.cu:36-53
inline bool hit_box(const pt_box& bx, const Ray& r, float& t) {
  f3 t0 = (ld3(bx.bmin) - r.o) / r.d;
  f3 t1 = (ld3(bx.bmax) - r.o) / r.d;
  float tmin = max_component(min_vec(t0, t1));
  float tmax = min_component(max_vec(t0, t1));
  if (tmin <= tmax) {
    if (tmin > r.tmin && tmin < r.tmax) {
      t = tmin;
      return true;
    } else if (tmax > r.tmin && tmax < r.tmax) {
      t = tmax;
      return true;
    }
  }
  return false;
}
// P/hitables/triangle.cu:26-58
inline bool hit_triangle(const float* v9, const Ray& r, float& t, float& bu, float& bv) {
  f3 a = ld3(v9), b = ld3(v9 + 3), c = ld3(v9 + 6);
  f3 e1 = b - a;
  f3 e2 = c - a;
  f3 P = cross(r.d, e2);
  float A = dot(P, e1);
  if ((double)fabsf(A) < 1E-8) return false;
  f3 R = r.o - a;
  float u = dot(P, R) / A;
  if (u < 0.0 || u > 1.0) return false;
  f3 Q = cross(R, e1);
  float v = dot(Q, r.d) / A;
  if (v < 0.0 || u + v > 1.0) return false;
  float tt = dot(Q, e2) / A;
  if (tt > r.tmin && tt < r.tmax) {
    t = tt;
    bu = u;
    bv = v;
    return true;
  }
  return false;
}

// P/hitables/moving_sphere.cu:44-46
inline f3 moving_center(const pt_moving_sphere& q, float time) {
  return ld3(q.center0) + ((time - q.time0) / (q.time1 - q.time0)) * (ld3(q.center1) - ld3(q.center0));
}
// P/hitables/moving_sphere.cu:54-131
inline bool hit_moving_sphere(const pt_moving_sphere& q, const Ray& r, float& t) {
  pt_sphere s;
  st3(s.center, moving_center(q, r.time));
  s.radius = q.radius;
  return hit_sphere(s, r, t);
}
// pinned volume uniform: rnd(tea<4>(seed, prim))
inline float volume_uniform(uint32_t seed, uint32_t prim) {
  uint32_t s = tea<4>(seed, prim);
  return rnd(s);
}
// P/hitables/volume_sphere.cu:20-54
inline bool vol_boundary_sphere(const pt_volume_sphere& q, const Ray& r, float tmin, float tmax, float& rec) {
  pt_sphere s = {{q.center[0], q.center[1], q.center[2]}, q.radius};
  Ray rr = r;
  rr.tmin = tmin;
  rr.tmax = tmax;
  return hit_sphere(s, rr, rec);
}
// P/hitables/volume_box.cu:20-42
inline bool vol_boundary_box(const pt_volume_box& q, const Ray& r, float tmin, float tmax, float& rec) {
  f3 t0 = (ld3(q.bmin) - r.o) / r.d;
  f3 t1 = (ld3(q.bmax) - r.o) / r.d;
  float temp1 = max_component(min_vec(t0, t1));
  float temp2 = min_component(max_vec(t0, t1));
  if (temp1 > temp2) return false;
  if (temp1 < tmax && temp1 > tmin) {
    rec = temp1;
    return true;
  }
  if (temp2 < tmax && temp2 > tmin) {
    rec = temp2;
    return true;
  }
  return false;
}
// P/hitables/volume_sphere.cu:62-105, volume_box.cu:44-82 (as written: the sampled distance is
// not compared with the distance inside the boundary)
template <class Q, class B>
inline bool hit_volume(const Q& q, B boundary, const Ray& r, int prim, float& t) {
  float rec1, rec2;
  if (!boundary(q, r, -FLT_MAX, FLT_MAX, rec1)) return false;
  if (!boundary(q, r, (float)(rec1 + 0.0001), FLT_MAX, rec2)) return false;
  if (rec1 < r.tmin) rec1 = r.tmin;
  if (rec2 > r.tmax) rec2 = r.tmax;
  if (rec1 >= rec2) return false;
  if (rec1 < 0.f) rec1 = 0.f;
  float hit_distance = -(1.f / q.density) * logf(volume_uniform(r.seed, (uint32_t)prim));
  float temp = rec1 + hit_distance / length(r.d);
  if (temp > r.tmin && temp < r.tmax) {
    t = temp;
    return true;
  }
  return false;
}

inline Ray to_object(const orc_scene& s, const pt_prim_ref& p, const Ray& r) {
  if (p.xform < 0) return r;
  const float* m = s.d.xforms[p.xform].w2o;
  Ray o = r;
  o.o = xf_point(m, r.o);
  o.d = xf_vector(m, r.d);
  o.tmin = r.tmin;
  o.tmax = r.tmax;
  return o;
}

inline bool intersect_prim(const orc_scene& s, int id, const Ray& r, float& t, float& u, float& v) {
  const pt_prim_ref& p = s.d.prims[id];
  u = v = 0.f;
  switch (p.type) {
    case PT_PRIM_TRIANGLE:
      return hit_triangle(s.d.tri_v + 9 * (size_t)p.index, r, t, u, v);
    case PT_PRIM_SPHERE:
      return hit_sphere(s.d.spheres[p.index], to_object(s, p, r), t);
    case PT_PRIM_RECT:
      return hit_rect(s.d.rects[p.index], to_object(s, p, r), t);
    case PT_PRIM_BOX:
      return hit_box(s.d.boxes[p.index], to_object(s, p, r), t);
    case PT_PRIM_MOVING_SPHERE:
      return hit_moving_sphere(s.d.moving_spheres[p.index], to_object(s, p, r), t);
    case PT_PRIM_VOLUME_SPHERE:
      return hit_volume(s.d.volume_spheres[p.index], vol_boundary_sphere, to_object(s, p, r), id, t);
    case PT_PRIM_VOLUME_BOX:
      return hit_volume(s.d.volume_boxes[p.index], vol_boundary_box, to_object(s, p, r), id, t);
  }
  return false;
}

// --- world bounds (generously padded; the BVH must never cull a primitive the
// brute-force loop would hit) -------------------------------------------------
inline Box3 prim_bounds(const orc_scene& s, int id) {
  const pt_prim_ref& p = s.d.prims[id];
  f3 lo, hi;
  switch (p.type) {
    case PT_PRIM_TRIANGLE: {  // P/hitables/triangle.cu:61-83
      const float* v = s.d.tri_v + 9 * (size_t)p.index;
      lo = min_vec(min_vec(ld3(v), ld3(v + 3)), ld3(v + 6));
      hi = max_vec(max_vec(ld3(v), ld3(v + 3)), ld3(v + 6));
      break;
    }
    case PT_PRIM_SPHERE: {  // P/hitables/sphere.cu:96-100
      const pt_sphere& q = s.d.spheres[p.index];
      lo = ld3(q.center) - q.radius;
      hi = ld3(q.center) + q.radius;
      break;
    }
    case PT_PRIM_RECT: {  // P/hitables/aarect.cu:117-136
      const pt_rect& q = s.d.rects[p.index];
      if (q.axis == PT_X_AXIS) {
        lo = mk(q.k, q.a0, q.b0);
        hi = mk(q.k, q.a1, q.b1);
      } else if (q.axis == PT_Y_AXIS) {
        lo = mk(q.a0, q.k, q.b0);
        hi = mk(q.a1, q.k, q.b1);
      } else {
        lo = mk(q.a0, q.b0, q.k);
        hi = mk(q.a1, q.b1, q.k);
      }
      break;
    }
    case PT_PRIM_MOVING_SPHERE: {  // P/hitables/moving_sphere.cu:136-147 (+ camera shutter, see header)
      const pt_moving_sphere& q = s.d.moving_spheres[p.index];
      const float ts[4] = {q.time0, q.time1, s.d.camera.time0, s.d.camera.time1};
      lo = mk(INFINITY);
      hi = mk(-INFINITY);
      for (float tt : ts) {
        f3 c = moving_center(q, tt);
        if (!(c.x == c.x && c.y == c.y && c.z == c.z)) continue;
        lo = min_vec(lo, c - q.radius);
        hi = max_vec(hi, c + q.radius);
      }
      break;
    }
    case PT_PRIM_VOLUME_SPHERE: {  // P/hitables/volume_sphere.cu:111-115
      const pt_volume_sphere& q = s.d.volume_spheres[p.index];
      lo = ld3(q.center) - q.radius;
      hi = ld3(q.center) + q.radius;
      break;
    }
    case PT_PRIM_VOLUME_BOX: {  // P/hitables/volume_box.cu:86-90
      const pt_volume_box& q = s.d.volume_boxes[p.index];
      lo = min_vec(ld3(q.bmin), ld3(q.bmax));
      hi = max_vec(ld3(q.bmin), ld3(q.bmax));
      break;
    }
    default: {  // P/hitables/box.cu:84-88
      const pt_box& q = s.d.boxes[p.index];
      lo = ld3(q.bmin);
      hi = ld3(q.bmax);
      break;
    }
  }
  if (p.xform >= 0) {
    const float* m = s.d.xforms[p.xform].o2w;
    f3 nlo = mk(INFINITY), nhi = mk(-INFINITY);
    for (int c = 0; c < 8; c++) {
      f3 q = xf_point(m, mk((c & 1) ? hi.x : lo.x, (c & 2) ? hi.y : lo.y, (c & 4) ? hi.z : lo.z));
      nlo = min_vec(nlo, q);
      nhi = max_vec(nhi, q);
    }
    lo = nlo;
    hi = nhi;
  }
  // padding: reference pads 1e-4..1e-3 absolute; add a relative term so the
  // slab test stays conservative for large coordinates
  f3 m = max_vec(mk(fabsf(lo.x), fabsf(lo.y), fabsf(lo.z)), mk(fabsf(hi.x), fabsf(hi.y), fabsf(hi.z)));
  f3 pad = mk(1e-3f) + m * 1e-5f;
  return {lo - pad, hi + pad};
}

inline Box3 merge(const Box3& a, const Box3& b) { return {min_vec(a.lo, b.lo), max_vec(a.hi, b.hi)}; }
inline float half_area(const Box3& b) {
  f3 e = b.hi - b.lo;
  return e.x * e.y + e.y * e.z + e.z * e.x;
}

void build_bvh(orc_scene& s) {
  int n = s.d.n_prims;
  s.pbox.resize(n);
  s.order.resize(n);
  std::vector<f3> cen(n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    s.pbox[i] = prim_bounds(s, i);
    s.order[i] = i;
    cen[i] = (s.pbox[i].lo + s.pbox[i].hi) * 0.5f;
  }
  s.nodes.clear();
  s.nodes.reserve(2 * (size_t)n + 2);
  s.nodes.push_back({});
  struct Job {
    int node, first, count;
  };
  std::vector<Job> stack;
  stack.push_back({0, 0, n});
  const int NB = 16;
  while (!stack.empty()) {
    Job j = stack.back();
    stack.pop_back();
    Box3 bb = {mk(INFINITY), mk(-INFINITY)}, cb = bb;
    for (int i = j.first; i < j.first + j.count; i++) {
      bb = merge(bb, s.pbox[s.order[i]]);
      f3 c = cen[s.order[i]];
      cb.lo = min_vec(cb.lo, c);
      cb.hi = max_vec(cb.hi, c);
    }
    orc_scene::Node& nd = s.nodes[j.node];
    nd.box = bb;
    nd.first = j.first;
    nd.count = 0;
    nd.left = -1;
    if (j.count <= 4) {
      nd.count = j.count;
      continue;
    }
    // binned SAH over the widest centroid axis candidates
    int bestAxis = -1, bestBin = -1;
    float bestCost = INFINITY;
    for (int ax = 0; ax < 3; ax++) {
      float lo = (&cb.lo.x)[ax], hi = (&cb.hi.x)[ax];
      if (!(hi > lo)) continue;
      Box3 bins[NB];
      int cnt[NB];
      for (int b = 0; b < NB; b++) {
        bins[b] = {mk(INFINITY), mk(-INFINITY)};
        cnt[b] = 0;
      }
      float scale = NB / (hi - lo);
      for (int i = j.first; i < j.first + j.count; i++) {
        int id = s.order[i];
        int b = std::min(NB - 1, std::max(0, (int)(((&cen[id].x)[ax] - lo) * scale)));
        bins[b] = merge(bins[b], s.pbox[id]);
        cnt[b]++;
      }
      float rightArea[NB];
      int rightCnt[NB];
      Box3 acc = {mk(INFINITY), mk(-INFINITY)};
      int c = 0;
      for (int b = NB - 1; b > 0; b--) {
        acc = merge(acc, bins[b]);
        c += cnt[b];
        rightArea[b] = c ? half_area(acc) : 0.f;
        rightCnt[b] = c;
      }
      acc = {mk(INFINITY), mk(-INFINITY)};
      c = 0;
      for (int b = 0; b < NB - 1; b++) {
        acc = merge(acc, bins[b]);
        c += cnt[b];
        if (c == 0 || rightCnt[b + 1] == 0) continue;
        float cost = half_area(acc) * c + rightArea[b + 1] * rightCnt[b + 1];
        if (cost < bestCost) {
          bestCost = cost;
          bestAxis = ax;
          bestBin = b;
        }
      }
    }
    int mid;
    if (bestAxis < 0) {
      mid = j.first + j.count / 2;
    } else {
      float lo = (&cb.lo.x)[bestAxis], hi = (&cb.hi.x)[bestAxis];
      float scale = NB / (hi - lo);
      int* beg = s.order.data() + j.first;
      int* p = std::partition(beg, beg + j.count, [&](int id) {
        int b = std::min(NB - 1, std::max(0, (int)(((&cen[id].x)[bestAxis] - lo) * scale)));
        return b <= bestBin;
      });
      mid = (int)(p - s.order.data());
      if (mid == j.first || mid == j.first + j.count) mid = j.first + j.count / 2;
    }
    int left = (int)s.nodes.size();
    s.nodes.push_back({});
    s.nodes.push_back({});
    s.nodes[j.node].left = left;
    stack.push_back({left, j.first, mid - j.first});
    stack.push_back({left + 1, mid, j.first + j.count - mid});
  }
}

inline bool slab(const Box3& b, const Ray& r, f3 inv, float tbest) {
  float t0x = (b.lo.x - r.o.x) * inv.x, t1x = (b.hi.x - r.o.x) * inv.x;
  float t0y = (b.lo.y - r.o.y) * inv.y, t1y = (b.hi.y - r.o.y) * inv.y;
  float t0z = (b.lo.z - r.o.z) * inv.z, t1z = (b.hi.z - r.o.z) * inv.z;
  float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), r.tmin));
  float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tbest));
  // conservative: widen by a few ulp (Ize, "Robust BVH ray traversal")
  return tn <= tf * 1.0000004f + 1e-30f || !(tn == tn) || !(tf == tf);
}

struct Hit {
  float t;
  int prim;
  float u, v;
};

inline Hit closest_hit(const orc_scene& s, const Ray& r, uint64_t* nodeTests, uint64_t* primTests) {
  Hit best = {r.tmax, -1, 0.f, 0.f};
  auto consider = [&](int id) {
    float t, u, v;
    if (intersect_prim(s, id, r, t, u, v)) {
      if (t < best.t || (t == best.t && (best.prim < 0 || id < best.prim))) best = {t, id, u, v};
    }
  };
  if (s.mode == 1) {
    for (int i = 0; i < s.d.n_prims; i++) consider(i);
    if (primTests) *primTests += s.d.n_prims;
    return best;
  }
  auto safe_inv = [](float d) {
    const float eps = 1e-30f;
    if (fabsf(d) < eps) d = (d < 0.f || (d == 0.f && std::signbit(d))) ? -eps : eps;
    return 1.0f / d;
  };
  f3 inv = mk(safe_inv(r.d.x), safe_inv(r.d.y), safe_inv(r.d.z));
  int stack[96];
  int sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const orc_scene::Node& nd = s.nodes[stack[--sp]];
    if (nodeTests) ++*nodeTests;
    if (!slab(nd.box, r, inv, best.t)) continue;
    if (nd.count > 0) {
      for (int i = 0; i < nd.count; i++) consider(s.order[nd.first + i]);
      if (primTests) *primTests += nd.count;
    } else {
      stack[sp++] = nd.left;
      stack[sp++] = nd.left + 1;
    }
  }
  return best;
}

// --- hit records -------------------------------------------------------------
inline HitRecord get_hit_record(const orc_scene& s, const Ray& wr, const Hit& h) {
  const pt_prim_ref& p = s.d.prims[h.prim];
  HitRecord rec;
  rec.Wo = normalize(-wr.d);
  rec.index = 0;
  const float* o2w = p.xform >= 0 ? s.d.xforms[p.xform].o2w : nullptr;
  const float* w2o = p.xform >= 0 ? s.d.xforms[p.xform].w2o : nullptr;
  Ray r = to_object(s, p, wr);
  auto toWorldP = [&](f3 q) { return o2w ? xf_point(o2w, q) : q; };
  auto toWorldN = [&](f3 n) { return normalize(w2o ? xf_normal(w2o, n) : n); };
  switch (p.type) {
    case PT_PRIM_VOLUME_SPHERE:
    case PT_PRIM_VOLUME_BOX: {  // P/hitables/volume_sphere.cu:86-103, volume_box.cu:61-78
      f3 hit_point = r.o + h.t * r.d;
      rec.P = toWorldP(hit_point);
      rec.shading_normal = rec.geometric_normal = toWorldN(mk(1.f, 0.f, 0.f));
      rec.u = rec.v = 0.f;
      break;
    }
    case PT_PRIM_SPHERE:
    case PT_PRIM_MOVING_SPHERE: {  // P/hitables/sphere.cu:65-93, moving_sphere.cu:82-100
      pt_sphere q;
      if (p.type == PT_PRIM_SPHERE) {
        q = s.d.spheres[p.index];
      } else {
        const pt_moving_sphere& ms = s.d.moving_spheres[p.index];
        st3(q.center, moving_center(ms, wr.time));
        q.radius = ms.radius;
      }
      f3 hit_point = r.o + h.t * r.d;
      rec.P = toWorldP(hit_point);
      f3 T = (hit_point - ld3(q.center)) / q.radius;
      f3 normal = toWorldN(T);
      rec.shading_normal = rec.geometric_normal = normal;
      float phi = atan2f(T.z, T.x);
      float theta = asinf(T.y);
      rec.u = 1.f - (phi + PI_F) / (2.f * PI_F);
      rec.v = (theta + PI_F / 2.f) / PI_F;
      break;
    }
    case PT_PRIM_RECT: {  // P/hitables/aarect.cu:71-115
      const pt_rect& q = s.d.rects[p.index];
      f3 hit_point = r.o + h.t * r.d;
      rec.P = toWorldP(hit_point);
      f3 normal;
      if (q.axis == PT_X_AXIS) {
        normal = mk(1.f, 0.f, 0.f);
        rec.u = (hit_point.y - q.a0) / (q.a1 - q.a0);
        rec.v = (hit_point.z - q.b0) / (q.b1 - q.b0);
      } else if (q.axis == PT_Y_AXIS) {
        normal = mk(0.f, 1.f, 0.f);
        rec.u = (hit_point.x - q.a0) / (q.a1 - q.a0);
        rec.v = (hit_point.z - q.b0) / (q.b1 - q.b0);
      } else {
        normal = mk(0.f, 0.f, 1.f);
        rec.u = (hit_point.x - q.a0) / (q.a1 - q.a0);
        rec.v = (hit_point.y - q.b0) / (q.b1 - q.b0);
      }
      normal = q.flip ? -normal : normal;
      normal = toWorldN(normal);
      rec.shading_normal = rec.geometric_normal = normal;
      break;
    }
    case PT_PRIM_BOX: {  // P/hitables/box.cu:24-33,56-81
      const pt_box& q = s.d.boxes[p.index];
      f3 hit_point = r.o + h.t * r.d;
      rec.P = toWorldP(hit_point);
      f3 t0 = (ld3(q.bmin) - r.o) / r.d;
      f3 t1 = (ld3(q.bmax) - r.o) / r.d;
      float t = h.t;
      f3 neg = mk(t == t0.x ? 1 : 0, t == t0.y ? 1 : 0, t == t0.z ? 1 : 0);
      f3 pos = mk(t == t1.x ? 1 : 0, t == t1.y ? 1 : 0, t == t1.z ? 1 : 0);
      f3 normal = toWorldN(pos - neg);
      rec.shading_normal = rec.geometric_normal = normal;
      rec.u = rec.v = 0.f;
      break;
    }
    default: {  // P/hitables/triangle.cu:85-149 (identity transform: world-space mesh)
      const float* v9 = s.d.tri_v + 9 * (size_t)p.index;
      f3 a = ld3(v9), b = ld3(v9 + 3), c = ld3(v9 + 6);
      f3 e1 = b - a, e2 = c - a;
      float b0 = 1.f - h.u - h.v, b1 = h.u, b2 = h.v;
      rec.P = a * b0 + b * b1 + c * b2;
      f3 Ng = cross(e1, e2);
      float area = length(Ng) / 2.f;
      Ng /= 2.f * area;
      Ng = normalize(normalize(Ng));
      rec.geometric_normal = Ng;
      if (!s.d.tri_n) {
        rec.shading_normal = Ng;
      } else {
        const float* n9 = s.d.tri_n + 9 * (size_t)p.index;
        f3 Ns = ld3(n9) * b0 + ld3(n9 + 3) * b1 + ld3(n9 + 6) * b2;
        rec.shading_normal = normalize(Ns);
      }
      if (!s.d.tri_uv) {
        rec.u = rec.v = 0.f;
      } else {
        const float* t6 = s.d.tri_uv + 6 * (size_t)p.index;
        rec.u = t6[0] * b0 + t6[2] * b1 + t6[4] * b2;
        rec.v = t6[1] * b0 + t6[3] * b1 + t6[5] * b2;
      }
      rec.index = s.d.tri_tex ? s.d.tri_tex[p.index] : 0;
      break;
    }
  }
  return rec;
}

// --- textures — P/textures/*.cu ---------------------------------------------
inline f3 image_lookup(const pt_image& im, float u, float v) {
  if (!(std::isfinite(u) && std::isfinite(v)) || im.width <= 0 || im.height <= 0) return mk(0.f);
  float x = u * (float)im.width - 0.5f;
  float y = v * (float)im.height - 0.5f;
  float fx0 = floorf(x), fy0 = floorf(y);
  float fx = x - fx0, fy = y - fy0;
  // repeat addressing on (possibly huge) float coordinates
  auto wrap = [](float f, int n) {
    float m = fmodf(f, (float)n);
    int i = (int)m;
    if (i < 0) i += n;
    if (i >= n) i -= n;
    return i;
  };
  int x0 = wrap(fx0, im.width), y0 = wrap(fy0, im.height);
  int x1 = x0 + 1 == im.width ? 0 : x0 + 1;
  int y1 = y0 + 1 == im.height ? 0 : y0 + 1;
  auto T = [&](int xx, int yy) { return ld3(im.rgba + 4 * ((size_t)yy * im.width + xx)); };
  f3 c00 = T(x0, y0), c10 = T(x1, y0), c01 = T(x0, y1), c11 = T(x1, y1);
  f3 top = c00 + fx * (c10 - c00);
  f3 bot = c01 + fx * (c11 - c01);
  return top + fy * (bot - top);
}

// P/textures/noise_texture.cu:12-60 (ranvec / perm_x / perm_y / perm_z live in the texture-data table)
inline float perlin_interp(f3 c[2][2][2], float u, float v, float w) {
  float uu = u * u * (3 - 2 * u);
  float vv = v * v * (3 - 2 * v);
  float ww = w * w * (3 - 2 * w);
  float accum = 0;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      for (int k = 0; k < 2; k++) {
        f3 weight_v = mk(u - i, v - j, w - k);
        accum += (i * uu + (1 - i) * (1 - uu)) * (j * vv + (1 - j) * (1 - vv)) * (k * ww + (1 - k) * (1 - ww)) *
                 dot(c[i][j][k], weight_v);
      }
  return accum;
}
inline float noise(const float* data, f3 p) {
  float u = p.x - floorf(p.x);
  float v = p.y - floorf(p.y);
  float w = p.z - floorf(p.z);
  int i = (int)floorf(p.x);
  int j = (int)floorf(p.y);
  int k = (int)floorf(p.z);
  f3 c[2][2][2];
  for (int di = 0; di < 2; di++)
    for (int dj = 0; dj < 2; dj++)
      for (int dk = 0; dk < 2; dk++) {
        int idx = (int)data[768 + ((i + di) & 255)] ^ (int)data[1024 + ((j + dj) & 255)] ^ (int)data[1280 + ((k + dk) & 255)];
        c[di][dj][dk] = ld3(data + 3 * idx);
      }
  return perlin_interp(c, u, v, w);
}
inline float turb(const float* data, f3 p) {
  float accum = 0;
  f3 temp_p = p;
  float weight = 1.0f;
  for (int i = 0; i < 7; i++) {
    accum += weight * noise(data, temp_p);
    weight *= 0.5f;
    temp_p *= 2.f;
  }
  return fabsf(accum);
}

inline f3 sample_texture(const orc_scene& s, int tex, float u, float v, f3 p, int i) {
  for (int guard = 0; guard < 8; guard++) {
    if (tex < 0 || tex >= s.d.n_textures) return mk(0.f);
    const pt_texture& t = s.d.textures[tex];
    switch (t.type) {
      case PT_TEX_CONSTANT:  // constant_texture.cu:5-7
        return ld3(t.color);
      case PT_TEX_CHECKER: {  // checkered_texture.cu:8-15
        float sines = sinf(10 * p.x) * sinf(10 * p.y) * sinf(10 * p.z);
        tex = (sines < 0) ? t.a : t.b;
        i = 0;
        break;
      }
      case PT_TEX_IMAGE:  // image_texture.cu:3-7
        if (t.a < 0 || t.a >= s.d.n_images) return mk(0.f);
        return image_lookup(s.d.images[t.a], u, v);
      case PT_TEX_VECTOR:  // vector.cu:7-12
        if (i >= t.b || i < 0) return mk(0.f);
        tex = s.d.texvec[t.a + i];
        i = 0;
        break;
      case PT_TEX_NOISE: {  // noise_texture.cu:62-75: the axis switch falls through to the Z case
        const float scale = t.color[0];
        float sinValue = sinf(scale * p.z + 5 * turb(s.d.texdata + t.a, scale * p));
        return mk(1.f) * 0.5f * (1 + sinValue);
      }
      case PT_TEX_GRADIENT: {  // gradient_texture.cu:7-14
        const float* g = s.d.texdata + t.a;
        const f3 unit_direction = normalize(p);
        const float x = fabsf(unit_direction.x), y = fabsf(unit_direction.y), z = fabsf(unit_direction.z);
        return x * ld3(g) + y * ld3(g + 3) + z * ld3(g + 6);
      }
      default:
        return mk(0.f);
    }
  }
  return mk(0.f);
}

// --- path state ---------------------------------------------------------------
enum ScatterEvent { rayGotBounced, rayGotCancelled, rayHitLight, rayMissed };  // P/prd.cuh:24-29
struct PerRayData {  // P/prd.cuh:45-57
  uint32_t seed;
  float time;
  f3 throughput, radiance;
  ScatterEvent scatterEvent;
  bool isSpecular;
  f3 origin, direction;
  float last_pdf;  // env NEE mode only: BSDF pdf of the direction that produced the current ray
};
struct Counters {
  uint64_t closest = 0, shadow = 0, bounces = 0, nodes = 0, prims = 0;
  // probe support (orc_closest_hit_probe): the last shadow ray Direct_Light traced and its outcome
  int shadow_traced = 0, shadow_blocked = 0;
  float shadow_ray[6] = {0, 0, 0, 0, 0, 0};
};

f3 miss_color(const orc_scene& s, f3 d);
inline bool env_nee_active(const orc_scene& s);
inline f3 env_sample_dir(const orc_scene& s, uint32_t& seed, float& pdf);

// P/materials/light_sample.cuh:35-103
f3 Direct_Light(const orc_scene& s, const Surface& surface, f3 P, f3 Wo, f3 N, bool isLight,
                uint32_t& seed, Counters& cn, float time) {
  f3 directLight = mk(0.f);
  int numLights = s.d.n_lights;
  if (numLights == 0) return mk(0.f);
  int index = ((int)(rnd(seed) * numLights)) % numLights;
  if (isLight && numLights == 1) return mk(0.f);
  const pt_light& L = s.d.lights[index];
  f3 emission = ld3(L.emission);
  f3 Wi = Light_Sample(L, P, seed);
  float lightPDF = Light_PDF(L, P, Wi);
  if (dot(Wi, N) < 0.f) return mk(0.f);
  // shadow ray (type 1, tmin 1e-3, tmax inf) — pinned visibility, see header
  Ray sh = {P, Wi, 1e-3f, RT_DEFAULT_MAX, time, seed};
  Hit h = closest_hit(s, sh, &cn.nodes, &cn.prims);
  cn.shadow++;
  bool inShadow = true;
  if (h.prim >= 0) {
    const pt_material& m = s.d.materials[s.d.prims[h.prim].material];
    if (m.type == PT_MAT_DIFFUSE_LIGHT && dot(N, Wi) > 0.f) inShadow = false;
  }
  cn.shadow_traced++;
  cn.shadow_blocked = inShadow ? 1 : 0;
  st3(cn.shadow_ray, P);
  st3(cn.shadow_ray + 3, Wi);
  if (inShadow) return mk(0.f);
  if (lightPDF != 0.f && !isNull(emission)) {
    float matPDF = 0.f;
    f3 matValue = Evaluate(surface, P, Wo, Wi, N, matPDF);
    if (matPDF != 0.f && !isNull(matValue)) {
      float weight = PowerHeuristic(1, lightPDF, 1, matPDF);
      directLight += matValue * emission * weight / lightPDF;
    }
  }
  Wi = Sample(surface, P, Wo, N, seed);
  float matPDF = 0.f;
  f3 matValue = Evaluate(surface, P, Wo, Wi, N, matPDF);
  if (matPDF != 0.f && !isNull(matValue)) {
    lightPDF = Light_PDF(L, P, Wi);
    if (!lightPDF || isNull(emission)) return directLight;
    float weight = PowerHeuristic(1, matPDF, 1, lightPDF);
    directLight += matValue * emission * weight / matPDF;
  }
  return directLight;
}

// closest-hit programs — P/materials/*.cu
void closest_hit_program(const orc_scene& s, const Ray& ray, const Hit& h, PerRayData& prd,
                         Counters& cn) {
  HitRecord rec = get_hit_record(s, ray, h);
  const pt_material& m = s.d.materials[s.d.prims[h.prim].material];
  int index = rec.index;
  f3 P = rec.P, Wo = rec.Wo, N = rec.shading_normal;
  Surface surf;
  surf.type = m.type;
  surf.p0 = m.p0;
  surf.p1 = m.p1;
  surf.color = surf.color2 = mk(0.f);
  switch (m.type) {
    case PT_MAT_LAMBERTIAN:   // lambertian.cu:49-73
    case PT_MAT_OREN_NAYAR: {  // oren_nayar.cu:46-70
      surf.color = sample_texture(s, m.tex0, rec.u, rec.v, P, index);
      if (env_nee_active(s)) {  // NEW estimator (see env_* above)
        float lpdf;
        f3 WiL = env_sample_dir(s, prd.seed, lpdf);
        f3 direct = mk(0.f);
        bool need_shadow = false;
        if (dot(WiL, N) > 0.f && lpdf > 0.f) {
          float mpdf = 0.f;
          f3 f = Evaluate(surf, P, Wo, WiL, N, mpdf);
          if (mpdf != 0.f && !isNull(f)) {
            float w = PowerHeuristic(1, lpdf, 1, mpdf);
            direct = f * miss_color(s, WiL) * w / lpdf;
            need_shadow = true;
          }
        }
        const uint32_t shadow_seed = prd.seed;  // seed when the shadow ray is decided (after the 4 light draws)
        f3 Wi = Sample(surf, P, Wo, N, prd.seed);
        float pdf = 0.f;
        f3 attenuation = Evaluate(surf, P, Wo, Wi, N, pdf);
        if (need_shadow) {
          Ray sh = {P, WiL, 1e-3f, RT_DEFAULT_MAX, prd.time, shadow_seed};
          Hit hh = closest_hit(s, sh, &cn.nodes, &cn.prims);
          cn.shadow++;
          if (hh.prim < 0) prd.radiance = prd.radiance + prd.throughput * direct;
        }
        prd.scatterEvent = rayGotBounced;
        prd.origin = P;
        prd.direction = Wi;
        prd.throughput = prd.throughput * clamp3(attenuation / pdf, 0.f, 1.f);
        prd.isSpecular = false;
        prd.last_pdf = pdf;
        break;
      }
      f3 direct = Direct_Light(s, surf, P, Wo, N, false, prd.seed, cn, prd.time);
      prd.radiance += prd.throughput * direct;
      f3 Wi = Sample(surf, P, Wo, N, prd.seed);
      float pdf = 0.f;
      f3 attenuation = Evaluate(surf, P, Wo, Wi, N, pdf);
      prd.scatterEvent = rayGotBounced;
      prd.origin = P;
      prd.direction = Wi;
      prd.throughput *= clamp3(attenuation / pdf, 0.f, 1.f);
      prd.isSpecular = false;
      break;
    }
    case PT_MAT_ISOTROPIC: {  // isotropic.cu:34-60
      surf.color = sample_texture(s, m.tex0, rec.u, rec.v, P, index);
      f3 direct = Direct_Light(s, surf, P, Wo, N, false, prd.seed, cn, prd.time);
      prd.radiance += prd.throughput * direct;
      f3 Wi = Sample(surf, P, Wo, N, prd.seed);
      float pdf = 0.f;
      f3 attenuation = Evaluate(surf, P, Wo, Wi, N, pdf);
      prd.scatterEvent = rayGotBounced;
      prd.origin = P;
      prd.direction = Wi;
      prd.throughput *= attenuation / pdf;
      prd.isSpecular = false;
      break;
    }
    case PT_MAT_METAL: {  // metal.cu:38-56
      f3 color = sample_texture(s, m.tex0, rec.u, rec.v, P, index);
      f3 reflected = reflect(-Wo, N);
      prd.direction = reflected + m.p0 * random_in_unit_sphere(prd.seed);
      prd.scatterEvent = rayGotBounced;
      prd.origin = P;
      prd.throughput *= color;
      prd.isSpecular = true;
      break;
    }
    case PT_MAT_DIELECTRIC: {  // dielectric.cu:46-96
      f3 Wd = -rec.Wo;
      f3 base_color = sample_texture(s, m.tex0, rec.u, rec.v, P, index);
      f3 absorption = mk(1.f);
      float ref_idx = m.p0;
      float ni_over_nt;
      float cosine = dot(Wd, N);
      if (cosine > 0.f) {
        N = -N;
        ni_over_nt = ref_idx;
        cosine = ref_idx * cosine / length(Wd);
      } else {
        ni_over_nt = 1.f / ref_idx;
        cosine = -cosine / length(Wd);
      }
      f3 refracted = mk(0.f);
      float reflect_prob;
      if (Refract(Wd, N, ni_over_nt, refracted))
        reflect_prob = schlick(cosine, ref_idx);
      else
        reflect_prob = 1.f;
      if (rnd(prd.seed) < reflect_prob)
        prd.direction = reflect(Wd, N);
      else
        prd.direction = normalize(refracted);
      prd.scatterEvent = rayGotBounced;
      prd.origin = P;
      prd.throughput *= (base_color * absorption);
      prd.isSpecular = true;
      break;
    }
    case PT_MAT_DIFFUSE_LIGHT: {  // diffuse_light.cu:30-48
      surf.color = sample_texture(s, m.tex0, rec.u, rec.v, P, index);
      f3 direct = Direct_Light(s, surf, P, Wo, N, true, prd.seed, cn, prd.time);
      prd.radiance += prd.throughput * direct;
      if (dot(N, Wo) < 0.f) prd.throughput *= surf.color;
      prd.scatterEvent = rayHitLight;
      break;
    }
    case PT_MAT_TORRANCE_SPARROW: {  // torrance_sparrow.cu:39-59
      surf.color = sample_texture(s, m.tex0, rec.u, rec.v, P, index);
      f3 Wi = Sample(surf, P, Wo, N, prd.seed);
      float pdf = 0.f;
      f3 attenuation = Evaluate(surf, P, Wo, Wi, N, pdf);
      prd.scatterEvent = rayGotBounced;
      prd.origin = P;
      prd.direction = Wi;
      prd.throughput *= clamp3(attenuation / pdf, 0.f, 1.f);
      prd.isSpecular = true;
      break;
    }
    case PT_MAT_ASHIKHMIN_SHIRLEY: {  // ashikhmin_shirley.cu:48-68
      surf.color = sample_texture(s, m.tex0, rec.u, rec.v, P, index);
      surf.color2 = sample_texture(s, m.tex1, rec.u, rec.v, P, index);
      f3 Wi = Sample(surf, P, Wo, N, prd.seed);
      float pdf = 0.f;
      f3 attenuation = Evaluate(surf, P, Wo, Wi, N, pdf);
      prd.scatterEvent = rayGotBounced;
      prd.origin = P;
      prd.direction = Wi;
      prd.throughput *= clamp3(attenuation, 0.f, 1.f);
      prd.isSpecular = true;
      break;
    }
    default: {  // normal_shader.cu:22-35
      f3 n = (m.flags & 1) ? rec.shading_normal : rec.geometric_normal;
      prd.radiance = n * 0.5f + mk(0.5f);
      prd.scatterEvent = rayGotCancelled;
      break;
    }
  }
}

// texture coordinates of the image backgrounds — P/miss.cu:51-55 (image_background), :64-91 (environmental_mapping)
inline void miss_uv(const pt_miss& ms, f3 d, float& u, float& v) {
  if (ms.type == PT_MISS_IMG) {
    float theta = atan2f(d.x, d.z);
    float phi = OPTIX_PIf * 0.5f - acosf(d.y);
    u = (theta + OPTIX_PIf) * (0.5f * OPTIX_1_PIf);
    v = 0.5f * (1.f + sinf(phi));
  } else if (ms.spherical) {
    float r = length(d);
    float lon = atan2f(d.z, d.x);
    float lat = acosf(d.y / r);
    float rx = 1.f / (PI_F * 2.f), ry = 1.f / PI_F;
    u = lon * rx;
    v = lat * ry;
  } else {
    float theta = atan2f(d.x, d.z);
    theta = theta < 0.f ? theta + (2.f * PI_F) : theta;
    float phi = acosf(d.y);
    u = 1.f - (theta / (2.f * PI_F));
    v = phi / PI_F;
  }
}

// miss programs — P/miss.cu:30-95: background radiance factor for direction d
f3 miss_color(const orc_scene& s, f3 d) {
  const pt_miss& ms = s.d.miss;
  switch (ms.type) {
    case PT_MISS_GRADIENT: {
      const f3 unit_direction = normalize(d);
      const float t = 0.5f * (unit_direction.y + 1.f);
      f3 c = (1.f - t) * ld3(ms.color1);
      c += t * ld3(ms.color2);
      return c;
    }
    case PT_MISS_CONSTANT:
      return ld3(ms.color1);
    case PT_MISS_IMG: {
      float u, v;
      miss_uv(ms, d, u, v);
      return (ms.image >= 0 && ms.image < s.d.n_images) ? image_lookup(s.d.images[ms.image], u, v) : mk(0.f);
    }
    default: {  // HDR environmental_mapping
      float u, v;
      miss_uv(ms, d, u, v);
      f3 c = (ms.image >= 0 && ms.image < s.d.n_images) ? image_lookup(s.d.images[ms.image], u, v) : mk(0.f);
      return 2.f * c;
    }
  }
}

// ---- environment importance sampling (NEW mode pt_miss.env_nee; own definition, mirrored by
// the product's shading.h): alias table over texels, power-heuristic MIS, no clamp at miss ----
inline bool env_nee_active(const orc_scene& s) { return s.d.env_n > 0 && s.d.n_lights == 0; }
inline f3 env_dir_from_uv(float u, float v) {
  float lon = u * (2.f * PI_F), lat = v * PI_F;
  float st = sinf(lat);
  return mk(st * cosf(lon), cosf(lat), st * sinf(lon));
}
inline float env_pdf_dir(const orc_scene& s, f3 d) {
  const pt_image& im = s.d.images[s.d.miss.image];
  float r = length(d);
  float lon = atan2f(d.z, d.x);
  float lat = acosf(d.y / r);
  float u = lon * (1.f / (PI_F * 2.f));
  if (u < 0.f) u += 1.f;
  float v = lat * (1.f / PI_F);
  int x = (int)(u * (float)im.width), y = (int)(v * (float)im.height);
  x = x < 0 ? 0 : (x > im.width - 1 ? im.width - 1 : x);
  y = y < 0 ? 0 : (y > im.height - 1 ? im.height - 1 : y);
  float st = sinf(lat);
  if (!(st > 1e-6f)) return 0.f;
  return s.d.env_texel_pdf[(size_t)y * im.width + x] * ((float)s.d.env_n / (2.f * PI_F * PI_F * st));
}
inline f3 env_sample_dir(const orc_scene& s, uint32_t& seed, float& pdf) {
  const pt_image& im = s.d.images[s.d.miss.image];
  float u0 = rnd(seed), u1 = rnd(seed), u2 = rnd(seed), u3 = rnd(seed);
  int i = (int)(u0 * (float)s.d.env_n);
  i = i > s.d.env_n - 1 ? s.d.env_n - 1 : i;
  if (u1 >= s.d.env_alias_prob[i]) i = s.d.env_alias_idx[i];
  int x = i % im.width, y = i / im.width;
  float u = ((float)x + u2) / (float)im.width, v = ((float)y + u3) / (float)im.height;
  float st = sinf(v * PI_F);
  pdf = st > 1e-6f ? s.d.env_texel_pdf[i] * ((float)s.d.env_n / (2.f * PI_F * PI_F * st)) : 0.f;
  return env_dir_from_uv(u, v);
}

void miss_program(const orc_scene& s, const Ray& ray, PerRayData& prd) {
  prd.throughput *= miss_color(s, ray.d);
  prd.scatterEvent = rayMissed;
}

// P/raygen.cu:66-123
f3 color(const orc_scene& s, Ray ray, uint32_t& seed, int maxDepth, int russian, Counters& cn) {
  PerRayData prd;
  prd.seed = seed;
  prd.time = s.d.camera.time0 + rnd(prd.seed) * (s.d.camera.time1 - s.d.camera.time0);
  prd.throughput = mk(1.f);
  prd.radiance = mk(0.f);
  prd.isSpecular = false;
  prd.last_pdf = 0.f;
  bool previousHitSpecular = false;
  for (int depth = 0; depth < maxDepth; depth++) {
    ray.time = prd.time;
    ray.seed = prd.seed;
    Hit h = closest_hit(s, ray, &cn.nodes, &cn.prims);
    cn.closest++;
    cn.bounces++;
    const float pdf_in = prd.last_pdf;
    prd.last_pdf = 0.f;
    if (h.prim >= 0)
      closest_hit_program(s, ray, h, prd, cn);
    else
      miss_program(s, ray, prd);
    if (prd.scatterEvent == rayMissed) {
      if (env_nee_active(s)) {  // NEW estimator: MIS weight against the env pdf, no clamp
        float w = 1.f;
        if (pdf_in > 0.f) w = PowerHeuristic(1, pdf_in, 1, env_pdf_dir(s, ray.d));
        return prd.radiance + prd.throughput * w;
      }
      return prd.radiance + clamp3(prd.throughput, 0.f, 1.f);
    }
    else if (prd.scatterEvent == rayHitLight) {
      if (depth == 0 || previousHitSpecular) prd.radiance += prd.throughput;
      return prd.radiance;
    } else if (prd.scatterEvent == rayGotCancelled)
      return prd.radiance;
    else {
      ray = {prd.origin, prd.direction, 1e-3f, RT_DEFAULT_MAX, prd.time, prd.seed};
      previousHitSpecular = prd.isSpecular;
    }
    if (russian) {
      float prob = max_component(prd.throughput);
      if (depth > 10) {
        if (rnd(prd.seed) >= prob)
          return prd.radiance + prd.throughput;
        else
          prd.throughput *= 1.f / prob;
      }
    }
  }
  return mk(0.f);
}

inline f3 de_nan(f3 c) {  // P/raygen.cu:126-134
  if (!(c.x == c.x)) c.x = 0.f;
  if (!(c.y == c.y)) c.y = 0.f;
  if (!(c.z == c.z)) c.z = 0.f;
  return c;
}

// P/raygen.cu:50-64,148-162
inline Ray camera_ray(const orc_scene& s, int W, int H, int x, int y, int frame, uint32_t& seed) {
  const pt_camera& c = s.d.camera;
  seed = tea<64>((uint32_t)W * (uint32_t)y + (uint32_t)x, (uint32_t)frame);
  float u = float(x + rnd(seed)) / W;
  float v = float(y + rnd(seed)) / H;
  const f3 rd = c.lens_radius * random_in_unit_disk(seed);
  const f3 lens_offset = ld3(c.u) * rd.x + ld3(c.v) * rd.y;
  const f3 origin = ld3(c.origin) + lens_offset;
  const f3 direction =
      ld3(c.lower_left_corner) + u * ld3(c.horizontal) + v * ld3(c.vertical) - origin;
  return {origin, direction, 1e-6f, RT_DEFAULT_MAX, 0.f, 0u};
}

inline Surface make_surface(int mat, const float* color, const float* color2, float p0, float p1) {
  Surface s;
  s.type = mat;
  s.color = ld3(color);
  s.color2 = color2 ? ld3(color2) : mk(0.f);
  s.p0 = p0;
  s.p1 = p1;
  return s;
}

}  // namespace

// ---------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------
extern "C" {

orc_scene* orc_scene_create(const pt_scene_desc* desc, int mode) {
  orc_scene* s = new orc_scene();
  s->d = *desc;
  if (mode == 0) mode = desc->n_prims <= 64 ? 1 : 2;
  s->mode = mode;
  if (mode == 2) build_bvh(*s);
  return s;
}
void orc_scene_destroy(orc_scene* s) { delete s; }

int orc_trace(const orc_scene* s, const pt_ray* rays, int64_t n, pt_hit* hits, orc_stats* stats) {
  uint64_t nodes = 0, prims = 0;
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, prims)
  for (int64_t i = 0; i < n; i++) {
    // ray sets carry no per-ray data: time = the camera's time0, seed = ray index (as pt_trace_rays)
    Ray r = {ld3(rays[i].origin), ld3(rays[i].dir), rays[i].tmin, rays[i].tmax, s->d.camera.time0, (uint32_t)i};
    if (!(r.tmax > r.tmin)) {
      hits[i] = {r.tmax, -1, 0.f, 0.f};
      continue;
    }
    Hit h = closest_hit(*s, r, &nodes, &prims);
    hits[i] = {h.t, h.prim, h.u, h.v};
  }
  if (stats) {
    stats->closest_rays = (uint64_t)n;
    stats->node_tests = nodes;
    stats->prim_tests = prims;
    stats->threads = omp_get_max_threads();
    stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return 0;
}

int orc_camera_rays(const orc_scene* s, int W, int H, int frame, pt_ray* out) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      uint32_t seed;
      Ray r = camera_ray(*s, W, H, x, y, frame, seed);
      pt_ray& o = out[(size_t)y * W + x];
      st3(o.origin, r.o);
      st3(o.dir, r.d);
      o.tmin = r.tmin;
      o.tmax = r.tmax;
    }
  return 0;
}

int orc_bounce_rays(const orc_scene* s, const pt_ray* in, int64_t n, uint32_t seed0, pt_ray* out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < n; i++) {
    Ray r = {ld3(in[i].origin), ld3(in[i].dir), in[i].tmin, in[i].tmax, s->d.camera.time0, (uint32_t)i};
    pt_ray& o = out[i];
    memset(&o, 0, sizeof(o));
    o.tmax = -1.f;
    if (!(r.tmax > r.tmin)) continue;
    Hit h = closest_hit(*s, r, nullptr, nullptr);
    if (h.prim < 0) continue;
    HitRecord rec = get_hit_record(*s, r, h);
    uint32_t seed = tea<16>((uint32_t)i, seed0);
    float u1 = rnd(seed), u2 = rnd(seed);
    f3 N = rec.shading_normal;
    if (dot(N, rec.Wo) < 0.f) N = -N;
    f3 Wi = Sample_cosine(N, u1, u2);
    if (!(Wi.x == Wi.x && Wi.y == Wi.y && Wi.z == Wi.z)) continue;
    st3(o.origin, rec.P);
    st3(o.dir, Wi);
    o.tmin = 1e-3f;
    o.tmax = RT_DEFAULT_MAX;
  }
  return 0;
}

int orc_render(const orc_scene* s, int W, int H, int first_frame, int nframes, int max_depth,
               int russian, float* accum, orc_stats* stats) {
  Counters total;
  auto t0 = std::chrono::steady_clock::now();
  uint64_t closest = 0, shadow = 0, bounces = 0, nodes = 0, prims = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : closest, shadow, bounces, nodes, prims)
  for (int y = 0; y < H; y++) {
    Counters cn;
    for (int f = first_frame; f < first_frame + nframes; f++)
      for (int x = 0; x < W; x++) {
        uint32_t seed;
        Ray r = camera_ray(*s, W, H, x, y, f, seed);
        f3 col = de_nan(color(*s, r, seed, max_depth, russian, cn));
        float* px = accum + 4 * ((size_t)(H - 1 - y) * W + x);  // raygen.cu:153
        px[0] += col.x;
        px[1] += col.y;
        px[2] += col.z;
        px[3] += 1.f;
      }
    closest += cn.closest;
    shadow += cn.shadow;
    bounces += cn.bounces;
    nodes += cn.nodes;
    prims += cn.prims;
  }
  if (stats) {
    stats->samples = (uint64_t)W * H * nframes;
    stats->closest_rays = closest;
    stats->shadow_rays = shadow;
    stats->bounces = bounces;
    stats->node_tests = nodes;
    stats->prim_tests = prims;
    stats->threads = omp_get_max_threads();
    stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return 0;
}

int orc_render_pixel(const orc_scene* s, int W, int H, int x, int y, int frame, int max_depth,
                     int russian, float* rgb) {
  Counters cn;
  uint32_t seed;
  Ray r = camera_ray(*s, W, H, x, y, frame, seed);
  f3 col = de_nan(color(*s, r, seed, max_depth, russian, cn));
  st3(rgb, col);
  return 0;
}

int orc_tonemap_png8(const float* accum, int W, int H, int samples, uint8_t* rgb) {
  for (size_t i = 0; i < (size_t)W * H; i++) {  // H/image_save.hpp:13-29
    f3 col = mk(accum[4 * i], accum[4 * i + 1], accum[4 * i + 2]);
    col = col / float(samples);
    col = mk(sqrtf(col.x), sqrtf(col.y), sqrtf(col.z));
    col = 255.99f * clamp3(col, 0.f, 1.f);
    rgb[3 * i + 0] = (uint8_t)(int)col.x;
    rgb[3 * i + 1] = (uint8_t)(int)col.y;
    rgb[3 * i + 2] = (uint8_t)(int)col.z;
  }
  return 0;
}

uint32_t orc_tea64(uint32_t v0, uint32_t v1) { return tea<64>(v0, v1); }
float orc_rnd(uint32_t* seed) { return rnd(*seed); }
// P/raygen.cu:136-146 make_Color with (frame + 1) = the pixel's sample count
int orc_display_rgba8(const float* accum, int W, int H, uint8_t* rgba) {
  for (size_t i = 0; i < (size_t)W * H; i++) {
    const float* c = accum + 4 * i;
    f3 temp = mk(c[0], c[1], c[2]) / (c[3] > 0.f ? c[3] : 1.f);
    temp = clamp3(mk(sqrtf(temp.x), sqrtf(temp.y), sqrtf(temp.z)), 0.f, 1.f);
    rgba[4 * i + 0] = (uint8_t)int(255.99 * temp.x);
    rgba[4 * i + 1] = (uint8_t)int(255.99 * temp.y);
    rgba[4 * i + 2] = (uint8_t)int(255.99 * temp.z);
    rgba[4 * i + 3] = (uint8_t)int(255.99 * 1.f);
  }
  return 0;
}

void orc_random_in_unit_disk(uint32_t* seed, float* o) { st3(o, random_in_unit_disk(*seed)); }
void orc_random_in_unit_sphere(uint32_t* seed, float* o) { st3(o, random_in_unit_sphere(*seed)); }
void orc_random_on_unit_sphere(uint32_t* seed, float* o) { st3(o, random_on_unit_sphere(*seed)); }

void orc_bsdf_sample(int mat, const float* color, const float* color2, float p0, float p1,
                     const float* P, const float* Wo, const float* N, float u1, float u2,
                     float* Wi_out) {
  Surface s = make_surface(mat, color, color2, p0, p1);
  (void)P;
  f3 wi;
  switch (mat) {
    case PT_MAT_TORRANCE_SPARROW:
      wi = Sample_TS(s, ld3(Wo), ld3(N), u1, u2);
      break;
    case PT_MAT_ASHIKHMIN_SHIRLEY:
      wi = Sample_AS(s, ld3(Wo), ld3(N), u1, u2);
      break;
    default:
      wi = Sample_cosine(ld3(N), u1, u2);
      break;
  }
  st3(Wi_out, wi);
}
void orc_bsdf_eval(int mat, const float* color, const float* color2, float p0, float p1,
                   const float* P, const float* Wo, const float* Wi, const float* N, float* f_out,
                   float* pdf_out) {
  Surface s = make_surface(mat, color, color2, p0, p1);
  float pdf = 0.f;
  f3 f = Evaluate(s, ld3(P), ld3(Wo), ld3(Wi), ld3(N), pdf);
  st3(f_out, f);
  *pdf_out = pdf;
}
float orc_schlick(float cosine, float ref_idx) { return schlick(cosine, ref_idx); }
int orc_refract(const float* v, const float* n, float ni_over_nt, float* out3) {
  f3 r = mk(0.f);
  bool ok = Refract(ld3(v), ld3(n), ni_over_nt, r);
  st3(out3, r);
  return ok ? 1 : 0;
}
void orc_ggx_sample(const float* origin, float u1, float u2, float nu, float nv, float* out3) {
  st3(out3, GGX_Sample(ld3(origin), u1, u2, nu, nv));
}
float orc_ggx_d(const float* H, float nu, float nv) { return GGX_D(ld3(H), nu, nv); }
float orc_ggx_g(const float* Wo, const float* Wi, float nu, float nv) {
  return GGX_G(ld3(Wo), ld3(Wi), nu, nv);
}
void orc_rect_sample(const pt_light* l, const float* P, float u1, float u2, float* Wi_out) {
  st3(Wi_out, Rect_Sample(*l, ld3(P), u1, u2));
}
float orc_rect_pdf(const pt_light* l, const float* P, const float* Wi) {
  return Rect_PDF(*l, ld3(P), ld3(Wi));
}
int orc_intersect_prim(const orc_scene* s, int prim, const pt_ray* ray, float* t, float* u,
                       float* v) {
  Ray r = {ld3(ray->origin), ld3(ray->dir), ray->tmin, ray->tmax};
  return intersect_prim(*s, prim, r, *t, *u, *v) ? 1 : 0;
}
// the same with the per-ray data moving spheres (time) and volumes (seed) read, plus the hit record
// (P, geometric normal, shading normal, u, v, Wo) and the UNPADDED-by-the-BVH reference bounds check
int orc_intersect_prim_ex(const orc_scene* s, int prim, const pt_ray* ray, float time, uint32_t seed,
                          float* hit3, float* rec14) {
  Ray r = {ld3(ray->origin), ld3(ray->dir), ray->tmin, ray->tmax, time, seed};
  float t, u, v;
  if (!intersect_prim(*s, prim, r, t, u, v)) return 0;
  hit3[0] = t;
  hit3[1] = u;
  hit3[2] = v;
  Hit h = {t, prim, u, v};
  HitRecord rec = get_hit_record(*s, r, h);
  st3(rec14, rec.P);
  st3(rec14 + 3, rec.geometric_normal);
  st3(rec14 + 6, rec.shading_normal);
  rec14[9] = rec.u;
  rec14[10] = rec.v;
  st3(rec14 + 11, rec.Wo);
  return 1;
}
// One closest-hit program invocation (P/materials/*.cu closest_hit) on the hit (prim, t, u, v) of `ray`.
// prd_io: seed (bit pattern), time, throughput.xyz, radiance.xyz on entry; on return seed, time, throughput,
// radiance, origin.xyz, direction.xyz, scatterEvent, isSpecular (16 floats).  shadow4: number of shadow rays
// traced, then the last one's origin.xyz direction.xyz and whether it was blocked (8 floats).
void orc_closest_hit_probe(const orc_scene* s, const pt_ray* ray, int prim, float t, float u, float v,
                           float* prd_io, float* shadow8) {
  PerRayData prd;
  memcpy(&prd.seed, prd_io, 4);
  prd.time = prd_io[1];
  prd.throughput = ld3(prd_io + 2);
  prd.radiance = ld3(prd_io + 5);
  prd.origin = prd.direction = mk(0.f);
  prd.scatterEvent = rayGotBounced;
  prd.isSpecular = false;
  prd.last_pdf = 0.f;
  Ray r = {ld3(ray->origin), ld3(ray->dir), ray->tmin, ray->tmax, prd.time, prd.seed};
  Hit h = {t, prim, u, v};
  Counters cn;
  closest_hit_program(*s, r, h, prd, cn);
  memcpy(prd_io, &prd.seed, 4);
  prd_io[1] = prd.time;
  st3(prd_io + 2, prd.throughput);
  st3(prd_io + 5, prd.radiance);
  st3(prd_io + 8, prd.origin);
  st3(prd_io + 11, prd.direction);
  prd_io[14] = (float)(int)prd.scatterEvent;
  prd_io[15] = prd.isSpecular ? 1.f : 0.f;
  shadow8[0] = (float)cn.shadow_traced;
  memcpy(shadow8 + 1, cn.shadow_ray, 24);
  shadow8[7] = (float)cn.shadow_blocked;
}
// One rtTrace of a radiance ray as OptiX performs it: closest hit -> that material's closest-hit program, or
// the miss program (P/raygen.cu:77).  prd16 as in orc_closest_hit_probe.  The reference's own color() loop
// (compiled from raygen.cu, oracle/ref_ray_probe.cpp) is driven through this entry point.
void orc_trace_radiance(const orc_scene* s, const float* ray8, float* prd16) {
  PerRayData prd;
  memcpy(&prd.seed, prd16, 4);
  prd.time = prd16[1];
  prd.throughput = ld3(prd16 + 2);
  prd.radiance = ld3(prd16 + 5);
  prd.origin = ld3(prd16 + 8);
  prd.direction = ld3(prd16 + 11);
  prd.scatterEvent = (ScatterEvent)(int)prd16[14];
  prd.isSpecular = prd16[15] != 0.f;
  prd.last_pdf = 0.f;
  Ray r = {ld3(ray8), ld3(ray8 + 4), ray8[3], ray8[7], prd.time, prd.seed};
  Counters cn;
  Hit h = closest_hit(*s, r, &cn.nodes, &cn.prims);
  if (h.prim >= 0) closest_hit_program(*s, r, h, prd, cn);
  else miss_program(*s, r, prd);
  memcpy(prd16, &prd.seed, 4);
  prd16[1] = prd.time;
  st3(prd16 + 2, prd.throughput);
  st3(prd16 + 5, prd.radiance);
  st3(prd16 + 8, prd.origin);
  st3(prd16 + 11, prd.direction);
  prd16[14] = (float)(int)prd.scatterEvent;
  prd16[15] = prd.isSpecular ? 1.f : 0.f;
}
// miss program on a throughput; also the (u, v) an image background samples at
void orc_miss_probe(const orc_scene* s, const float* dir, const float* thr_in, float* thr_out, float* uv2) {
  PerRayData prd;
  prd.throughput = ld3(thr_in);
  Ray r = {mk(0.f), ld3(dir), 1e-3f, RT_DEFAULT_MAX};
  miss_program(*s, r, prd);
  st3(thr_out, prd.throughput);
  uv2[0] = uv2[1] = 0.f;
  if (s->d.miss.type == PT_MISS_IMG || s->d.miss.type == PT_MISS_HDR) miss_uv(s->d.miss, ld3(dir), uv2[0], uv2[1]);
}
// texture program `tex` of the scene at (u, v, P, index) — P/textures/*.cu sample_texture
void orc_sample_texture(const orc_scene* s, int tex, float u, float v, const float* P, int index, float* rgb) {
  st3(rgb, sample_texture(*s, tex, u, v, ld3(P), index));
}
void orc_prim_bounds(const orc_scene* s, int prim, float* lo_hi6) {
  Box3 b = prim_bounds(*s, prim);
  st3(lo_hi6, b.lo);
  st3(lo_hi6 + 3, b.hi);
}
// light sampling / pdf for either light kind; the sample consumes its draws from *seed
void orc_light_sample(const pt_light* l, const float* P, uint32_t* seed, float* Wi_out) {
  st3(Wi_out, Light_Sample(*l, ld3(P), *seed));
}
float orc_light_pdf(const pt_light* l, const float* P, const float* Wi) { return Light_PDF(*l, ld3(P), ld3(Wi)); }
uint32_t orc_volume_seed(uint32_t ray_seed, uint32_t prim) { return tea<4>(ray_seed, prim); }

}  // extern "C"
