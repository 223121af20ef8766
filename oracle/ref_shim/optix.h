// optix.h — minimal host shim so the reference's header-only device code
// (programs/*.cuh, *.hpp) compiles with g++.  TEST INFRASTRUCTURE ONLY.
// Restates the public OptiX SDK 6.0.0 optixu_math_namespace.h helpers the
// reference uses; nothing here is copied from the reference.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <cmath>
#include <cstdio>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

using std::abs;
using std::isinf;

typedef unsigned int uint;
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };

inline float2 make_float2(float x, float y) { return {x, y}; }
inline float2 make_float2(float s) { return {s, s}; }
inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float3 make_float3(float s) { return {s, s, s}; }
inline float3 make_float3(const float4& v) { return {v.x, v.y, v.z}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline int2 make_int2(int x, int y) { return {x, y}; }

#define RT_DEFAULT_MAX 1.e27f
#ifndef M_PIf
#define M_PIf 3.14159265358979323846f
#endif
#ifndef M_1_PIf
#define M_1_PIf 0.318309886183790671538f
#endif

namespace optix {

inline float3 operator+(const float3& a, const float3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator+(const float3& a, float b) { return {a.x + b, a.y + b, a.z + b}; }
inline float3 operator-(const float3& a, const float3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(const float3& a, float b) { return {a.x - b, a.y - b, a.z - b}; }
inline float3 operator-(const float3& a) { return {-a.x, -a.y, -a.z}; }
inline float3 operator*(const float3& a, const float3& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator*(const float3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, const float3& a) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator/(const float3& a, const float3& b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }
inline float3 operator/(float s, const float3& a) { return {s / a.x, s / a.y, s / a.z}; }
inline void operator+=(float3& a, const float3& b) { a = a + b; }
inline void operator-=(float3& a, const float3& b) { a = a - b; }
inline void operator*=(float3& a, const float3& b) { a = a * b; }
inline void operator*=(float3& a, float s) { a = a * s; }
inline void operator/=(float3& a, float s) { float inv = 1.0f / s; a = a * inv; }
inline float2 operator*(const float2& a, float s) { return {a.x * s, a.y * s}; }

inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(const float3& a, const float3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline float3 normalize(const float3& v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
inline float clamp(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
inline float3 clamp(const float3& v, float a, float b) {
  return {clamp(v.x, a, b), clamp(v.y, a, b), clamp(v.z, a, b)};
}
inline float3 reflect(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }

inline void cosine_sample_hemisphere(const float u1, const float u2, float3& p) {
  const float r = sqrtf(u1);
  const float phi = 2.0f * M_PIf * u2;
  p.x = r * cosf(phi);
  p.y = r * sinf(phi);
  p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
}

struct Onb {
  inline Onb(const float3& normal) {
    m_normal = normal;
    if (fabs(m_normal.x) > fabs(m_normal.z)) {
      m_binormal.x = -m_normal.y; m_binormal.y = m_normal.x; m_binormal.z = 0;
    } else {
      m_binormal.x = 0; m_binormal.y = -m_normal.z; m_binormal.z = m_normal.y;
    }
    m_binormal = normalize(m_binormal);
    m_tangent = cross(m_binormal, m_normal);
  }
  inline void inverse_transform(float3& p) const {
    p = p.x * m_tangent + p.y * m_binormal + p.z * m_normal;
  }
  float3 m_tangent, m_binormal, m_normal;
};

struct Ray {
  float3 origin, direction;
  unsigned ray_type;
  float tmin, tmax;
};
inline Ray make_Ray(float3 o, float3 d, unsigned type, float tmin, float tmax) {
  return {o, d, type, tmin, tmax};
}

}  // namespace optix

inline float saturate(float x) { return fmaxf(0.f, fminf(1.f, x)); }

template <typename T> struct rtCallableProgramId {};
template <typename T> struct rtCallableProgramX {};
