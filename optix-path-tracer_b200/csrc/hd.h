// hd.h — host/device portability layer for the kernel *cores*.
//
// Every per-thread kernel body in this library is written as a PT_HD function
// so that tests/hostsim can compile the very same code with g++ and run it in
// a serial loop (logic tests without a GPU).  The shipped library always runs
// these bodies inside __global__ kernels; nothing here is a CPU fallback.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define PT_HD __host__ __device__ __forceinline__
#define PT_D __device__ __forceinline__
#else
#define PT_HD inline
#define PT_D inline
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
#endif

// ---- bit casts / intrinsics with host equivalents ---------------------------
PT_HD float pt_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
PT_HD uint32_t pt_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PT_HD int pt_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
PT_HD int pt_clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}
PT_HD int pt_clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
// index of highest set bit (v != 0)
PT_HD int pt_bfind(uint32_t v) { return 31 - pt_clz32(v); }

// Exactly-rounded, never-contracted binary32 ops.  Used for everything whose
// result decides a hit id (primitive intersection, object-space ray
// transforms): the CPU oracle evaluates the same expressions with
// -ffp-contract=off, so results match bit for bit.
#ifdef __CUDA_ARCH__
#define XMUL(a, b) __fmul_rn((a), (b))
#define XADD(a, b) __fadd_rn((a), (b))
#define XSUB(a, b) __fsub_rn((a), (b))
#define XDIV(a, b) __fdiv_rn((a), (b))
#define XSQRT(a) __fsqrt_rn((a))
#define XFMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define XMUL(a, b) ((a) * (b))
#define XADD(a, b) ((a) + (b))
#define XSUB(a, b) ((a) - (b))
#define XDIV(a, b) ((a) / (b))
#define XSQRT(a) sqrtf((a))
#define XFMA(a, b, c) fmaf((a), (b), (c))
#endif

// ---- float3 helpers (OptiX SDK semantics: v/s = v*(1/s), normalize = v*rsqrt) -
PT_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
PT_HD float3 f3(float s) { return make_float3(s, s, s); }
PT_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
PT_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
PT_HD float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
PT_HD float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
PT_HD float3 operator/(float3 a, float s) { float inv = 1.0f / s; return a * inv; }
PT_HD float3 operator+(float3 a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
PT_HD float3 operator-(float3 a, float s) { return f3(a.x - s, a.y - s, a.z - s); }
PT_HD float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
PT_HD float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
PT_HD float3& operator*=(float3& a, float s) { a = a * s; return a; }
PT_HD float3& operator/=(float3& a, float s) { a = a / s; return a; }
PT_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD float3 cross(float3 a, float3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PT_HD float length(float3 v) { return sqrtf(dot(v, v)); }
PT_HD float3 normalize(float3 v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
PT_HD float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
PT_HD float3 clamp3(float3 v, float a, float b) {
  return f3(clampf(v.x, a, b), clampf(v.y, a, b), clampf(v.z, a, b));
}
PT_HD float3 reflect3(float3 i, float3 n) { return i - 2.0f * n * dot(n, i); }
PT_HD float3 fmin3(float3 a, float3 b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
PT_HD float3 fmax3(float3 a, float3 b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
PT_HD float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }
PT_HD float comp(float3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// exact (hit-deciding) vector ops
PT_HD float xdot(float3 a, float3 b) {
  return XADD(XADD(XMUL(a.x, b.x), XMUL(a.y, b.y)), XMUL(a.z, b.z));
}
PT_HD float3 xcross(float3 a, float3 b) {
  return f3(XSUB(XMUL(a.y, b.z), XMUL(a.z, b.y)), XSUB(XMUL(a.z, b.x), XMUL(a.x, b.z)),
            XSUB(XMUL(a.x, b.y), XMUL(a.y, b.x)));
}
PT_HD float3 xsub(float3 a, float3 b) { return f3(XSUB(a.x, b.x), XSUB(a.y, b.y), XSUB(a.z, b.z)); }
// row-major 3x4 times point / vector (programs: rtTransformPoint / rtTransformVector)
PT_HD float3 xf_point(const float* m, float3 p) {
  return f3(XADD(XADD(XADD(XMUL(m[0], p.x), XMUL(m[1], p.y)), XMUL(m[2], p.z)), m[3]),
            XADD(XADD(XADD(XMUL(m[4], p.x), XMUL(m[5], p.y)), XMUL(m[6], p.z)), m[7]),
            XADD(XADD(XADD(XMUL(m[8], p.x), XMUL(m[9], p.y)), XMUL(m[10], p.z)), m[11]));
}
PT_HD float3 xf_vector(const float* m, float3 v) {
  return f3(XADD(XADD(XMUL(m[0], v.x), XMUL(m[1], v.y)), XMUL(m[2], v.z)),
            XADD(XADD(XMUL(m[4], v.x), XMUL(m[5], v.y)), XMUL(m[6], v.z)),
            XADD(XADD(XMUL(m[8], v.x), XMUL(m[9], v.y)), XMUL(m[10], v.z)));
}
// rtTransformNormal(RT_OBJECT_TO_WORLD, n) = transpose(world->object 3x3) * n
PT_HD float3 xf_normal(const float* w2o, float3 n) {
  return f3(w2o[0] * n.x + w2o[4] * n.y + w2o[8] * n.z, w2o[1] * n.x + w2o[5] * n.y + w2o[9] * n.z,
            w2o[2] * n.x + w2o[6] * n.y + w2o[10] * n.z);
}

#define PT_PI_F 3.141592654f                 /* programs/common.hpp:18 PI_F  */
#define PT_OPTIX_PI 3.14159265358979323846f  /* optixu M_PIf                  */
#define PT_OPTIX_1_PI 0.318309886183790671538f
#define PT_RAY_MAX 1.e27f                    /* RT_DEFAULT_MAX                */
