// traverse.h — primitive intersection + compressed-BVH8 traversal core.
//
// Replaces OptiX's rtTrace / intersection-program dispatch
// (programs/raygen.cu:77, programs/hitables/{sphere,aarect,box,triangle}.cu).
// Semantics (pinned with the oracle): closest hit with t in the OPEN interval
// (tmin, tmax); equal-t hits resolve to the lowest primitive id.
//
// All hit-deciding arithmetic uses the X* exact ops from hd.h so hit ids are
// bit-identical to the CPU oracle.  The node (box) tests are allowed to use
// FMA: boxes are padded (bvh_build) so the test only has to be conservative.
#pragma once
#include "bvh_core.h"
#include "scene_dev.h"

// Exact divisions of the triangle test deferred to the retire batch (isect_triangle_deferred below): C4 +1.7 %,
// incoherent rays +1.3 %, 10 M triangles +1.5 %, results bit-identical.  -DPT_NO_DEFER_DIV builds the immediate path.
#if !defined(PT_NO_DEFER_DIV) && !defined(PT_DEFER_DIV)
#define PT_DEFER_DIV
#endif

struct TraceHit {
  float t;
  int leaf;  // leaf record index, -1 = miss
  int prim;  // primitive id
  float u, v;
#ifdef PT_DEFER_DIV
  // den != 0: the hit is a triangle whose three exact divisions are still to be done (hit_resolve): u, v hold the
  // numerators, tn the numerator of t, den the determinant, and t an UPPER BOUND of the exact t (good for culling)
  float den, tn;
#endif
};

// ---- per-primitive tests: candidate t in (tmin,tmax) of the ray -----------------
// programs/hitables/triangle.cu:26-58.  `tbest` (<= tmax) is the closest hit so far: a candidate
// beyond it can never win.  The "certainly out" pre-tests decide from numerators and |A| alone
// when the quotient is clearly outside its range (margins of 2^-18 >> the 2^-24 rounding of an
// IEEE quotient); everything near a boundary falls through to the exact divisions, so the
// accept/reject decision and the returned t,u,v are bit-identical to the reference sequence.
PT_HD bool isect_triangle(float3 a, float3 b, float3 c, float3 o, float3 d, float tmin, float tmax, float tbest,
                          float& t, float& bu, float& bv) {
  float3 e1 = xsub(b, a);
  float3 e2 = xsub(c, a);
  float3 P = xcross(d, e2);
  float A = xdot(P, e1);
  const float absA = fabsf(A);
  if ((double)absA < 1E-8) return false;
  float3 R = xsub(o, a);
  const float uN = xdot(P, R);
  const float hiA = absA * 1.000004f;   // > |A| (1 + 2^-18)
  const float tiny = absA * 1e-30f;     // quotients above this cannot underflow to zero
  const bool opp_u = (uN < 0.f) != (A < 0.f);
  if (fabsf(uN) > hiA) return false;                 // |u| > 1: u < 0 or u > 1
  if (opp_u && fabsf(uN) > tiny) return false;       // u < 0
  float3 Q = xcross(R, e1);
  const float vN = xdot(Q, d);
  const bool opp_v = (vN < 0.f) != (A < 0.f);
  if (fabsf(vN) > hiA) return false;                 // v < 0 or v > 1 >= 1 - u
  if (opp_v && fabsf(vN) > tiny) return false;       // v < 0
  const float tN = xdot(Q, e2);
  if (tmin >= 0.f) {
    const bool opp_t = (tN < 0.f) != (A < 0.f);
    if (opp_t && fabsf(tN) > tiny) return false;     // t < 0 <= tmin
    if (tbest < 1e30f && fabsf(tN) > hiA * tbest) return false;  // t > tbest
  }
  float u = XDIV(uN, A);
  if (u < 0.0f || u > 1.0f) return false;
  float v = XDIV(vN, A);
  if (v < 0.0f || XADD(u, v) > 1.0f) return false;
  float tt = XDIV(tN, A);
  if (tt > tmin && tt < tmax) {
    t = tt;
    bu = u;
    bv = v;
    return true;
  }
  return false;
}
// programs/hitables/sphere.cu:32-62
PT_HD bool isect_sphere(float3 center, float radius, float3 o, float3 d, float tmin, float tmax,
                        float& t) {
  float3 oc = xsub(o, center);
  float a = xdot(d, d);
  float b = xdot(oc, d);
  float c = XSUB(xdot(oc, oc), XMUL(radius, radius));
  float disc = XSUB(XMUL(b, b), XMUL(a, c));
  if (disc < 0.f) return false;
  float sq = XSQRT(disc);
  float t0 = XDIV(XSUB(-b, sq), a);
  if (t0 > tmin && t0 < tmax) {
    t = t0;
    return true;
  }
  float t1 = XDIV(XADD(-b, sq), a);
  if (t1 > tmin && t1 < tmax) {
    t = t1;
    return true;
  }
  return false;
}
// programs/hitables/aarect.cu:20-69 ; params: a0,a1,b0,b1,k ; axis
PT_HD bool isect_rect(float a0, float a1, float b0, float b1, float k, int axis, float3 o, float3 d,
                      float tmin, float tmax, float& t) {
  float oa, ob, ok, da, db, dk;
  if (axis == PT_X_AXIS) {
    ok = o.x; dk = d.x; oa = o.y; da = d.y; ob = o.z; db = d.z;
  } else if (axis == PT_Y_AXIS) {
    ok = o.y; dk = d.y; oa = o.x; da = d.x; ob = o.z; db = d.z;
  } else {
    ok = o.z; dk = d.z; oa = o.x; da = d.x; ob = o.y; db = d.y;
  }
  float tt = XDIV(XSUB(k, ok), dk);
  float a = XADD(oa, XMUL(tt, da));
  float b = XADD(ob, XMUL(tt, db));
  if (a < a0 || a > a1 || b < b0 || b > b1) return false;
  if (tt > tmin && tt < tmax) {
    t = tt;
    return true;
  }
  return false;
}
// programs/hitables/box.cu:36-53 (ffmin/ffmax semantics of programs/vec.hpp:57-84)
PT_HD float pt_ffmin(float a, float b) { return a < b ? a : b; }
PT_HD float pt_ffmax(float a, float b) { return a > b ? a : b; }
PT_HD bool isect_box(float3 bmin, float3 bmax, float3 o, float3 d, float tmin, float tmax, float& t) {
  float3 t0 = f3(XDIV(XSUB(bmin.x, o.x), d.x), XDIV(XSUB(bmin.y, o.y), d.y), XDIV(XSUB(bmin.z, o.z), d.z));
  float3 t1 = f3(XDIV(XSUB(bmax.x, o.x), d.x), XDIV(XSUB(bmax.y, o.y), d.y), XDIV(XSUB(bmax.z, o.z), d.z));
  float tn = pt_ffmax(pt_ffmax(pt_ffmin(t0.x, t1.x), pt_ffmin(t0.y, t1.y)), pt_ffmin(t0.z, t1.z));
  float tf = pt_ffmin(pt_ffmin(pt_ffmax(t0.x, t1.x), pt_ffmax(t0.y, t1.y)), pt_ffmax(t0.z, t1.z));
  if (tn <= tf) {
    if (tn > tmin && tn < tmax) {
      t = tn;
      return true;
    } else if (tf > tmin && tf < tmax) {
      t = tf;
      return true;
    }
  }
  return false;
}

// ---- time-dependent and volumetric primitives (rare: kept out of line on the device so the
// hot triangle/sphere path of the traversal kernels keeps its register budget) ----------------
// Per-ray data these primitives read: the sample's shutter time (prd.time, raygen.cu:69) and an
// RNG seed.  PINNED: the reference draws `rnd(prd.seed)` inside the volume intersection programs
// (volume_sphere.cu:80, volume_box.cu:55), i.e. once per intersection-program invocation in
// OptiX's traversal order with OptiX's culling - unknowable and order-dependent.  Here the uniform
// of (ray, volume primitive) is a pure function of the ray's seed and the primitive id:
// rnd(tea<4>(seed, prim)).  Same distribution, independent of traversal order, and the path's
// own stream is not advanced by traversal.
struct RayAux {
  float time;
  uint32_t seed;
};
// Aux providers: `enabled == false` compiles the special-primitive branch out of the leaf test
// (scenes without such primitives run kernels whose hot path is exactly the plain one).
struct NoAux {
  static constexpr bool enabled = false;
  PT_HD RayAux operator()() const { return RayAux{0.f, 0u}; }
};
PT_HD float pt_volume_uniform(uint32_t seed, uint32_t prim) {
  uint32_t v0 = seed, v1 = prim, s0 = 0;  // programs/random.cuh:33-47 tea<4>
  for (int n = 0; n < 4; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  v0 = 1664525u * v0 + 1013904223u;  // random.cuh:50-65 lcg + rnd
  return (float)(v0 & 0x00FFFFFFu) / (float)0x01000000;
}
// programs/hitables/volume_sphere.cu:20-54 hit_boundary
PT_HD bool vol_boundary_sphere(float3 center, float radius, float3 o, float3 d, float tmin, float tmax, float& rec) {
  return isect_sphere(center, radius, o, d, tmin, tmax, rec);
}
// programs/hitables/volume_box.cu:20-42 hit_boundary
PT_HD bool vol_boundary_box(float3 bmin, float3 bmax, float3 o, float3 d, float tmin, float tmax, float& rec) {
  float3 t0 = f3(XDIV(XSUB(bmin.x, o.x), d.x), XDIV(XSUB(bmin.y, o.y), d.y), XDIV(XSUB(bmin.z, o.z), d.z));
  float3 t1 = f3(XDIV(XSUB(bmax.x, o.x), d.x), XDIV(XSUB(bmax.y, o.y), d.y), XDIV(XSUB(bmax.z, o.z), d.z));
  float temp1 = pt_ffmax(pt_ffmax(pt_ffmin(t0.x, t1.x), pt_ffmin(t0.y, t1.y)), pt_ffmin(t0.z, t1.z));
  float temp2 = pt_ffmin(pt_ffmin(pt_ffmax(t0.x, t1.x), pt_ffmax(t0.y, t1.y)), pt_ffmax(t0.z, t1.z));
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
// volume_sphere.cu:62-105 / volume_box.cu:44-82: scattering distance inside a constant-density
// medium.  As written in the reference the exit distance only bounds the entry (rec1 < rec2); the
// sampled distance itself is accepted anywhere in the ray interval.
#ifdef __CUDACC__
static __host__ __device__ __noinline__
#else
inline
#endif
bool isect_special(int kind, float4 La, float4 Lb, float4 Lc, float3 o, float3 d, float tmin, float tmax,
                   float time, uint32_t seed, float& t) {
  if (kind == PT_PRIM_MOVING_SPHERE) {  // moving_sphere.cu:54-131
    const float c0[3] = {La.x, La.y, La.z}, c1[3] = {Lc.x, Lc.y, Lc.z};
    return isect_sphere(moving_center(c0, c1, Lb.y, Lb.z, time), Lb.x, o, d, tmin, tmax, t);
  }
  const float FMAX = 3.402823466e+38f;
  const bool box = kind == PT_PRIM_VOLUME_BOX;
  const float density = box ? Lc.x : Lb.y;
  float rec1, rec2;
  if (!(box ? vol_boundary_box(xyz(La), xyz(Lb), o, d, -FMAX, FMAX, rec1)
            : vol_boundary_sphere(xyz(La), Lb.x, o, d, -FMAX, FMAX, rec1)))
    return false;
  const float after = (float)((double)rec1 + 0.0001);  // `rec1 + 0.0001` is a double sum in the reference
  if (!(box ? vol_boundary_box(xyz(La), xyz(Lb), o, d, after, FMAX, rec2)
            : vol_boundary_sphere(xyz(La), Lb.x, o, d, after, FMAX, rec2)))
    return false;
  if (rec1 < tmin) rec1 = tmin;
  if (rec2 > tmax) rec2 = tmax;
  if (rec1 >= rec2) return false;
  if (rec1 < 0.f) rec1 = 0.f;
  const float len = XSQRT(xdot(d, d));
  const float hit_distance = XMUL(-XDIV(1.f, density), logf(pt_volume_uniform(seed, pt_f2u(La.w))));
  const float temp = XADD(rec1, XDIV(hit_distance, len));
  if (temp > tmin && temp < tmax) {  // rtPotentialIntersection
    t = temp;
    return true;
  }
  return false;
}

// one leaf record against the (world) ray
template <class Aux>
PT_HD bool isect_leaf(const SceneDev& sc, const LeafRec& L, float3 o, float3 d, float tmin, float tmax, float tbest,
                      float& t, float& u, float& v, const Aux& aux) {
  uint32_t w = pt_f2u(L.b.w);
  int kind = PT_LEAF_KIND(w);
  u = 0.f;
  v = 0.f;
  if (kind == PT_PRIM_TRIANGLE) return isect_triangle(xyz(L.a), xyz(L.b), xyz(L.c), o, d, tmin, tmax, tbest, t, u, v);
  int xf = PT_LEAF_XFORM(w);
  if (xf >= 0) {  // Transform node: ray to object space, t is shared
    const float* m = sc.xforms[xf].w2o;
    o = xf_point(m, o);
    d = xf_vector(m, d);
  }
  if (Aux::enabled && kind >= PT_PRIM_MOVING_SPHERE) {
    const RayAux a = aux();
    return isect_special(kind, L.a, L.b, L.c, o, d, tmin, tmax, a.time, a.seed, t);
  }
  if (kind == PT_PRIM_SPHERE) return isect_sphere(xyz(L.a), L.b.x, o, d, tmin, tmax, t);
  if (kind == PT_PRIM_RECT) {
    int ax = (int)pt_f2u(L.b.z);
    return isect_rect(L.a.x, L.a.y, L.a.z, L.b.x, L.b.y, ax & 3, o, d, tmin, tmax, t);
  }
  return isect_box(xyz(L.a), xyz(L.b), o, d, tmin, tmax, t);
}

// ---- BVH8 traversal ------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define PT_LDG4(p) __ldg(p)
#define PT_LDGF4(p) __ldg(p)
#else
#define PT_LDG4(p) (*(p))
#define PT_LDGF4(p) (*(p))
#endif

PT_HD uint32_t pt_sign_extend_s8x4(uint32_t v) {
  // each byte with bit7 set -> 0xff, else 0x00
  return ((v >> 7) & 0x01010101u) * 0xffu;
}
PT_HD uint32_t pt_byte(uint32_t v, int i) { return (v >> (8 * i)) & 0xffu; }

// Quantised plane byte -> float.  I2F runs at 1/8 rate and dominated the node test in the first
// profile, so on the device one PRMT builds the float (2^15 + q) directly (q lands in mantissa
// bits 8..15) and the 2^15 bias is folded into the FMA's addend once per node and axis:
//   t = (2^15 + q) * s + (b - 2^15 * s).
// Folding rounds the addend at magnitude 2^15*s, an error of at most 2^-10 quantisation steps;
// the builder quantises child boxes with a 2^-8-step guard band (bvh_core.h), so the test stays
// conservative.  The host build converts exactly.
// SASS PRMT takes ONE immediate.  With the 0x47000000 operand a compile-time constant ptxas keeps it
// as the immediate and puts the four selectors in registers, which under the kernels' register
// cap are re-materialised before most of the 48 PRMTs of a node test (the first ncu source view:
// 62 of 320 node-step instructions were such moves).  Passing the constant in as a run-time value
// (SceneDev::qbias) makes it the register operand and the selector the immediate.
#ifdef __CUDA_ARCH__
#define PT_QBIAS 32768.0f
PT_HD float pt_byte_to_float(uint32_t w, int j, uint32_t qbias) {
  return __uint_as_float(__byte_perm(w, qbias, 0x7404u | ((uint32_t)j << 4)));
}
#else
#define PT_QBIAS 0.0f
PT_HD float pt_byte_to_float(uint32_t w, int j, uint32_t) { return (float)pt_byte(w, j); }
#endif

struct RayCtx {
  float3 o, d;
  float idx, idy, idz;  // guarded reciprocal direction
  float tmin, tmax;     // the ray's own interval (candidate selection)
  uint32_t octinv;
};

PT_HD void ray_setup(RayCtx& r, float3 o, float3 d, float tmin, float tmax) {
  const float ooeps = 8.2718061e-25f;  // 2^-80
  r.o = o;
  r.d = d;
  r.tmin = tmin;
  r.tmax = tmax;
#if defined(__CUDA_ARCH__) && !defined(PT_EXACT_RCP)
  // the reciprocal only feeds the (conservative, padded) box test: 1 ulp of MUFU.RCP is far inside the padding;
  // three IEEE divisions per fetched ray were ~25 instructions of every refill (C4 +0.4 %, incoherent rays +1.4 %)
  r.idx = __fdividef(1.0f, fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
  r.idy = __fdividef(1.0f, fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
  r.idz = __fdividef(1.0f, fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
#else
  r.idx = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
  r.idy = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
  r.idz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
#endif
  const uint32_t oct = (d.x < 0.f ? 4u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 1u : 0u);
  r.octinv = 7u - oct;
}

PT_HD void hit_reset(TraceHit& hit, float tmax) {
  hit.t = tmax;
  hit.leaf = -1;
  hit.prim = -1;
  hit.u = hit.v = 0.f;
#ifdef PT_DEFER_DIV
  hit.den = 0.f;
  hit.tn = 0.f;
#endif
}
// the deferred divisions of a triangle hit (triangle.cu:41-49: u, v, t = numerator / A), exactly as the immediate path
PT_HD void hit_resolve(TraceHit& hit) {
#ifdef PT_DEFER_DIV
  if (hit.den != 0.f) {
    hit.u = XDIV(hit.u, hit.den);
    hit.v = XDIV(hit.v, hit.den);
    hit.t = XDIV(hit.tn, hit.den);
    hit.den = 0.f;
  }
#endif
}

// Hit-mask assembly of the node step: 8-bit child-slot mask -> octant-ordered priority mask, bit s -> bit s ^ octinv,
// as a 2 KB table (L1-resident on the device).
struct Perm8Table {
  uint8_t v[8][256];
  constexpr Perm8Table() : v{} {
    for (int o = 0; o < 8; o++)
      for (int m = 0; m < 256; m++) {
        int r = 0;
        for (int b = 0; b < 8; b++)
          if (m & (1 << b)) r |= 1 << (b ^ o);
        v[o][m] = (uint8_t)r;
      }
  }
};
#ifdef __CUDACC__
static __device__ const Perm8Table g_perm8_dev = Perm8Table();
#endif
static const Perm8Table g_perm8_host = Perm8Table();
PT_HD uint32_t pt_perm8(uint32_t m, uint32_t octinv) {
#ifdef __CUDA_ARCH__
  return g_perm8_dev.v[octinv][m];
#else
  return g_perm8_host.v[octinv][m];
#endif
}

// Pops the nearest pending child of node group `ng` (which must have node bits), fetches that
// node (5 x 128-bit), tests its 8 quantised child boxes against the ray clipped to [tmin, tbest]
// and returns: `rest` = what is left of the popped-from group (push it if it still has node bits),
// `ng` = hit inner children of the fetched node (octant-ordered priority bits 24..31 | imask),
// `lg` = hit leaf slots of the fetched node (leaf base, bit per leaf record).
PT_HD void bvh8_node_step(const SceneDev& sc, const RayCtx& r, float tbest, uint2& ng, uint2& lg, uint2& rest) {
  const uint32_t bit = (uint32_t)pt_bfind(ng.y);
  ng.y &= ~(1u << bit);
  rest = ng;
  const uint32_t imask_p = ng.y & 0xffu;
  const uint32_t slot = (bit - 24u) ^ r.octinv;
  const uint32_t rel = (uint32_t)pt_popc(imask_p & ~(0xffffffffu << slot));
  const uint32_t ni = ng.x + rel;

  const uint4* np = sc.nodes8 + (size_t)ni * 5;
  const uint4 n0 = PT_LDG4(np + 0);
  const uint4 n1 = PT_LDG4(np + 1);
  const uint4 n2 = PT_LDG4(np + 2);
  const uint4 n3 = PT_LDG4(np + 3);
  const uint4 n4 = PT_LDG4(np + 4);

  const float sx = pt_u2f((n0.w & 0xffu) << 23) * r.idx;
  const float sy = pt_u2f(((n0.w >> 8) & 0xffu) << 23) * r.idy;
  const float sz = pt_u2f(((n0.w >> 16) & 0xffu) << 23) * r.idz;
  const float bx = XFMA(-PT_QBIAS, sx, (pt_u2f(n0.x) - r.o.x) * r.idx);
  const float by = XFMA(-PT_QBIAS, sy, (pt_u2f(n0.y) - r.o.y) * r.idy);
  const float bz = XFMA(-PT_QBIAS, sz, (pt_u2f(n0.z) - r.o.z) * r.idz);
  const uint32_t qb = sc.qbias;

  // One fixed bit per hit child slot; inner children are then permuted into octant order by table and leaf children
  // (0.6 hit per node on average) expanded to their record bits in a short loop.  Empty slots carry the inverted box
  // (255 > 0), and even if one passed it would add nothing (meta 0, imask bit clear).  (Assembling `count << index`
  // per child under the hit predicate - 5 instructions for each of the 8 children whichever were hit, plus the
  // 4-wide field preparation - was 25 instructions longer per step: C2 -2.3 %, C4 equal; profiles/r01_hitmask_ab.md.)
  uint32_t hit8 = 0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t qlox = h ? n2.y : n2.x, qloy = h ? n2.w : n2.z, qloz = h ? n3.y : n3.x;
    const uint32_t qhix = h ? n3.w : n3.z, qhiy = h ? n4.y : n4.x, qhiz = h ? n4.w : n4.z;
    const uint32_t nx = r.idx < 0.f ? qhix : qlox, fx = r.idx < 0.f ? qlox : qhix;
    const uint32_t ny = r.idy < 0.f ? qhiy : qloy, fy = r.idy < 0.f ? qloy : qhiy;
    const uint32_t nz = r.idz < 0.f ? qhiz : qloz, fz = r.idz < 0.f ? qloz : qhiz;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float t0x = XFMA(pt_byte_to_float(nx, j, qb), sx, bx);
      const float t0y = XFMA(pt_byte_to_float(ny, j, qb), sy, by);
      const float t0z = XFMA(pt_byte_to_float(nz, j, qb), sz, bz);
      const float t1x = XFMA(pt_byte_to_float(fx, j, qb), sx, bx);
      const float t1y = XFMA(pt_byte_to_float(fy, j, qb), sy, by);
      const float t1z = XFMA(pt_byte_to_float(fz, j, qb), sz, bz);
      const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, r.tmin));
      const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tbest));
      if (cmin <= cmax) hit8 |= 1u << (4 * h + j);
    }
  }
  const uint32_t imask = n0.w >> 24;
  uint32_t leaf8 = hit8 & ~imask;
  uint32_t hitmask = pt_perm8(hit8 & imask, r.octinv) << 24;
  while (leaf8) {
    const uint32_t s = (uint32_t)pt_bfind(leaf8);
    leaf8 ^= 1u << s;
    const uint32_t m = ((s & 4u) ? n1.w : n1.z) >> ((s & 3u) * 8u);
    hitmask |= ((m >> 5) & 7u) << (m & 31u);
  }
  ng.x = n1.x;
  ng.y = (hitmask & 0xff000000u) | (n0.w >> 24);
  lg.x = n1.y;
  lg.y = hitmask & 0x00ffffffu;
}

#ifdef PT_DEFER_DIV
// Triangle test with the exact divisions deferred.  Leaf steps run the three IEEE divisions of Moeller-Trumbore with
// ~3 active lanes; a candidate that is CERTAINLY a hit and CERTAINLY closer than the best so far (every quotient at
// least 2^-18 inside its range - decided from numerators and |A| alone) does not need them yet: it becomes the best
// hit as numerators, with an upper bound of t for culling, and the divisions run when the ray retires (a batch of
// >= 20 lanes) or when an ambiguous candidate needs an exact comparison.  Same numerators, same XDIVs: the resolved
// t, u, v are bit-identical to the immediate path.  Returns true when `hit` changed.
PT_HD bool isect_triangle_deferred(float3 a, float3 b, float3 c, const RayCtx& r, TraceHit& hit, int leaf, int prim) {
  float3 e1 = xsub(b, a);
  float3 e2 = xsub(c, a);
  float3 P = xcross(r.d, e2);
  float A = xdot(P, e1);
  const float absA = fabsf(A);
  if ((double)absA < 1E-8) return false;
  float3 R = xsub(r.o, a);
  const float uN = xdot(P, R);
  const float hiA = absA * 1.000004f;  // > |A| (1 + 2^-18)
  const float tiny = absA * 1e-30f;
  const bool neg = A < 0.f;
  const bool opp_u = (uN < 0.f) != neg;
  if (fabsf(uN) > hiA) return false;
  if (opp_u && fabsf(uN) > tiny) return false;
  float3 Q = xcross(R, e1);
  const float vN = xdot(Q, r.d);
  const bool opp_v = (vN < 0.f) != neg;
  if (fabsf(vN) > hiA) return false;
  if (opp_v && fabsf(vN) > tiny) return false;
  const float tN = xdot(Q, e2);
  const float tbest = hit.t;  // exact, or an upper bound of a pending hit: either way nothing beyond it can win
  const bool opp_t = (tN < 0.f) != neg;
  if (r.tmin >= 0.f) {
    if (opp_t && fabsf(tN) > tiny) return false;
    if (tbest < 1e30f && fabsf(tN) > hiA * tbest) return false;
  }
  // certainly inside, certainly in the interval, certainly closer than the best so far?
  const float loA = absA * 0.999996f;  // < |A| (1 - 2^-18)
  const float eps = absA * 3.8147e-6f;  // |A| 2^-18
  const float lower = hit.den != 0.f ? tbest * 0.999996f : tbest;  // lower bound of the best t so far
  const bool certain = r.tmin >= 0.f && !opp_u && !opp_v && !opp_t && fabsf(uN) > eps && fabsf(vN) > eps &&
                       fabsf(uN) + fabsf(vN) < loA && fabsf(tN) > hiA * r.tmin && fabsf(tN) > eps && fabsf(tN) < loA * lower &&
                       lower < 1e30f;
  if (certain) {
#ifdef __CUDA_ARCH__
    hit.t = __fdividef(fabsf(tN), absA) * 1.000004f;
#else
    hit.t = (fabsf(tN) / absA) * 1.000004f;
#endif
    hit.den = A;
    hit.tn = tN;
    hit.u = uN;
    hit.v = vN;
    hit.leaf = leaf;
    hit.prim = prim;
    return true;
  }
  float u = XDIV(uN, A);
  if (u < 0.0f || u > 1.0f) return false;
  float v = XDIV(vN, A);
  if (v < 0.0f || XADD(u, v) > 1.0f) return false;
  float tt = XDIV(tN, A);
  if (!(tt > r.tmin && tt < r.tmax)) return false;
  hit_resolve(hit);  // an exact comparison needs the exact best
  if (tt < hit.t || (tt == hit.t && (hit.prim < 0 || prim < hit.prim))) {
    hit.t = tt;
    hit.leaf = leaf;
    hit.prim = prim;
    hit.u = u;
    hit.v = v;
    return true;
  }
  return false;
}
#endif

// Tests ONE pending leaf record of `lg` (highest bit) and updates the closest hit.
template <class Aux>
PT_HD void bvh8_leaf_step(const SceneDev& sc, const RayCtx& r, uint2& lg, TraceHit& hit, const Aux& aux) {
  const uint32_t b = (uint32_t)pt_bfind(lg.y);
  lg.y &= ~(1u << b);
  const uint32_t li = lg.x + b;
  const float4* lp = (const float4*)(sc.leaves + li);
  LeafRec L;
  L.a = PT_LDGF4(lp + 0);
  L.b = PT_LDGF4(lp + 1);
  L.c = PT_LDGF4(lp + 2);
#ifdef PT_DEFER_DIV
  if (PT_LEAF_KIND(pt_f2u(L.b.w)) == PT_PRIM_TRIANGLE) {
    isect_triangle_deferred(xyz(L.a), xyz(L.b), xyz(L.c), r, hit, (int)li, (int)pt_f2u(L.a.w));
    return;
  }
#endif
  float t, u, v;
  if (isect_leaf(sc, L, r.o, r.d, r.tmin, r.tmax, hit.t, t, u, v, aux)) {
    const int prim = (int)pt_f2u(L.a.w);
    hit_resolve(hit);
    if (t < hit.t || (t == hit.t && (hit.prim < 0 || prim < hit.prim))) {
      hit.t = t;
      hit.leaf = (int)li;
      hit.prim = prim;
      hit.u = u;
      hit.v = v;
    }
  }
}

#define PT_STACK 48

// Straightforward single-ray loop over the step functions (host simulation, reference for the
// warp-scheduled GPU driver in trace_sched.cuh).
struct FixedAux {
  static constexpr bool enabled = true;
  RayAux a;
  PT_HD RayAux operator()() const { return a; }
};
template <bool COUNT>
PT_HD void bvh8_closest(const SceneDev& sc, float3 o, float3 d, float tmin, float tmax,
                        TraceHit& hit, uint32_t& n_nodes, uint32_t& n_leaves, RayAux ra = RayAux{0.f, 0u},
                        int* overflow = nullptr) {
  const FixedAux aux{ra};
  RayCtx r;
  ray_setup(r, o, d, tmin, tmax);
  hit_reset(hit, tmax);
  uint2 stack[PT_STACK];
  int sp = 0;
  uint2 ng = make_uint2(0u, 0x80000000u);  // root
  uint2 lg = make_uint2(0u, 0u);
  for (;;) {
    if (ng.y & 0xff000000u) {
      uint2 rest;
      bvh8_node_step(sc, r, hit.t, ng, lg, rest);
      if (COUNT) n_nodes++;
      if (rest.y & 0xff000000u) {
        if (sp < PT_STACK) stack[sp++] = rest;
        else if (overflow) *overflow = 1;  // never silently wrong: the caller turns this into an error
      }
    }
    while (lg.y) {
      bvh8_leaf_step(sc, r, lg, hit, aux);
      if (COUNT) n_leaves++;
    }
    if ((ng.y & 0xff000000u) == 0) {
      if (sp == 0) break;
      ng = stack[--sp];
    }
  }
  hit_resolve(hit);
}
