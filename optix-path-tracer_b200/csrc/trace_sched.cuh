// trace_sched.cuh — warp-synchronous persistent traversal scheduler (device only).
//
// Why: one-ray-per-thread traversal leaves most SIMT lanes idle — rays of a warp finish
// after very different numbers of nodes, and node tests and leaf tests diverge.  The
// round-1 profile of the simple loop showed 6.9 active lanes per instruction (9 in the node
// test, 2-4 in the leaf test).  Here each warp is persistent and runs a vote-driven loop:
//
//   retire  lanes whose ray is finished emit their result (policy.emit)
//   refill  idle lanes fetch the next ray indices with ONE aggregated atomic per warp
//   vote    L = lanes with a pending leaf record, N = lanes with a pending node;
//           if L >= N the warp executes one leaf test for every L lane,
//           else one node step (pop nearest child, 5x128-bit fetch, 8 box tests) for every N lane.
//
// A lane with pending leaves waits for a leaf step, so leaf work accumulates across lanes
// until it is the majority: every step runs with at least half of the busy lanes active.
// The traversal stack (node groups only) lives in shared memory, PT_SMEM_STACK entries per
// thread, spilling to local memory beyond that.
//
// Policy interface (all __device__):
//   bool load(int i, float3& o, float3& d, float& tmin, float& tmax)   false = empty interval
//   bool any_hit(int i)                                                 true = stop at the first hit
//   RayAux aux(int i)                                                   shutter time + seed of ray i (read only
//                                                                        when a moving / volumetric leaf is tested)
//   void emit(int i, const TraceHit& h, unsigned finished_mask)        called convergently by
//                                                                        exactly the lanes in the mask
#pragma once
#include "traverse.h"

#ifndef PT_SMEM_STACK
#define PT_SMEM_STACK 12
#endif
#define PT_TRACE_THREADS 128

struct SchedParams {
  int refill_min;   // refill when at least this many lanes are idle (or the whole warp is)
  int leaf_thresh;  // run a leaf step when >= this many lanes have a pending leaf (or leaf lanes are the majority)
  int fetch_chunk;  // rays a warp claims from the queue per atomic (surplus stays warp-private); 0 = exactly the idle lanes
  int flat;         // tiny scene: skip the BVH, every lane tests every leaf record (trace_flat)
};

template <class Policy, bool SPECIAL>
struct PolicyAux {
  static constexpr bool enabled = SPECIAL;
  const Policy& pol;
  int ray;
  __device__ __forceinline__ RayAux operator()() const { return pol.aux(ray); }
};

// SPECIAL: the scene holds moving / volumetric primitives (SceneDev::special_mask != 0)
// BURST: node steps a lane may run back to back inside one vote round while it has nothing but node work.  Deep
// trees (triangle meshes: 11.5 node steps per ray on 1 M triangles) gain from the fewer vote rounds (+2.2 % on C4 with
// 2, +0.8 % with 3); shallow ones lose lane occupancy (C2: -4.5 %), and even the loop with a trip count of 1 costs
// 1.7 % - hence a compile-time parameter with the plain single step as its own instantiation.
template <bool COUNT, bool SPECIAL, int BURST = 1, class Policy>
__device__ __forceinline__ void trace_persistent(const SceneDev& sc, int n, int* fetch_counter, Policy& pol,
                                                 uint2* s_stack, const SchedParams sp_, int* error_flag, uint32_t& cn,
                                                 uint32_t& cl) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint2* my_stack = s_stack + threadIdx.x;  // entry k at my_stack[k * PT_TRACE_THREADS]
  uint2 lstack[PT_STACK - PT_SMEM_STACK];

  int ray = -1;
  bool anyhit = false;
  bool exhausted = false;  // warp-uniform: the ray queue is empty
  int w_next = 0, w_end = 0;  // warp-uniform: the warp's private claim [w_next, w_end) of queue indices
  RayCtx r;
  TraceHit hit;
  uint2 ng = make_uint2(0u, 0u), lg = make_uint2(0u, 0u);
  int sp = 0;
  hit_reset(hit, 0.f);

  for (;;) {
    // ---- pop / detect finished lanes ---------------------------------------------------------
    if (ray >= 0 && lg.y == 0u && (ng.y & 0xff000000u) == 0u && sp > 0) {
      --sp;
      ng = sp < PT_SMEM_STACK ? my_stack[sp * PT_TRACE_THREADS] : lstack[sp - PT_SMEM_STACK];
    }
    const bool finished = ray >= 0 && lg.y == 0u && (ng.y & 0xff000000u) == 0u;
    const unsigned fin = __ballot_sync(FULL, finished);
    unsigned idle = __ballot_sync(FULL, ray < 0);
    const unsigned busy = ~(fin | idle);
    // ---- retire + refill, batched: results are emitted (and their queue atomics issued) only when
    // enough lanes can be recycled at once, or when nothing else is left to do -------------------
    const bool can_recycle = fin != 0u || (!exhausted && idle != 0u);
    if (can_recycle && (__popc(fin | idle) >= sp_.refill_min || busy == 0u)) {
      if (finished) {
        hit_resolve(hit);  // deferred divisions of a triangle hit: here the retiring lanes do them together
        pol.emit(ray, hit, fin);
        ray = -1;
      }
      idle |= fin;
      if (!exhausted) {
        // same-address L2 atomics are the scarce resource of these kernels (one unit serialises them):
        // claim fetch_chunk indices per atomic and hand them out from the warp-private range
        const int cnt = __popc(idle);
        const int avail = w_end - w_next;
        const int chunk = max(sp_.fetch_chunk, cnt - avail);  // fetch_chunk 0: claim exactly what is needed
        int base = 0;
        if (avail < cnt) {
          if (lane == 0) base = atomicAdd(fetch_counter, chunk);
          base = __shfl_sync(FULL, base, 0);
        }
        if (ray < 0) {
          const int k = __popc(idle & lt_mask);
          const int idx = k < avail ? w_next + k : base + (k - avail);
          if (idx < n) {
            ray = idx;
            float3 o, d;
            float tmin, tmax;
            const bool ok = pol.load(idx, o, d, tmin, tmax);
            anyhit = pol.any_hit(idx);
            ray_setup(r, o, d, tmin, tmax);
            hit_reset(hit, tmax);
            ng = make_uint2(0u, ok ? 0x80000000u : 0u);  // root group (or nothing: reports a miss)
            lg = make_uint2(0u, 0u);
            sp = 0;
          }
        }
        if (avail < cnt) {
          w_next = base + (cnt - avail);
          w_end = base + chunk;
        } else {
          w_next += cnt;
        }
        if (w_next >= n) exhausted = true;
      }
      continue;  // re-evaluate lane states (fresh rays, rays loaded with an empty interval)
    }
    if (busy == 0u) break;  // nothing running, nothing to retire, queue exhausted
    // ---- vote + step ---------------------------------------------------------------------------
    const bool hasL = ray >= 0 && lg.y != 0u;
    const bool hasN = ray >= 0 && !hasL && (ng.y & 0xff000000u) != 0u;
    const unsigned mL = __ballot_sync(FULL, hasL);
    const unsigned mN = __ballot_sync(FULL, hasN);
    if (mL != 0u && (__popc(mL) >= sp_.leaf_thresh || __popc(mL) >= __popc(mN))) {
      if (hasL) {
        bvh8_leaf_step(sc, r, lg, hit, PolicyAux<Policy, SPECIAL>{pol, ray});
        if (COUNT) cl++;
        if (anyhit && hit.leaf >= 0) {  // occlusion query: any hit ends the ray
          ng.y = 0u;
          lg.y = 0u;
          sp = 0;
        }
      }
    } else if (mN != 0u) {
      if (hasN) {
        if (BURST <= 1) {
          uint2 rest;
          bvh8_node_step(sc, r, hit.t, ng, lg, rest);
          if (COUNT) cn++;
          if (rest.y & 0xff000000u) {
            if (sp < PT_SMEM_STACK) my_stack[sp * PT_TRACE_THREADS] = rest;
            else if (sp < PT_STACK) lstack[sp - PT_SMEM_STACK] = rest;
            if (sp < PT_STACK) sp++;
            else atomicExch(error_flag, 1);  // never silently wrong: the host turns this into an error
          }
        } else {
#pragma unroll 1
          for (int b = 0;;) {
            uint2 rest;
            bvh8_node_step(sc, r, hit.t, ng, lg, rest);
            if (COUNT) cn++;
            if (rest.y & 0xff000000u) {
              if (sp < PT_SMEM_STACK) my_stack[sp * PT_TRACE_THREADS] = rest;
              else if (sp < PT_STACK) lstack[sp - PT_SMEM_STACK] = rest;
              if (sp < PT_STACK) sp++;
              else atomicExch(error_flag, 1);
            }
            if (++b >= BURST || lg.y != 0u) break;  // a hit leaf waits for the warp's next leaf step
            if ((ng.y & 0xff000000u) == 0u) {
              if (sp == 0) break;
              --sp;
              ng = sp < PT_SMEM_STACK ? my_stack[sp * PT_TRACE_THREADS] : lstack[sp - PT_SMEM_STACK];
            }
          }
        }
      }
    }
  }
}

// ---- flat traversal for tiny scenes --------------------------------------------------------------
// Cornell-type scenes hold 7..16 primitives in ONE BVH8 node.  Walking that node through the
// vote-scheduled machinery costs a 320-instruction node test, a handful of scheduler rounds and leaf
// tests that run with 2-8 active lanes.  Here every lane tests every leaf record in the same order:
// the record address is warp-uniform (one broadcast load), the primitive kind is warp-uniform (no
// divergence), there is no stack and no scheduling.  Same closest hit as the BVH walk: exact tests,
// open interval, ties to the lowest primitive id; any-hit rays stop at their first hit.
template <bool COUNT, bool SPECIAL, class Policy>
__device__ __forceinline__ void trace_flat(const SceneDev& sc, int n, Policy& pol, uint32_t& cl) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * blockDim.x;
  const int n_leaves = sc.n_prims;
  for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {  // warp-uniform trip count
    const int i = i0 + lane;
    const bool act = i < n;
    TraceHit hit;
    hit_reset(hit, 0.f);
    if (act) {
      float3 o, d;
      float tmin, tmax;
      const bool ok = pol.load(i, o, d, tmin, tmax);
      const bool anyhit = pol.any_hit(i);
      hit_reset(hit, tmax);
      if (ok) {
        const PolicyAux<Policy, SPECIAL> aux{pol, i};
        for (int k = 0; k < n_leaves; k++) {
          const float4* lp = (const float4*)(sc.leaves + k);
          LeafRec L;
          L.a = __ldg(lp + 0);
          L.b = __ldg(lp + 1);
          L.c = __ldg(lp + 2);
          float t, u, v;
          if (isect_leaf(sc, L, o, d, tmin, tmax, hit.t, t, u, v, aux)) {
            const int prim = (int)pt_f2u(L.a.w);
            if (t < hit.t || (t == hit.t && (hit.prim < 0 || prim < hit.prim))) {
              hit.t = t;
              hit.leaf = k;
              hit.prim = prim;
              hit.u = u;
              hit.v = v;
            }
            if (anyhit) break;
          }
        }
        if (COUNT && lane == 0) cl += (uint32_t)n_leaves;  // one broadcast fetch per record serves the whole warp
      }
    }
    const unsigned m = __ballot_sync(FULL, act);
    if (act) pol.emit(i, hit, m);
  }
}
