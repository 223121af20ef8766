// integrator.cu — wavefront path integrator with path regeneration.
//
// Replaces the reference's megakernel launch loop (main.cpp:264-293 calling
// context->launch once per sample; programs/raygen.cu:66-167) with a pool of
// paths that is kept full: every iteration runs
//   k_iter_begin   1 thread: queue bookkeeping (counts stay on the device)
//   k_generate     pinhole/thin-lens camera rays for the free slots
//   k_trace        closest-hit BVH8 traversal, bins paths by material type
//                  (k_trace_deep: the same with two node steps per vote round,
//                  for scenes of >= 64 Ki primitives)
//   k_shade<MAT>   one kernel per material type present (+ k_miss); writes
//                  compacted continuing paths (warp-aggregated atomics) or a
//                  shadow request with both continuations (see shading.h)
//   k_shadow       shadow traversal + resolve + Russian roulette
// Path state moves with the queue (double-buffered pool of 64-byte records, 4
// float4 per path; shadow requests one array per float4), so appends are
// coalesced 128-bit stores; only the material-sorted shade kernels gather.
// Completion: k_iter_begin publishes its loop state in the first line of the control block and the host reads
// it with asynchronous copies that trail the launches by one batch, so the GPU never waits for the host.  Once
// every sample has been generated and at most PT_DRAIN_MAX paths are alive, one k_drain launch runs each of them
// to its end in a single thread (trace + shade + shadow in one loop) instead of ~45 nearly empty wavefront
// iterations of six or more kernels each.
#include <cooperative_groups.h>

#include <algorithm>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "renderer_internal.h"
#include "sample_order.h"
#include "scene_host.h"
#include "shading.h"
#include "trace_sched.cuh"
#include "traverse.h"

namespace cg = cooperative_groups;

namespace {

constexpr int NBINS = PT_MAT_TYPE_COUNT + 1;  // PT_MAT_* + miss
constexpr int BIN_MISS = PT_MAT_TYPE_COUNT;
constexpr int NSHADOW = 8;

// Path state: 64 B / path = 4 x float4
//   0 o_seed  origin.xyz, seed
//   1 d_pack  dir.xyz, pix | depth<<23 | prevSpecular<<31
//   2 thr     throughput.xyz, BSDF pdf of this ray (env NEE MIS)
//   3 rad     radiance.xyz, prd.time (shutter time of the sample, raygen.cu:69)
// Shadow request: 7 float4 (+ an 8th, s[7] = (prd.time, seed at shadow time), only in scenes with
// moving / volumetric primitives).
// Layout (measured, profiles/r01_layout_ab.md): the path state is one 64-byte record per path (AoS of
// float4) - appends by consecutive slots stay coalesced either way, and the material-sorted shade
// kernels, which GATHER by path index, touch two full 32-byte sectors per path instead of a half-used
// sector in each of four arrays (C2 +5 %, C3 +2 %).  Shadow requests stay one array per float4: they are
// written once by consecutive slots and most of a 128-byte record would be a partial-line write
// (AoS requests: C1 -9 %).  -DPT_PATH_SOA / -DPT_SHADOW_AOS build the other layouts.
#ifndef PT_SHADOW_AOS
#define PT_SHADOW_SOA
#endif
#ifdef PT_PATH_SOA
struct PathBuf {
  float4* a[4];
  __device__ __forceinline__ float4& f(int k, int i) const { return a[k][i]; }
};
#else
struct PathBuf {
  float4* base;
  __device__ __forceinline__ float4& f(int k, int i) const { return base[(size_t)i * 4 + k]; }
};
#endif
#ifdef PT_SHADOW_SOA
struct ShadowBuf {
  float4* a[NSHADOW];
  __device__ __forceinline__ float4& s(int k, int i) const { return a[k][i]; }
};
#else
struct ShadowBuf {
  float4* base;
  __device__ __forceinline__ float4& s(int k, int i) const { return base[(size_t)i * NSHADOW + k]; }
};
#endif
#define P_OSEED 0
#define P_DPACK 1
#define P_THR 2
#define P_RAD 3

// Device control block.  Every atomically updated counter sits on its own 128-byte line: queue
// tails, bin counters and fetch cursors are hammered by warp-aggregated atomics from all SMs while
// other fields are read as loop bounds, and sharing L2 lines between them (or with whatever small
// allocation happens to be adjacent) made whole kernels up to 50 % slower depending on placement.
#ifndef PT_CTR_STRIDE
#define PT_CTR_STRIDE 1024  // measured: 128/256/512-B spacing still left +-5 % placement luck (L2 atomic units pair
                            // neighbouring lines); at 1 KiB the C1 throughput no longer depends on where Ctrl lands
#endif
struct alignas(PT_CTR_STRIDE) PaddedInt {
  int v;
  int pad[PT_CTR_STRIDE / 4 - 1];
};
struct alignas(PT_CTR_STRIDE) Ctrl {
  // written by k_iter_begin only, read as loop bounds; the first 64 bytes are what the host polls
  int n_cur, gen_base, gen_n;
  int first_frame;
  unsigned long long sample_base, next_sample, total_samples;
  unsigned long long stat_closest, stat_shadow, stat_iters;
  // atomics
  PaddedInt n_next_, n_shadow_, fetch_closest_, fetch_shadow_;
  PaddedInt bin_count[NBINS];
  alignas(PT_CTR_STRIDE) unsigned long long stat_nodes;
  alignas(PT_CTR_STRIDE) unsigned long long stat_leaves;
  alignas(PT_CTR_STRIDE) int error;  // 1 = traversal stack overflow
};
// what the host polls: the leading bytes of Ctrl, written by k_iter_begin
struct CtrlPoll {
  int n_cur, gen_base, gen_n, first_frame;
  unsigned long long sample_base, next_sample, total_samples;
};
#define n_next n_next_.v
#define n_shadow n_shadow_.v
#define fetch_closest fetch_closest_.v
#define fetch_shadow fetch_shadow_.v

struct Params {
  int W, H, max_depth, russian, pool;  // (the first frame index lives in Ctrl so captured graphs stay valid)
  int tiled;                           // pixels enumerated in 8x4 tiles, Morton order inside (W % 8 == 0 && H % 4 == 0)
  // An image of more than PT_BAND_PIXELS pixels is rendered as horizontal bands, one after the other: W x H are the
  // rows [y0, y0 + H) of a W x Himg image, path records carry the pixel index WITHIN the band (23 bits), and the
  // seed, the camera ray and the accumulation address use the image row y0 + y.  One band: y0 = 0, H = Himg.
  int y0, Himg, ytop;  // ytop = Himg - 1 - y0: the accumulation buffer's row of the band's first image row
  int group_log2;      // camera samples are enumerated in groups of 2^group_log2 frames (k_generate)
  float4* accum;
  uint32_t c16, c1;  // the constants 16 and 1 as run-time values, see pt_tea64_balanced
};

// tea<64> (programs/random.cuh:33-47) is 64 rounds of shift / add / xor: as written every operation
// lands on the ALU pipe (LEA, LEA.HI, VIADD, LOP3 - 9 of 10 instructions per round) while the FMA
// pipe idles, and one seed per camera ray made k_generate 6-7 % of the headline run.  `(v << 4) + k`
// and `v + s` are integer multiply-adds; with 16 and 1 passed in as run-time values ptxas cannot
// turn them back into shifts, so they issue as IMAD on the FMA pipe and the two pipes carry
// 2.5 operations per half-round each.  Integer arithmetic: bit-identical to pt_tea64.
__device__ __forceinline__ uint32_t pt_tea64_balanced(uint32_t val0, uint32_t val1, uint32_t c16, uint32_t c1) {
  uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll 8
  for (int n = 0; n < 64; n++) {
    s0 += 0x9e3779b9u;
    v0 += (v1 * c16 + 0xa341316cu) ^ (v1 * c1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += (v0 * c16 + 0xad90777du) ^ (v0 * c1 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}

__device__ __forceinline__ uint32_t pack_state(uint32_t pix, int depth, bool spec) {
  return pix | ((uint32_t)depth << 23) | (spec ? 0x80000000u : 0u);
}

__device__ __forceinline__ void accumulate(const Params& p, uint32_t pix, float3 col) {
  col = pt_de_nan(col);
  int x = pix % p.W, y = pix / p.W;
  float4* dst = p.accum + (size_t)(p.ytop - y) * p.W + x;  // raygen.cu:153 (row Himg - 1 - (y0 + y))
  atomicAdd(dst, make_float4(col.x, col.y, col.z, 1.f));
}

// warp-aggregated append: one atomic per converged group
__device__ __forceinline__ int push_slot(int* counter) {
  auto g = cg::coalesced_threads();
  int base = 0;
  if (g.thread_rank() == 0) base = atomicAdd(counter, (int)g.size());
  base = g.shfl(base, 0);
  return base + (int)g.thread_rank();
}

// raygen.cu:97-118 tail of one loop iteration: Russian roulette and the depth cut-off.  Returns true when
// the path continues (thr, depth, seed updated); otherwise the sample has been accumulated.
__device__ __forceinline__ bool bounce_survives(const Params& p, uint32_t pix, int& depth, float3& thr, float3 rad,
                                                uint32_t& seed) {
  if (p.russian) {
    float prob = pt_ffmax(pt_ffmax(thr.x, thr.y), thr.z);
    if (depth > 10) {
      if (pt_rnd(seed) >= prob) {
        accumulate(p, pix, rad + thr);
        return false;
      }
      thr *= 1.f / prob;
    }
  }
  depth++;
  if (depth >= p.max_depth) {  // raygen.cu:122: recursion did not terminate -> black
    accumulate(p, pix, f3(0.f));
    return false;
  }
  return true;
}
// ... then the path is appended to the next wavefront
__device__ __forceinline__ void finish_bounce(const Params& p, Ctrl* c, const PathBuf& next, uint32_t pix,
                                              int depth, float3 o, float3 d, float3 thr, float3 rad,
                                              uint32_t seed, bool specular, float pdf, float time) {
  if (!bounce_survives(p, pix, depth, thr, rad, seed)) return;
  int slot = push_slot(&c->n_next);
  next.f(P_OSEED, slot) = make_float4(o.x, o.y, o.z, __uint_as_float(seed));
  next.f(P_DPACK, slot) = make_float4(d.x, d.y, d.z, __uint_as_float(pack_state(pix, depth, specular)));
  next.f(P_THR, slot) = make_float4(thr.x, thr.y, thr.z, pdf);  // .w: BSDF pdf of this ray (env NEE MIS)
  next.f(P_RAD, slot) = make_float4(rad.x, rad.y, rad.z, time);
}

__global__ void k_iter_begin(Ctrl* c, int pool) {
  c->stat_shadow += (unsigned long long)c->n_shadow;
  int n = c->n_next;  // survivors of the previous iteration (buffers were swapped by the host)
  unsigned long long remaining = c->total_samples - c->next_sample;
  int room = pool - n;
  int g = (int)(remaining < (unsigned long long)room ? remaining : (unsigned long long)room);
  c->gen_base = n;
  c->gen_n = g;
  c->sample_base = c->next_sample;
  c->next_sample += (unsigned long long)g;
  c->n_cur = n + g;
  c->n_next = 0;
  c->n_shadow = 0;
  c->fetch_closest = 0;
  c->fetch_shadow = 0;
  for (int i = 0; i < NBINS; i++) c->bin_count[i].v = 0;
  c->stat_closest += (unsigned long long)c->n_cur;
  c->stat_iters += 1;
}

// camera rays for the free slots of the pool, in the order sample_order.h defines (frames in groups, pixels in tiles)
__global__ void k_generate(SceneDev sc, Params p, const Ctrl* c, PathBuf cur) {
  const int n = c->gen_n, base = c->gen_base;
  const unsigned long long sbase = c->sample_base;
  const uint32_t npix = (uint32_t)p.W * (uint32_t)p.H;
  // sample id -> (whole frames before it, offset): 64-bit divisions once per thread; inside the loop the offset from
  // the first sample of this launch is < npix + pool < 2^26, so 32-bit arithmetic does
  const uint32_t fb0 = (uint32_t)(sbase / npix);
  const uint32_t rem0 = (uint32_t)(sbase - (unsigned long long)fb0 * npix);
  const SampleOrder so = sample_order_make(p.W, p.H, (uint32_t)(c->total_samples / npix), p.group_log2, p.tiled);
  const int first_frame = c->first_frame;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t t = rem0 + (uint32_t)i;
    const uint32_t q = t / npix;
    int x, y;
    uint32_t fr;
    sample_order_decode(so, fb0 + q, t - q * npix, x, y, fr);
    const int frame = first_frame + (int)fr;
    const uint32_t pix = (uint32_t)y * (uint32_t)p.W + (uint32_t)x;
    uint32_t seed;
    float3 o, d;
    const int yi = y + p.y0;  // row of the image (the band starts at y0)
    seed = pt_tea64_balanced((uint32_t)p.W * (uint32_t)yi + (uint32_t)x, (uint32_t)frame, p.c16, p.c1);
    camera_ray_seeded(sc.camera, p.W, p.Himg, x, yi, seed, o, d);
    const float time = sc.camera.time0 + pt_rnd(seed) * (sc.camera.time1 - sc.camera.time0);  // raygen.cu:69
    int s = base + i;
    cur.f(P_OSEED, s) = make_float4(o.x, o.y, o.z, __uint_as_float(seed));
    cur.f(P_DPACK, s) = make_float4(d.x, d.y, d.z, __uint_as_float(pack_state(pix, 0, false)));
    cur.f(P_THR, s) = make_float4(1.f, 1.f, 1.f, 0.f);
    cur.f(P_RAD, s) = make_float4(0.f, 0.f, 0.f, time);
  }
}

// closest-hit traversal of the current wavefront: persistent warps, vote-scheduled (trace_sched.cuh)
struct ClosestPolicy {
  const SceneDev& sc;
  Ctrl* c;
  PathBuf cur;
  float4* hits;
  uint32_t* bin_q;
  int pool;
  __device__ __forceinline__ bool load(int i, float3& o, float3& d, float& tmin, float& tmax) const {
    float4 o4 = cur.f(P_OSEED, i), d4 = cur.f(P_DPACK, i);
    int depth = (__float_as_uint(d4.w) >> 23) & 0xff;
    o = f3(o4.x, o4.y, o4.z);
    d = f3(d4.x, d4.y, d4.z);
    tmin = depth == 0 ? 1e-6f : 1e-3f;  // raygen.cu:61 camera rays, :102 bounced rays
    tmax = PT_RAY_MAX;
    return true;
  }
  __device__ __forceinline__ bool any_hit(int) const { return false; }
  __device__ __forceinline__ RayAux aux(int i) const {
    return RayAux{cur.f(P_RAD, i).w, __float_as_uint(cur.f(P_OSEED, i).w)};
  }
  __device__ __forceinline__ void emit(int i, const TraceHit& h, unsigned fin) const {
#ifdef PT_HITS_BY_PATH
    hits[i] = make_float4(h.t, __int_as_float(h.leaf), h.u, h.v);
#endif
    int key = BIN_MISS;
    if (h.leaf >= 0) key = (int)(__float_as_uint(__ldg(&sc.leaves[h.leaf].c.w)) >> 24);
    // material binning: finished lanes with the same key share one atomic
    unsigned peers = __match_any_sync(fin, key);
    int leader = __ffs(peers) - 1;
    int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == leader) base = atomicAdd(&c->bin_count[key].v, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const size_t slot = (size_t)key * pool + base + __popc(peers & ((1u << lane) - 1));
    bin_q[slot] = (uint32_t)i;
#ifndef PT_HITS_BY_PATH
    // the hit record goes into the BIN slot, next to the queue entry: the shade kernels then read it with
    // coalesced loads (indexed by path it was a 16-byte gather that cost a 64-byte DRAM access per path)
    if (key != BIN_MISS) hits[slot] = make_float4(h.t, __int_as_float(h.leaf), h.u, h.v);
#endif
  }
};

#ifndef PT_TRACE_MINBLOCKS
#define PT_TRACE_MINBLOCKS 8  // 64 registers: measured best (9 and 10 CTAs/SM at 56 / 48 registers were 2-5 % slower)
#endif
template <bool COUNT, bool SPECIAL>
__global__ void __launch_bounds__(PT_TRACE_THREADS, PT_TRACE_MINBLOCKS) k_trace(SceneDev sc, Ctrl* c, PathBuf cur, float4* hits,
                                                            uint32_t* bin_q, int pool, SchedParams sched) {
  __shared__ uint2 s_stack[PT_SMEM_STACK * PT_TRACE_THREADS];
  ClosestPolicy pol{sc, c, cur, hits, bin_q, pool};
  uint32_t cn = 0, cl = 0;
  if (sched.flat) trace_flat<COUNT, SPECIAL>(sc, c->n_cur, pol, cl);
  else trace_persistent<COUNT, SPECIAL>(sc, c->n_cur, &c->fetch_closest, pol, s_stack, sched, &c->error, cn, cl);
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      cn += __shfl_xor_sync(0xffffffffu, cn, o);
      cl += __shfl_xor_sync(0xffffffffu, cl, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&c->stat_nodes, (unsigned long long)cn);
      atomicAdd(&c->stat_leaves, (unsigned long long)cl);
    }
  }
}

// closest-hit traversal of deep trees: two node steps per vote round (trace_sched.cuh, BURST); never flat
#ifndef PT_DEEP_BURST
#define PT_DEEP_BURST 2
#endif
template <bool COUNT, bool SPECIAL>
__global__ void __launch_bounds__(PT_TRACE_THREADS, PT_TRACE_MINBLOCKS) k_trace_deep(SceneDev sc, Ctrl* c, PathBuf cur, float4* hits,
                                                                 uint32_t* bin_q, int pool, SchedParams sched) {
  __shared__ uint2 s_stack[PT_SMEM_STACK * PT_TRACE_THREADS];
  ClosestPolicy pol{sc, c, cur, hits, bin_q, pool};
  uint32_t cn = 0, cl = 0;
  trace_persistent<COUNT, SPECIAL, PT_DEEP_BURST>(sc, c->n_cur, &c->fetch_closest, pol, s_stack, sched, &c->error, cn, cl);
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      cn += __shfl_xor_sync(0xffffffffu, cn, o);
      cl += __shfl_xor_sync(0xffffffffu, cl, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&c->stat_nodes, (unsigned long long)cn);
      atomicAdd(&c->stat_leaves, (unsigned long long)cl);
    }
  }
}

#ifndef PT_SHADE_MINBLOCKS
#define PT_SHADE_MINBLOCKS 8
#endif
template <int MAT, int LK>
__global__ void __launch_bounds__(128, PT_SHADE_MINBLOCKS) k_shade(SceneDev sc, Params p, Ctrl* c, PathBuf cur, PathBuf next, ShadowBuf sb,
                                               const float4* hits, const uint32_t* bin_q) {
  const int n = c->bin_count[MAT].v;
  const uint32_t* q = bin_q + (size_t)MAT * p.pool;
  // (a software pipeline of this gather - queue entry read two iterations ahead, prefetch.global.L2 of the
  // next path record and hit - was measured 4 % SLOWER on C2: the kernels are not latency-bound on it)
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int i = (int)q[j];
#ifdef PT_HITS_BY_PATH
    const float4 h4 = hits[i];
#else
    const float4 h4 = hits[(size_t)MAT * p.pool + j];
#endif
    float4 o4 = cur.f(P_OSEED, i), d4 = cur.f(P_DPACK, i), t4 = cur.f(P_THR, i), r4 = cur.f(P_RAD, i);
    uint32_t pk = __float_as_uint(d4.w);
    PathIn in;
    in.o = f3(o4.x, o4.y, o4.z);
    in.d = f3(d4.x, d4.y, d4.z);
    in.thr = f3(t4.x, t4.y, t4.z);
    in.rad = f3(r4.x, r4.y, r4.z);
    in.seed = __float_as_uint(o4.w);
    in.depth = (pk >> 23) & 0xff;
    in.prev_specular = (pk >> 31) != 0;
    in.last_pdf = t4.w;
    in.time = r4.w;
    const uint32_t pix = pk & 0x7fffffu;
    ShadeOut out;
    shade_hit<MAT, LK>(sc, in, h4.x, __float_as_int(h4.y), h4.z, h4.w, out);
    if (out.kind == SH_TERMINATE) {
      accumulate(p, pix, out.col);
    } else if (out.kind == SH_BOUNCE) {
      finish_bounce(p, c, next, pix, in.depth, out.P, out.dir, out.thr, out.rad, out.seed, out.specular, out.pdf,
                    in.time);
    } else {
      int s = push_slot(&c->n_shadow);
      uint32_t pk2 = pack_state(pix, in.depth, out.terminal);
      sb.s(0, s) = make_float4(out.P.x, out.P.y, out.P.z, __uint_as_float(out.seed));
      sb.s(1, s) = make_float4(out.WiL.x, out.WiL.y, out.WiL.z, __uint_as_float(out.seedS));
      sb.s(2, s) = make_float4(out.dir.x, out.dir.y, out.dir.z, __uint_as_float(pk2));
      sb.s(3, s) = make_float4(out.dirS.x, out.dirS.y, out.dirS.z, out.thr.x);
      sb.s(4, s) = make_float4(out.thr.y, out.thr.z, out.rad.x, out.rad.y);
      sb.s(5, s) = make_float4(out.rad.z, out.thrS.x, out.thrS.y, out.thrS.z);
      sb.s(6, s) = make_float4(out.radS.x, out.radS.y, out.radS.z, out.env_shadow ? out.pdf : -1.f);
#if defined(PT_SHADOW_SOA) && !defined(PT_SHADOW_FULL)
      if (sc.special_mask)
#endif
        sb.s(7, s) = make_float4(in.time, __uint_as_float(out.seedShadow), 0.f, 0.f);
    }
  }
}

// (evaluating cheap miss programs - gradient, constant - in k_trace's epilogue instead was measured SLOWER:
// C2 -3 %, C3 -12 %, and C4 -4 % although its HDR background does not even take that branch; the traversal
// kernel sits at its 64-register budget and pays for every instruction added to it)
__global__ void __launch_bounds__(128) k_miss(SceneDev sc, Params p, Ctrl* c, PathBuf cur, const uint32_t* bin_q) {
  const int n = c->bin_count[BIN_MISS].v;
  const uint32_t* q = bin_q + (size_t)BIN_MISS * p.pool;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int i = (int)q[j];
    float4 d4 = cur.f(P_DPACK, i), t4 = cur.f(P_THR, i), r4 = cur.f(P_RAD, i);
    PathIn in;
    in.d = f3(d4.x, d4.y, d4.z);
    in.thr = f3(t4.x, t4.y, t4.z);
    in.rad = f3(r4.x, r4.y, r4.z);
    in.last_pdf = t4.w;
    accumulate(p, __float_as_uint(d4.w) & 0x7fffffu, shade_miss(sc, in));
  }
}

// shadow ray: light_sample.cuh:62-72 + hit.cu:16-24 under the pinned semantics
// (visible <=> closest hit in (1e-3, inf) is a Diffuse_Light); then resolve + Russian roulette
struct ShadowPolicy {
  const SceneDev& sc;
  const Params& p;
  Ctrl* c;
  PathBuf next;
  ShadowBuf sb;
  __device__ __forceinline__ bool load(int i, float3& o, float3& d, float& tmin, float& tmax) const {
    float4 s0 = sb.s(0, i), s1 = sb.s(1, i);
    o = f3(s0.x, s0.y, s0.z);
    d = f3(s1.x, s1.y, s1.z);
    tmin = 1e-3f;
    tmax = PT_RAY_MAX;
    return true;
  }
  // environment shadow rays are occlusion queries; rect-light shadow rays need the closest hit
#ifdef PT_NO_ANYHIT
  __device__ __forceinline__ bool any_hit(int) const { return false; }
#else
  __device__ __forceinline__ bool any_hit(int i) const { return sb.s(6, i).w >= 0.f; }
#endif
  // PINNED: the reference's volume programs would read prd.seed through a PerRayData_Shadow
  // payload (undefined); shadow rays use prd.seed as it is when Direct_Light traces them
  __device__ __forceinline__ RayAux aux(int i) const {
    const float4 s7 = sb.s(7, i);
    return RayAux{s7.x, __float_as_uint(s7.y)};
  }
  __device__ __forceinline__ void emit(int i, const TraceHit& h, unsigned) const {
    float4 s0 = sb.s(0, i), s1 = sb.s(1, i), s2 = sb.s(2, i), s3 = sb.s(3, i), s4 = sb.s(4, i), s5 = sb.s(5, i),
           s6 = sb.s(6, i);
    const bool env = s6.w >= 0.f;  // shadow ray towards the environment: visible <=> nothing hit
    bool visible = env ? h.leaf < 0 : false;
    if (!env && h.leaf >= 0) visible = (__float_as_uint(__ldg(&sc.leaves[h.leaf].c.w)) >> 24) == PT_MAT_DIFFUSE_LIGHT;
    uint32_t pk = __float_as_uint(s2.w);
    uint32_t pix = pk & 0x7fffffu;
    int depth = (pk >> 23) & 0xff;
    bool terminal = (pk >> 31) != 0;
    float3 dir, thr, rad;
    uint32_t seed;
    if (visible) {
      dir = f3(s2.x, s2.y, s2.z);
      thr = f3(s3.w, s4.x, s4.y);
      rad = f3(s4.z, s4.w, s5.x);
      seed = __float_as_uint(s0.w);
    } else {
      dir = f3(s3.x, s3.y, s3.z);
      thr = f3(s5.y, s5.z, s5.w);
      rad = f3(s6.x, s6.y, s6.z);
      seed = __float_as_uint(s1.w);
    }
    if (terminal) accumulate(p, pix, rad);
    else finish_bounce(p, c, next, pix, depth, f3(s0.x, s0.y, s0.z), dir, thr, rad, seed, false, env ? s6.w : 0.f,
                       sc.special_mask ? sb.s(7, i).x : 0.f);
  }
};

template <bool COUNT, bool SPECIAL>
__global__ void __launch_bounds__(PT_TRACE_THREADS, PT_TRACE_MINBLOCKS) k_shadow(SceneDev sc, Params p, Ctrl* c, PathBuf next, ShadowBuf sb,
                                                             SchedParams sched) {
  __shared__ uint2 s_stack[PT_SMEM_STACK * PT_TRACE_THREADS];
  ShadowPolicy pol{sc, p, c, next, sb};
  uint32_t cn = 0, cl = 0;
  if (sched.flat) trace_flat<COUNT, SPECIAL>(sc, c->n_shadow, pol, cl);
  else trace_persistent<COUNT, SPECIAL>(sc, c->n_shadow, &c->fetch_shadow, pol, s_stack, sched, &c->error, cn, cl);
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      cn += __shfl_xor_sync(0xffffffffu, cn, o);
      cl += __shfl_xor_sync(0xffffffffu, cl, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&c->stat_nodes, (unsigned long long)cn);
      atomicAdd(&c->stat_leaves, (unsigned long long)cl);
    }
  }
}

// standalone traversal of a ray set (pt_trace_rays); counters: [0] nodes, [1] leaves, [2] fetch cursor, [3] error flag
struct RaySetPolicy {
  const pt_ray* rays;
  pt_hit* hits;
  float time;  // ray sets carry no per-ray data: every ray sees the scene at the camera's time0, seed = ray index
  __device__ __forceinline__ RayAux aux(int i) const { return RayAux{time, (uint32_t)i}; }
  __device__ __forceinline__ bool load(int i, float3& o, float3& d, float& tmin, float& tmax) const {
    const float4* rp = (const float4*)(rays + i);
    float4 a = __ldg(rp), b = __ldg(rp + 1);
    o = f3(a.x, a.y, a.z);
    d = f3(b.x, b.y, b.z);
    tmin = a.w;
    tmax = b.w;
    return b.w > a.w;
  }
  __device__ __forceinline__ bool any_hit(int) const { return false; }
  __device__ __forceinline__ void emit(int i, const TraceHit& h, unsigned) const {
    pt_hit o;
    o.t = h.t;
    o.prim = h.prim;
    o.u = h.u;
    o.v = h.v;
    hits[i] = o;
  }
};

template <bool COUNT, bool SPECIAL>
__global__ void __launch_bounds__(PT_TRACE_THREADS, PT_TRACE_MINBLOCKS) k_trace_rays(SceneDev sc, const pt_ray* rays, int n, pt_hit* hits,
                                                                 unsigned long long* counters, SchedParams sched) {
  __shared__ uint2 s_stack[PT_SMEM_STACK * PT_TRACE_THREADS];
  RaySetPolicy pol{rays, hits, sc.camera.time0};
  uint32_t cn = 0, cl = 0;
  if (sched.flat) trace_flat<COUNT, SPECIAL>(sc, n, pol, cl);
  else trace_persistent<COUNT, SPECIAL>(sc, n, (int*)(counters + 2), pol, s_stack, sched, (int*)(counters + 3), cn, cl);
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      cn += __shfl_xor_sync(0xffffffffu, cn, o);
      cl += __shfl_xor_sync(0xffffffffu, cl, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(counters, (unsigned long long)cn);
      atomicAdd(counters + 1, (unsigned long long)cl);
    }
  }
}


// ray sets on deep trees: two node steps per vote round, like k_trace_deep (never flat)
template <bool COUNT, bool SPECIAL>
__global__ void __launch_bounds__(PT_TRACE_THREADS, PT_TRACE_MINBLOCKS) k_trace_rays_deep(SceneDev sc, const pt_ray* rays, int n, pt_hit* hits,
                                                                      unsigned long long* counters, SchedParams sched) {
  __shared__ uint2 s_stack[PT_SMEM_STACK * PT_TRACE_THREADS];
  RaySetPolicy pol{rays, hits, sc.camera.time0};
  uint32_t cn = 0, cl = 0;
  trace_persistent<COUNT, SPECIAL, PT_DEEP_BURST>(sc, n, (int*)(counters + 2), pol, s_stack, sched, (int*)(counters + 3), cn, cl);
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      cn += __shfl_xor_sync(0xffffffffu, cn, o);
      cl += __shfl_xor_sync(0xffffffffu, cl, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(counters, (unsigned long long)cn);
      atomicAdd(counters + 1, (unsigned long long)cl);
    }
  }
}

// ---- drain: the last few thousand paths, one thread each, run to their end -------------------------------
// After the last sample has been generated a render used to finish in ~45 nearly empty wavefront iterations (a few
// paths survive Russian roulette up to depth 50): each costs the dependent latency of one traversal plus six or
// more kernel boundaries, ~2 % of a C2 step and far more of a short per-rank step under a sample split.  Here each
// surviving path is walked by one thread through the same functions the wavefront kernels call (bvh8_closest is
// the single-ray loop over the same node / leaf steps; shade_hit / shade_miss / bounce_survives are shared), so
// the RNG stream and every arithmetic operation of a path are unchanged.
#ifndef PT_DRAIN_MAX
#define PT_DRAIN_MAX 16384
#endif
template <int LK>
__device__ __noinline__ void drain_shade(const SceneDev& sc, int mat, const PathIn& in, const TraceHit& h, ShadeOut& out) {
  switch (mat) {
    case PT_MAT_LAMBERTIAN: shade_hit<PT_MAT_LAMBERTIAN, LK>(sc, in, h.t, h.leaf, h.u, h.v, out); break;
    case PT_MAT_METAL: shade_hit<PT_MAT_METAL, 0>(sc, in, h.t, h.leaf, h.u, h.v, out); break;
    case PT_MAT_DIELECTRIC: shade_hit<PT_MAT_DIELECTRIC, 0>(sc, in, h.t, h.leaf, h.u, h.v, out); break;
    case PT_MAT_DIFFUSE_LIGHT: shade_hit<PT_MAT_DIFFUSE_LIGHT, LK>(sc, in, h.t, h.leaf, h.u, h.v, out); break;
    case PT_MAT_OREN_NAYAR: shade_hit<PT_MAT_OREN_NAYAR, LK>(sc, in, h.t, h.leaf, h.u, h.v, out); break;
    case PT_MAT_TORRANCE_SPARROW: shade_hit<PT_MAT_TORRANCE_SPARROW, 0>(sc, in, h.t, h.leaf, h.u, h.v, out); break;
    case PT_MAT_ASHIKHMIN_SHIRLEY: shade_hit<PT_MAT_ASHIKHMIN_SHIRLEY, 0>(sc, in, h.t, h.leaf, h.u, h.v, out); break;
    case PT_MAT_NORMAL: shade_hit<PT_MAT_NORMAL, 0>(sc, in, h.t, h.leaf, h.u, h.v, out); break;
    default: shade_hit<PT_MAT_ISOTROPIC, LK>(sc, in, h.t, h.leaf, h.u, h.v, out); break;
  }
}

template <int LK>
__global__ void __launch_bounds__(64) k_drain(SceneDev sc, Params p, Ctrl* c, PathBuf cur, int count_fetches) {
  const int n = c->n_next;  // survivors of the last wavefront iteration, appended to `cur`
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 o4 = cur.f(P_OSEED, i), d4 = cur.f(P_DPACK, i), t4 = cur.f(P_THR, i), r4 = cur.f(P_RAD, i);
  const uint32_t pk = __float_as_uint(d4.w);
  const uint32_t pix = pk & 0x7fffffu;
  PathIn in;
  in.o = f3(o4.x, o4.y, o4.z);
  in.d = f3(d4.x, d4.y, d4.z);
  in.thr = f3(t4.x, t4.y, t4.z);
  in.rad = f3(r4.x, r4.y, r4.z);
  in.seed = __float_as_uint(o4.w);
  in.depth = (pk >> 23) & 0xff;
  in.prev_specular = (pk >> 31) != 0;
  in.last_pdf = t4.w;
  in.time = r4.w;
  uint32_t q_closest = 0, q_shadow = 0, cn = 0, cl = 0;
  int overflow = 0;
  for (;;) {
    TraceHit h;
    q_closest++;
    bvh8_closest<true>(sc, in.o, in.d, in.depth == 0 ? 1e-6f : 1e-3f, PT_RAY_MAX, h, cn, cl, RayAux{in.time, in.seed}, &overflow);
    if (h.leaf < 0) {
      accumulate(p, pix, shade_miss(sc, in));
      break;
    }
    const int mat = (int)(__float_as_uint(__ldg(&sc.leaves[h.leaf].c.w)) >> 24);
    ShadeOut out;
    drain_shade<LK>(sc, mat, in, h, out);
    if (out.kind == SH_TERMINATE) {
      accumulate(p, pix, out.col);
      break;
    }
    float3 dir = out.dir, thr = out.thr, rad = out.rad;
    uint32_t seed = out.seed;
    bool specular = out.specular;
    float pdf = out.pdf;
    if (out.kind == SH_SHADOW) {  // ShadowPolicy::emit, inline
      q_shadow++;
      TraceHit sh;
      bvh8_closest<true>(sc, out.P, out.WiL, 1e-3f, PT_RAY_MAX, sh, cn, cl, RayAux{in.time, out.seedShadow}, &overflow);
      const bool env = out.env_shadow && out.pdf >= 0.f;  // the request's tag (k_shade: s6.w = env ? pdf : -1)
      bool visible = env ? sh.leaf < 0 : false;
      if (!env && sh.leaf >= 0)
        visible = (__float_as_uint(__ldg(&sc.leaves[sh.leaf].c.w)) >> 24) == PT_MAT_DIFFUSE_LIGHT;
      if (!visible) {
        dir = out.dirS;
        thr = out.thrS;
        rad = out.radS;
        seed = out.seedS;
      }
      if (out.terminal) {
        accumulate(p, pix, rad);
        break;
      }
      specular = false;
      pdf = env ? out.pdf : 0.f;
    }
    int depth = in.depth;
    if (!bounce_survives(p, pix, depth, thr, rad, seed)) break;
    in.o = out.P;
    in.d = dir;
    in.thr = thr;
    in.rad = rad;
    in.seed = seed;
    in.depth = depth;
    in.prev_specular = specular;
    in.last_pdf = pdf;
  }
  atomicAdd(&c->stat_closest, (unsigned long long)q_closest);
  if (q_shadow) atomicAdd(&c->stat_shadow, (unsigned long long)q_shadow);
  if (count_fetches) {
    atomicAdd(&c->stat_nodes, (unsigned long long)cn);
    atomicAdd(&c->stat_leaves, (unsigned long long)cl);
  }
  if (overflow) atomicExch(&c->error, 1);
}
// after the drain nothing is alive
__global__ void k_drain_end(Ctrl* c) {
  c->n_next = 0;
  c->n_cur = 0;
}

// raygen.cu:76,122 with maxDepth == 0: the loop body never runs, every sample is black but still counted
__global__ void k_count_samples(float4* accum, int n, float nframes) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) accum[i].w += nframes;
}

// make_Color (raygen.cu:136-146): the preview pixel of the display buffer
__global__ void k_display(const float4* __restrict__ accum, uchar4* __restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 c = accum[i];
    const float inv = 1.0f / (c.w > 0.f ? c.w : 1.f);  // float3 / (frame + 1) = float3 * (1 / s)
    const float r = fminf(fmaxf(sqrtf(c.x * inv), 0.f), 1.f);
    const float g = fminf(fmaxf(sqrtf(c.y * inv), 0.f), 1.f);
    const float b = fminf(fmaxf(sqrtf(c.z * inv), 0.f), 1.f);
    out[i] = make_uchar4((unsigned char)(int)(255.99 * (double)r), (unsigned char)(int)(255.99 * (double)g),
                         (unsigned char)(int)(255.99 * (double)b), (unsigned char)(int)(255.99 * 1.0));
  }
}

// per-device caches, indexed by device ordinal (a multi-GPU renderer drives several devices from one process)
int g_sm_count[PT_MAX_DEVICES] = {};
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < PT_MAX_DEVICES ? dev : 0;
}
int sm_count() {
  int dev = current_device();
  if (!g_sm_count[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_sm_count[dev] = n > 0 ? n : 148;
  }
  return g_sm_count[dev];
}

// scheduler knobs; PTB200_REFILL_MIN / PTB200_LEAF_THRESH override the tuned defaults (dev only)
// scenes with at most this many primitives are traversed flat (trace_flat); PTB200_FLAT_MAX overrides (dev)
int flat_max() {
  static int v = [] {
    const char* e = getenv("PTB200_FLAT_MAX");
    return e ? atoi(e) : 16;
  }();
  return v;
}
SchedParams sched_params_base() {
  static SchedParams p = [] {
    // refill threshold re-tuned for the final kernels (profiles/r01_sched_sweep3.txt): 12 -> 20 is +4 % on C2 / C3,
    // neutral on the triangle scene; 28 costs it 9 %
    SchedParams q{20, 8, 0, 0};
    if (const char* e = getenv("PTB200_REFILL_MIN")) q.refill_min = std::max(1, atoi(e));
    if (const char* e = getenv("PTB200_LEAF_THRESH")) q.leaf_thresh = std::max(1, atoi(e));
    if (const char* e = getenv("PTB200_FETCH_CHUNK")) q.fetch_chunk = std::max(0, atoi(e));
    return q;
  }();
  return p;
}
SchedParams sched_params(const SceneDev& sc) {
  SchedParams p = sched_params_base();
  p.flat = sc.n_prims <= flat_max() ? 1 : 0;
  return p;
}
// ray sets handed to pt_trace_rays are arbitrary (the benchmark's are incoherent secondary rays): their lanes
// finish at very different times and the smaller retire batch wins (round 1: 12 instead of 20, 1.73 vs 1.62 Gray/s on
// 1 M triangles; with the round-2 tree 8: 2.05 vs 2.00 Gray/s, 10 and 14 in between)
SchedParams sched_params_rays(const SceneDev& sc) {
  SchedParams p = sched_params(sc);
  if (!getenv("PTB200_REFILL_MIN")) p.refill_min = 8;
  return p;
}

// frames per group in the enumeration of the camera samples (k_generate), as a power of two; PTB200_FRAME_GROUP
// overrides (dev: 1, 2, 4, ... 256 frames)
// (measured, profiles/r02_frame_groups.md: every configuration gains monotonically up to 128-256 frames per group)
#ifndef PT_FRAME_GROUP_LOG2
#define PT_FRAME_GROUP_LOG2 8
#endif
int frame_group_log2() {
  int k = PT_FRAME_GROUP_LOG2;
  if (const char* e = getenv("PTB200_FRAME_GROUP")) {
    const int f = atoi(e);
    k = 0;
    while (k < 8 && (2 << k) <= f) k++;
  }
  return k;
}

// deep trees (>= PTB200_DEEP_MIN primitives, default 64 Ki) take two node steps per vote round
long long deep_min_prims() {
  static const long long v = getenv("PTB200_DEEP_MIN") ? atoll(getenv("PTB200_DEEP_MIN")) : 65536;
  return v;
}
bool rays_deep(const SceneDev& sc, const SchedParams& sched) {
  static const int on = getenv("PTB200_DEEP_RAYS") ? atoi(getenv("PTB200_DEEP_RAYS")) : 1;
  return on && deep_min_prims() > 0 && sc.n_prims >= deep_min_prims() && !sched.flat;
}

// resident CTAs per SM of the persistent traversal kernels (they must all be co-resident)
int g_trace_bpm[PT_MAX_DEVICES][2] = {};
int trace_blocks_per_sm(bool special) {
  int& bpm = g_trace_bpm[current_device()][special ? 1 : 0];
  if (!bpm) {
    int a = 0, b = 0, c = 0, d = 0;
    if (special) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_trace<false, true>, PT_TRACE_THREADS, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d, k_trace_deep<false, true>, PT_TRACE_THREADS, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_shadow<false, true>, PT_TRACE_THREADS, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_trace_rays<false, true>, PT_TRACE_THREADS, 0);
      int c2 = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c2, k_trace_rays_deep<false, true>, PT_TRACE_THREADS, 0);
      c = std::min(c, c2);
    } else {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_trace<false, false>, PT_TRACE_THREADS, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d, k_trace_deep<false, false>, PT_TRACE_THREADS, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_shadow<false, false>, PT_TRACE_THREADS, 0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_trace_rays<false, false>, PT_TRACE_THREADS, 0);
      int c2 = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c2, k_trace_rays_deep<false, false>, PT_TRACE_THREADS, 0);
      c = std::min(c, c2);
    }
    bpm = std::max(1, std::min(std::min(a, d), std::min(b, c)));
#ifdef PT_TRACE_BPM
    bpm = PT_TRACE_BPM;
#endif
  }
  return bpm;
}

// launches KERNEL<COUNT, SPECIAL> for the run-time (count, special) pair
#define PT_LAUNCH_TRACE(KERNEL, count, special, grid, stream, ...)                                        \
  do {                                                                                                    \
    if (special) {                                                                                        \
      if (count) KERNEL<true, true><<<grid, PT_TRACE_THREADS, 0, stream>>>(__VA_ARGS__);                  \
      else KERNEL<false, true><<<grid, PT_TRACE_THREADS, 0, stream>>>(__VA_ARGS__);                       \
    } else {                                                                                              \
      if (count) KERNEL<true, false><<<grid, PT_TRACE_THREADS, 0, stream>>>(__VA_ARGS__);                 \
      else KERNEL<false, false><<<grid, PT_TRACE_THREADS, 0, stream>>>(__VA_ARGS__);                      \
    }                                                                                                     \
  } while (0)

}  // namespace

// ---------------------------------------------------------------------------------------------
struct pt_renderer {
  pt_scene* scene = nullptr;
  int device = 0;  // the device everything below lives on (= the scene's)
  int W = 0, H = 0;
  int flags = 0;
  long long pool_req = 0;
  int pool = 0;
  DevBuf accum_own, display;
  float4* accum = nullptr;
  // all per-path arrays live in one slab; array k starts at a skewed offset so that the same
  // path index in different arrays never maps to the same DRAM/L2 partition (arrays of exactly
  // 2^26 bytes allocated back to back alias badly: measured -30 % on NEE scenes)
  DevBuf slab, ctrl;
  float4 *pathA[4] = {}, *pathB[4] = {}, *shadow[NSHADOW] = {}, *hits = nullptr;  // AoS: element [0] is the base
  uint32_t* bin_q = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_poll[2] = {nullptr, nullptr};
  bool started = false;
  // the wavefront iteration captured as two CUDA graphs (pools A->B and B->A), reused across
  // pt_render calls while the launch parameters stay the same
  cudaGraphExec_t gexec[2] = {nullptr, nullptr};
  int g_max_depth = -1, g_russian = -1, g_pool = -1, g_flags = -1, g_launches = 0;
  int g_y0 = -1, g_rows = -1;  // the band of image rows the graphs were captured for
  int g_group = -1;            // ... and the frame grouping of the camera samples
  unsigned long long g_commit_gen = 0;
  SceneDev g_scene{};  // kernel arguments are baked into the graphs: re-capture when they change
  pt_stats stats{};
  // pinned host memory of this renderer: the full control block (final statistics), two poll records that
  // trail the launches, and the staging buffer of pt_read_accum for pageable destinations (grow-only)
  Ctrl* h_ctrl = nullptr;
  CtrlPoll* h_poll = nullptr;
  float* stage = nullptr;
  size_t stage_bytes = 0;
  pt_multi* multi = nullptr;  // != null: this handle fans out to one renderer per device (multi.cu)
};

// Process-level cache (guarded by a mutex; one slot per device): the path-pool slab of a destroyed renderer is
// parked and handed to the next renderer on the same device.  Creating a renderer per job otherwise frees and
// re-allocates 2-3 GB each time, and the stream-ordered allocator answered that churn with re-mapping stalls of
// 100-700 ms (seen as e2e C4 1346 -> 1053 Msamples/s).  pt_release_caches() empties it.
static DevBuf g_slab_cache[PT_MAX_DEVICES];
static std::mutex g_cache_mu;

static int ensure_pool(pt_renderer* r, long long want) {
  // Paths in flight.  Longer launches amortise the tail of every persistent traversal launch (the last warps of a
  // launch run alone) and the per-iteration kernel boundaries: 8 Mi -> 32 Mi paths is +5.3 % on C4 (+4.8 % even on a
  // 64-spp share of it, i.e. four fills), +5.2 % on C3, +7 % on 10 M triangles.  The shading kernels pay for it: they
  // gather path records by index over a buffer four times the size (C2, three material bins, shading 31 -> 34 % of the
  // step while its traversal gets 2 % faster: -3 % in total).  So: traversal-dominated scenes (the deep ones, >= 64 Ki
  // primitives) take 32 Mi paths (14 GB of path state, shadow requests, hit records and bins, of the 180), the others
  // 1/32 of the job between 8 Mi (3.5 GB) and 32 Mi.
  const bool deep_scene = deep_min_prims() > 0 && r->scene->dev.n_prims >= deep_min_prims();
  long long pool = r->pool_req > 0 ? r->pool_req
                   : deep_scene   ? (1ll << 25)
                                  : std::max<long long>(1ll << 23, std::min<long long>(1ll << 25, want / 32));
  pool = std::min<long long>(pool, std::max<long long>(want, 1024));
  pool = std::min<long long>(pool, 1ll << 25);
  if (pool <= r->pool) return PT_OK;
  r->pool = (int)pool;
  const size_t arr = (size_t)pool * 16;
  const size_t skew = 33 * 256;  // odd multiple of 256 B
  const size_t gap = 64 * 1024;
  const int n_arr = 4 + 4 + NSHADOW + NBINS;
  const size_t total = (arr + gap + 7 * skew + 256) * n_arr + (size_t)pool * 4 * NBINS + 64 * 1024;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    DevBuf& parked = g_slab_cache[r->device];
    if (parked.p && parked.bytes >= total && !r->slab.p) {
      std::swap(r->slab.p, parked.p);
      std::swap(r->slab.bytes, parked.bytes);
      std::swap(r->slab.dev, parked.dev);
    }
  }
  if (!r->slab.p || r->slab.bytes < total) PT_CUDA(r->slab.alloc(total));
  char* base = r->slab.as<char>();
  size_t off = 0;
  int k = 0;
  auto carve = [&](size_t bytes) {
    char* p = base + off + (size_t)(k % 7) * skew;
    off += bytes + gap + 7 * skew;
    off = (off + 255) & ~(size_t)255;
    k++;
    return p;
  };
#ifdef PT_PATH_SOA
  for (int i = 0; i < 4; i++) r->pathA[i] = (float4*)carve(arr);
  for (int i = 0; i < 4; i++) r->pathB[i] = (float4*)carve(arr);
#else
  r->pathA[0] = (float4*)carve(arr * 4);
  r->pathB[0] = (float4*)carve(arr * 4);
#endif
#ifdef PT_SHADOW_SOA
  for (int i = 0; i < NSHADOW; i++) r->shadow[i] = (float4*)carve(arr);
#else
  r->shadow[0] = (float4*)carve(arr * NSHADOW);
#endif
#ifdef PT_HITS_BY_PATH
  r->hits = (float4*)carve(arr);
#else
  r->hits = (float4*)carve(arr * (NBINS - 1));  // one pool-sized segment per material bin
#endif
  r->bin_q = (uint32_t*)(base + off);
  return PT_OK;
}

// ---- internal accessors for multi.cu ------------------------------------------------------------------
float4* pt_r_accum(pt_renderer* r) { return r->accum; }
cudaStream_t pt_r_stream(pt_renderer* r) { return r->stream; }
int pt_r_device(pt_renderer* r) { return r->device; }
int pt_r_width(pt_renderer* r) { return r->W; }
int pt_r_height(pt_renderer* r) { return r->H; }
void pt_r_mark_started(pt_renderer* r) { r->started = true; }
void pt_r_set_ms_reduce(pt_renderer* r, float ms) { r->stats.ms_reduce = ms; }
pt_multi* pt_r_multi(pt_renderer* r) { return r->multi; }
pt_renderer* pt_r_new_multi_handle(pt_scene* scene, int w, int h, pt_multi* m) {
  pt_renderer* r = new pt_renderer();
  r->scene = scene;
  r->W = w;
  r->H = h;
  r->multi = m;
  return r;
}
void pt_r_delete_handle(pt_renderer* r) { delete r; }
void pt_release_slab_cache() {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  for (auto& b : g_slab_cache) b.release();
}

namespace {
// ends a stream capture that an error interrupted, so the stream is usable again
struct CaptureGuard {
  cudaStream_t st;
  bool active = false;
  ~CaptureGuard() {
    if (!active) return;
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(st, &g);
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
  }
};
struct EventSet {  // per-kernel-class timing events of one pt_render call
  cudaEvent_t e[6] = {};
  ~EventSet() {
    for (auto& ev : e)
      if (ev) cudaEventDestroy(ev);
  }
};
void drop_graphs(pt_renderer* r) {
  for (auto& g : r->gexec) {
    if (g) cudaGraphExecDestroy(g);
    g = nullptr;
  }
  r->g_max_depth = -1;
}
}  // namespace

static int render_rows(pt_renderer* r, int y0, int rows, int32_t first_frame, int32_t nframes, int32_t max_depth,
                       int32_t russian);

extern "C" {

int pt_renderer_create(pt_scene* scene, int32_t width, int32_t height, void* d_accum, pt_renderer** out) {
  if (!scene || !out || width <= 0 || height <= 0) {
    pt_set_error("pt_renderer_create: bad arguments");
    return PT_ERR_INVALID;
  }
  if (!scene->committed) {
    pt_set_error("pt_renderer_create: scene not committed");
    return PT_ERR_STATE;
  }
  if ((long long)width * height > PT_MAX_PIXELS || width > PT_BAND_PIXELS) {
    pt_set_error("pt_renderer_create: more than 2^28 pixels, or a row of more than 2^23");
    return PT_ERR_INVALID;
  }
  DeviceGuard dg(scene->device);
  pt_renderer* r = new pt_renderer();
  r->scene = scene;
  r->device = scene->device >= 0 && scene->device < PT_MAX_DEVICES ? scene->device : 0;
  r->W = width;
  r->H = height;
  if (d_accum) {
    r->accum = (float4*)d_accum;
  } else {
    if (r->accum_own.alloc((size_t)width * height * 16) != cudaSuccess) {
      pt_set_error("accumulation buffer allocation failed");
      delete r;
      return PT_ERR_OOM;
    }
    r->accum = r->accum_own.as<float4>();
  }
  if (cudaMemset(r->accum, 0, (size_t)width * height * 16) != cudaSuccess ||
      r->ctrl.alloc(sizeof(Ctrl)) != cudaSuccess || cudaStreamCreate(&r->stream) != cudaSuccess ||
      cudaEventCreate(&r->ev0) != cudaSuccess || cudaEventCreate(&r->ev1) != cudaSuccess ||
      cudaEventCreateWithFlags(&r->ev_poll[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&r->ev_poll[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaMallocHost((void**)&r->h_ctrl, sizeof(Ctrl) + 2 * sizeof(CtrlPoll)) != cudaSuccess) {
    pt_set_error("renderer setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    pt_renderer_destroy(r);
    return PT_ERR_CUDA;
  }
  r->h_poll = (CtrlPoll*)(r->h_ctrl + 1);
  *out = r;
  return PT_OK;
}

int pt_renderer_destroy(pt_renderer* r) {
  if (!r) return PT_OK;
  if (r->multi) {
    int rc = pt_multi_destroy(r->multi);
    delete r;
    return rc;
  }
  DeviceGuard dg(r->device);
  if (r->stream) {
    cudaStreamSynchronize(r->stream);
    cudaStreamDestroy(r->stream);
  }
  if (r->ev0) cudaEventDestroy(r->ev0);
  if (r->ev1) cudaEventDestroy(r->ev1);
  for (auto& e : r->ev_poll)
    if (e) cudaEventDestroy(e);
  if (r->h_ctrl) cudaFreeHost(r->h_ctrl);
  if (r->stage) cudaFreeHost(r->stage);
  if (r->slab.p) {  // park the larger slab, free the smaller
    std::lock_guard<std::mutex> lk(g_cache_mu);
    DevBuf& parked = g_slab_cache[r->device];
    if (r->slab.bytes > parked.bytes) {
      std::swap(r->slab.p, parked.p);
      std::swap(r->slab.bytes, parked.bytes);
      std::swap(r->slab.dev, parked.dev);
    }
  }
  for (auto& g : r->gexec)
    if (g) cudaGraphExecDestroy(g);
  delete r;
  return PT_OK;
}

int pt_renderer_set_flags(pt_renderer* r, int32_t flags) {
  if (!r) return PT_ERR_INVALID;
  if (r->multi) return pt_multi_set_flags(r->multi, flags);
  r->flags = flags;
  return PT_OK;
}
int pt_renderer_set_pool(pt_renderer* r, int64_t paths) {
  if (!r || paths < 0) return PT_ERR_INVALID;
  if (r->multi) return pt_multi_set_pool(r->multi, paths);
  r->pool_req = paths;
  return PT_OK;
}

int pt_render(pt_renderer* r, int32_t first_frame, int32_t nframes, int32_t max_depth, int32_t russian,
              int32_t clear_accum) {
  if (!r || nframes < 0 || max_depth < 0 || max_depth > PT_MAX_DEPTH) {
    pt_set_error("pt_render: bad arguments (max_depth must be 0..255: the path state packs the depth into 8 bits)");
    return PT_ERR_INVALID;
  }
  if (r->multi) return pt_multi_render(r->multi, first_frame, nframes, max_depth, russian, clear_accum);
  pt_scene* s = r->scene;
  if (!s->committed) {
    // add_* / set_* calls and a failed commit release or replace device buffers the kernels would read
    pt_set_error("pt_render: the scene was modified after the last successful pt_scene_commit");
    return PT_ERR_STATE;
  }
  DeviceGuard dg(r->device);
  const long long total = (long long)r->W * r->H * nframes;
  // Images of more than PT_BAND_PIXELS pixels are rendered in bands of whole rows, one after the other (the path
  // record has 23 bits for the pixel index; see Params).  Band heights are multiples of 4 so that the 8x4 tiling of
  // the camera samples applies.  PTB200_BAND_PIXELS (dev / tests) lowers the band size.
  long long band_pixels = PT_BAND_PIXELS;
  if (const char* e = getenv("PTB200_BAND_PIXELS")) band_pixels = std::max<long long>(1, std::min<long long>(atoll(e), PT_BAND_PIXELS));
  int band_rows = r->H;
  if ((long long)r->W * r->H > band_pixels) {
    band_rows = (int)std::max<long long>(1, band_pixels / r->W);
    if (band_rows >= 4) band_rows &= ~3;
  }
  int rc = ensure_pool(r, (long long)r->W * band_rows * nframes);
  if (rc) return rc;
  cudaStream_t st = r->stream;
  if (clear_accum || (first_frame == 0 && !r->started))
    PT_CUDA(cudaMemsetAsync(r->accum, 0, (size_t)r->W * r->H * 16, st));
  r->started = true;
  if (total == 0 || max_depth == 0) {
    // raygen.cu:76,122: the loop body never runs, every sample is black (but counted)
    if (total) {
      k_count_samples<<<sm_count() * 4, 256, 0, st>>>(r->accum, r->W * r->H, (float)nframes);
      PT_CUDA(cudaGetLastError());
      PT_CUDA(cudaStreamSynchronize(st));
    }
    r->stats = pt_stats{};
    r->stats.samples = (uint64_t)total;
    return PT_OK;
  }
  pt_stats sum{};
  for (int y0 = 0; y0 < r->H; y0 += band_rows) {
    rc = render_rows(r, y0, std::min(band_rows, r->H - y0), first_frame, nframes, max_depth, russian);
    if (rc) return rc;
    const pt_stats& b = r->stats;
    sum.samples += b.samples;
    sum.closest_rays += b.closest_rays;
    sum.shadow_rays += b.shadow_rays;
    sum.nodes_fetched += b.nodes_fetched;
    sum.leaves_fetched += b.leaves_fetched;
    sum.iterations += b.iterations;
    sum.kernel_launches += b.kernel_launches;
    sum.ms_total += b.ms_total;
    sum.ms_generate += b.ms_generate;
    sum.ms_trace_closest += b.ms_trace_closest;
    sum.ms_shade += b.ms_shade;
    sum.ms_trace_shadow += b.ms_trace_shadow;
    sum.drained_paths |= b.drained_paths;
  }
  r->stats = sum;
  return PT_OK;
}

}  // extern "C"

// the wavefront loop over rows [y0, y0 + rows) of the image: all samples of `nframes` frames of those pixels
static int render_rows(pt_renderer* r, int y0, int rows, int32_t first_frame, int32_t nframes, int32_t max_depth,
                       int32_t russian) {
  pt_scene* s = r->scene;
  cudaStream_t st = r->stream;
  const long long total = (long long)r->W * rows * nframes;
  Ctrl init{};
  init.total_samples = (unsigned long long)total;
  init.first_frame = first_frame;
  PT_CUDA(cudaMemcpyAsync(r->ctrl.p, &init, sizeof(init), cudaMemcpyHostToDevice, st));

  Params p;
  p.W = r->W;
  p.H = rows;
  p.y0 = y0;
  p.Himg = r->H;
  p.ytop = r->H - 1 - y0;
  p.max_depth = max_depth;
  p.russian = russian;
  p.pool = r->pool;
  p.tiled = (r->W % 8 == 0 && rows % 4 == 0 && !getenv("PTB200_NO_TILES")) ? 1 : 0;
  p.group_log2 = frame_group_log2();
  while (p.group_log2 > 0 && ((long long)r->W * rows << p.group_log2) > (1ll << 30)) p.group_log2--;  // 32-bit ids in k_generate
  p.accum = r->accum;
  p.c16 = 16u;
  p.c1 = 1u;
  PathBuf A, B;
  ShadowBuf sb;
#ifdef PT_PATH_SOA
  for (int k = 0; k < 4; k++) {
    A.a[k] = r->pathA[k];
    B.a[k] = r->pathB[k];
  }
#else
  A.base = r->pathA[0];
  B.base = r->pathB[0];
#endif
#ifdef PT_SHADOW_SOA
  for (int k = 0; k < NSHADOW; k++) sb.a[k] = r->shadow[k];
#else
  sb.base = r->shadow[0];
#endif
  Ctrl* c = r->ctrl.as<Ctrl>();
  float4* hits = r->hits;
  uint32_t* bq = r->bin_q;
  const SceneDev& sc = s->dev;
  const bool count = (r->flags & 1) != 0;
  const bool timing = (r->flags & 2) != 0;
  const int sms = sm_count();
  const bool special = sc.special_mask != 0;
  const int g_trace = sms * trace_blocks_per_sm(special), g_shade = sms * PT_SHADE_MINBLOCKS;
  const int g_gen = sms * 8;  // 2048 threads per SM hide the 64-round dependency chain of tea<64> (4 -> 8 blocks: +0.3 %)
#ifndef PT_MISS_BLOCKS
#define PT_MISS_BLOCKS 8
#endif
  const int g_miss = sms * PT_MISS_BLOCKS;
  const uint32_t mask = s->material_type_mask;
  const SchedParams sched = sched_params(sc);
  // deep trees (>= 64 Ki primitives: five or more BVH8 levels) take two node steps per vote round for closest-hit
  // rays; PTB200_DEEP_MIN overrides the size (dev), 0 switches it off
  const bool deep = deep_min_prims() > 0 && sc.n_prims >= deep_min_prims() && !sched.flat;
  // shadow rays keep the BVH walk even in tiny scenes: they start on a surface and aim at a light, the one
  // node test culls most primitives, and testing all of them flat was 7 % slower end to end (C1, C3);
  // closest-hit rays are 25-45 % faster flat.  PTB200_FLAT_SHADOW=1 overrides (dev)
  SchedParams sched_shadow = sched;
  if (!getenv("PTB200_REFILL_MIN")) sched_shadow.refill_min = 28;  // shadow rays are short: larger retire batches pay (C3 +3 %)
  {
    static const int fs = getenv("PTB200_FLAT_SHADOW") ? atoi(getenv("PTB200_FLAT_SHADOW")) : 0;
    sched_shadow.flat = sched.flat && fs != 0;
  }
  const bool need_shadow = (mask & ((1u << PT_MAT_LAMBERTIAN) | (1u << PT_MAT_OREN_NAYAR) | (1u << PT_MAT_DIFFUSE_LIGHT) |
                                    (1u << PT_MAT_ISOTROPIC))) &&
                           (s->lights.size() > 0 || sc.env_n > 0);

  bool sphere_lights = false;
  for (const pt_light& l : s->lights) sphere_lights |= l.type == PT_LIGHT_SPHERE;

  EventSet evs;
  cudaEvent_t* e = evs.e;
  float ms_gen = 0, ms_trace = 0, ms_shade = 0, ms_shadow = 0;
  if (timing)
    for (int k = 0; k < 6; k++) PT_CUDA(cudaEventCreate(&e[k]));

  // one wavefront iteration reading pool `cur`, appending survivors to pool `nxt`; returns #launches
  auto iteration = [&](const PathBuf& cur, const PathBuf& nxt, bool timed) {
    int n = 0;
    k_iter_begin<<<1, 1, 0, st>>>(c, r->pool);
    if (timed) cudaEventRecord(e[0], st);
    k_generate<<<g_gen, 256, 0, st>>>(sc, p, c, cur);
    if (timed) cudaEventRecord(e[1], st);
    if (deep) PT_LAUNCH_TRACE(k_trace_deep, count, special, g_trace, st, sc, c, cur, hits, bq, r->pool, sched);
    else PT_LAUNCH_TRACE(k_trace, count, special, g_trace, st, sc, c, cur, hits, bq, r->pool, sched);
    if (timed) cudaEventRecord(e[2], st);
    n += 3;
#define LAUNCH_SHADE(M, LK)                                                     \
  if (mask & (1u << M)) {                                                       \
    k_shade<M, LK><<<g_shade, 128, 0, st>>>(sc, p, c, cur, nxt, sb, hits, bq);  \
    n++;                                                                        \
  }
    // the three NEE materials exist in a rectangle-lights-only and an any-light build
    if (sphere_lights) {
      LAUNCH_SHADE(PT_MAT_LAMBERTIAN, 1)
      LAUNCH_SHADE(PT_MAT_DIFFUSE_LIGHT, 1)
      LAUNCH_SHADE(PT_MAT_OREN_NAYAR, 1)
      LAUNCH_SHADE(PT_MAT_ISOTROPIC, 1)
    } else {
      LAUNCH_SHADE(PT_MAT_LAMBERTIAN, 0)
      LAUNCH_SHADE(PT_MAT_DIFFUSE_LIGHT, 0)
      LAUNCH_SHADE(PT_MAT_OREN_NAYAR, 0)
      LAUNCH_SHADE(PT_MAT_ISOTROPIC, 0)
    }
    LAUNCH_SHADE(PT_MAT_METAL, 0)
    LAUNCH_SHADE(PT_MAT_DIELECTRIC, 0)
    LAUNCH_SHADE(PT_MAT_TORRANCE_SPARROW, 0)
    LAUNCH_SHADE(PT_MAT_ASHIKHMIN_SHIRLEY, 0)
    LAUNCH_SHADE(PT_MAT_NORMAL, 0)
#undef LAUNCH_SHADE
    k_miss<<<g_miss, 128, 0, st>>>(sc, p, c, cur, bq);
    if (timed) cudaEventRecord(e[3], st);
    n++;
    if (need_shadow) {
      PT_LAUNCH_TRACE(k_shadow, count, special, g_trace, st, sc, p, c, nxt, sb, sched_shadow);
      n++;
    }
    if (timed) {
      cudaEventRecord(e[4], st);
      cudaEventSynchronize(e[4]);
      float t;
      cudaEventElapsedTime(&t, e[0], e[1]); ms_gen += t;
      cudaEventElapsedTime(&t, e[1], e[2]); ms_trace += t;
      cudaEventElapsedTime(&t, e[2], e[3]); ms_shade += t;
      cudaEventElapsedTime(&t, e[3], e[4]); ms_shadow += t;
    }
    return n;
  };

  // Capture the iteration once per parameter set as two CUDA graphs (A->B, B->A): replaying a graph
  // removes the per-kernel launch gaps, which are ~10 % of an iteration on small scenes.  The
  // per-kernel timing mode launches directly instead.
  const bool use_graph = !timing && getenv("PTB200_NO_GRAPH") == nullptr;
  if (use_graph && (r->gexec[0] == nullptr || r->gexec[1] == nullptr || r->g_max_depth != max_depth ||
                    r->g_russian != russian || r->g_pool != r->pool || r->g_flags != r->flags || r->g_y0 != y0 ||
                    r->g_rows != rows || r->g_group != p.group_log2 || r->g_commit_gen != s->commit_gen || memcmp(&r->g_scene, &sc, sizeof(SceneDev)) != 0)) {
    drop_graphs(r);
    for (int parity = 0; parity < 2; parity++) {
      cudaGraph_t graph = nullptr;
      CaptureGuard cap{st};
      PT_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      cap.active = true;
      r->g_launches = parity == 0 ? iteration(A, B, false) : iteration(B, A, false);
      cap.active = false;
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (ce == cudaSuccess) ce = cudaGraphInstantiate(&r->gexec[parity], graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (ce != cudaSuccess) {
        drop_graphs(r);
        pt_set_error("pt_render: capturing the wavefront iteration failed: %s", cudaGetErrorString(ce));
        cudaGetLastError();
        return PT_ERR_CUDA;
      }
    }
    r->g_max_depth = max_depth;
    r->g_russian = russian;
    r->g_pool = r->pool;
    r->g_flags = r->flags;
    r->g_y0 = y0;
    r->g_rows = rows;
    r->g_group = p.group_log2;
    r->g_commit_gen = s->commit_gen;
    memcpy(&r->g_scene, &sc, sizeof(SceneDev));
  }

  PT_CUDA(cudaEventRecord(r->ev0, st));
  unsigned long long launches = 0;
  long long iter = 0;
  // Completion detection without stalling the device.  k_iter_begin leaves its loop state in the first bytes of
  // the control block; after every batch of iterations the host enqueues an asynchronous copy of those bytes and
  // only then waits for the copy of the PREVIOUS batch, so one batch is always queued behind the one being looked
  // at.  While more than a pool-full of samples is ungenerated at least remaining / pool further iterations are
  // certain and the batch is that long; near the end it is 4.  Batches are even so that the pool parity of
  // iteration k is k & 1 and the survivors of a finished batch are in pool A.
  const int drain_max = getenv("PTB200_DRAIN_MAX") ? std::min(atoi(getenv("PTB200_DRAIN_MAX")), 1 << 20) : PT_DRAIN_MAX;  // 0 = off (dev)
  const bool use_drain = drain_max > 0 && !timing;
  auto even = [](long long v) { return v + (v & 1); };
  long long batch = even(std::max<long long>(4, total / std::max(1, r->pool)));
  int slot = 0;
  bool outstanding = false, finished = false, drained = false;
  while (!finished) {
    for (long long k = 0; k < batch; k++, iter++) {
      if (use_graph) {
        PT_CUDA(cudaGraphLaunch(r->gexec[iter & 1], st));
        launches += (unsigned long long)r->g_launches;
      } else {
        launches += (unsigned long long)((iter & 1) ? iteration(B, A, timing) : iteration(A, B, timing));
      }
    }
    PT_CUDA(cudaGetLastError());
    PT_CUDA(cudaMemcpyAsync(&r->h_poll[slot], c, sizeof(CtrlPoll), cudaMemcpyDeviceToHost, st));
    PT_CUDA(cudaEventRecord(r->ev_poll[slot], st));
    if (outstanding) {
      PT_CUDA(cudaEventSynchronize(r->ev_poll[slot ^ 1]));
      // state at the BEGIN of the last iteration of the previous batch: n_cur paths were in flight then, so at
      // most that many are alive now if nothing is left to generate
      const CtrlPoll& q = r->h_poll[slot ^ 1];
      const unsigned long long remaining = q.total_samples - q.next_sample;
      if (remaining == 0 && q.n_cur <= (use_drain ? drain_max : 0)) {
        finished = true;
        drained = use_drain && q.n_cur > 0;
      }
      batch = even(std::max<long long>(4, (long long)(remaining / (unsigned)std::max(1, r->pool))));
    } else {
      batch = 4;
    }
    outstanding = true;
    slot ^= 1;
    if (iter > 100000000ll) {
      pt_set_error("pt_render: did not converge");
      cudaStreamSynchronize(st);
      return PT_ERR_INVALID;
    }
  }
  if (drained) {
    const int g_drain = (drain_max + 63) / 64;
    if (sphere_lights) k_drain<1><<<g_drain, 64, 0, st>>>(sc, p, c, A, count ? 1 : 0);
    else k_drain<0><<<g_drain, 64, 0, st>>>(sc, p, c, A, count ? 1 : 0);
    k_drain_end<<<1, 1, 0, st>>>(c);
    PT_CUDA(cudaGetLastError());
    launches += 2;
  }
  PT_CUDA(cudaMemcpyAsync(r->h_ctrl, c, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
  PT_CUDA(cudaEventRecord(r->ev1, st));
  PT_CUDA(cudaEventSynchronize(r->ev1));
  float ms = 0.f;
  PT_CUDA(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
  const Ctrl& hc = *r->h_ctrl;
  if (hc.error) {
    pt_set_error("pt_render: BVH traversal stack overflow (tree deeper than %d node groups)", PT_STACK);
    return PT_ERR_INVALID;
  }
  if (hc.next_sample < hc.total_samples || hc.n_next != 0) {
    pt_set_error("pt_render: internal error, %llu samples ungenerated and %d paths alive at the end",
                 hc.total_samples - hc.next_sample, hc.n_next);
    return PT_ERR_INVALID;
  }
  r->stats.samples = (uint64_t)total;
  r->stats.closest_rays = hc.stat_closest;
  r->stats.shadow_rays = hc.stat_shadow + (uint64_t)hc.n_shadow;  // the last iteration's count was never folded in
  r->stats.nodes_fetched = hc.stat_nodes;
  r->stats.leaves_fetched = hc.stat_leaves;
  r->stats.iterations = hc.stat_iters;
  r->stats.kernel_launches = launches;
  r->stats.ms_total = ms;
  r->stats.ms_generate = ms_gen;
  r->stats.ms_trace_closest = ms_trace;
  r->stats.ms_shade = ms_shade;
  r->stats.ms_trace_shadow = ms_shadow;
  r->stats.ms_reduce = 0.f;
  r->stats.drained_paths = drained ? 1u : 0u;
  return PT_OK;
}

extern "C" {

int pt_renderer_sync(pt_renderer* r) {
  if (!r) return PT_ERR_INVALID;
  if (r->multi) return pt_multi_sync(r->multi);
  DeviceGuard dg(r->device);
  PT_CUDA(cudaStreamSynchronize(r->stream));
  return PT_OK;
}

int pt_read_accum(pt_renderer* r, float* host_rgba) {
  if (!r || !host_rgba) return PT_ERR_INVALID;
  if (r->multi) {  // the per-device buffers are merged (NCCL sum-reduce) into the root's first
    int rc = pt_multi_merge(r->multi);
    return rc ? rc : pt_read_accum(pt_multi_root(r->multi), host_rgba);
  }
  DeviceGuard dg(r->device);
  const size_t bytes = (size_t)r->W * r->H * 16;
  // The caller's buffer is usually pageable, and a pageable device-to-host copy is staged by the driver at
  // ~2 GB/s (7.4 ms for the 15 MB of a 1200x800 image).  Copy into a pinned buffer of the renderer's own at
  // link speed (0.35 ms) and memcpy from there; a caller that passes pinned memory gets the direct copy.
  cudaPointerAttributes at{};
  const bool caller_pinned = cudaPointerGetAttributes(&at, host_rgba) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();  // unregistered host pointers make cudaPointerGetAttributes report an error on old drivers
  if (caller_pinned) {
    PT_CUDA(cudaMemcpyAsync(host_rgba, r->accum, bytes, cudaMemcpyDeviceToHost, r->stream));
    PT_CUDA(cudaStreamSynchronize(r->stream));
    return PT_OK;
  }
  if (r->stage_bytes < bytes) {  // grow-only, owned by the renderer
    if (r->stage) cudaFreeHost(r->stage);
    r->stage = nullptr;
    r->stage_bytes = 0;
    PT_CUDA(cudaMallocHost((void**)&r->stage, bytes));
    r->stage_bytes = bytes;
  }
  PT_CUDA(cudaMemcpyAsync(r->stage, r->accum, bytes, cudaMemcpyDeviceToHost, r->stream));
  PT_CUDA(cudaStreamSynchronize(r->stream));
  memcpy(host_rgba, r->stage, bytes);
  return PT_OK;
}

int pt_read_display(pt_renderer* r, uint8_t* host_rgba8) {
  if (!r || !host_rgba8) return PT_ERR_INVALID;
  if (r->multi) {
    int rc = pt_multi_merge(r->multi);
    return rc ? rc : pt_read_display(pt_multi_root(r->multi), host_rgba8);
  }
  DeviceGuard dg(r->device);
  const int n = r->W * r->H;
  if (!r->display.p) PT_CUDA(r->display.alloc((size_t)n * 4));
  k_display<<<sm_count() * 4, 256, 0, r->stream>>>(r->accum, r->display.as<uchar4>(), n);
  PT_CUDA(cudaGetLastError());
  PT_CUDA(cudaMemcpyAsync(host_rgba8, r->display.p, (size_t)n * 4, cudaMemcpyDeviceToHost, r->stream));
  PT_CUDA(cudaStreamSynchronize(r->stream));
  return PT_OK;
}

int pt_write_accum(pt_renderer* r, const float* host_rgba) {
  if (!r || !host_rgba) return PT_ERR_INVALID;
  if (r->multi) return pt_multi_write_accum(r->multi, host_rgba);
  DeviceGuard dg(r->device);
  PT_CUDA(cudaMemcpyAsync(r->accum, host_rgba, (size_t)r->W * r->H * 16, cudaMemcpyHostToDevice, r->stream));
  PT_CUDA(cudaStreamSynchronize(r->stream));
  r->started = true;  // a restored buffer must not be cleared by the next pt_render(first_frame = 0)
  return PT_OK;
}

int pt_accum_device_ptr(pt_renderer* r, void** d_ptr) {
  if (!r || !d_ptr) return PT_ERR_INVALID;
  if (r->multi) {  // the merged buffer on the root device
    int rc = pt_multi_merge(r->multi);
    if (rc) return rc;
    *d_ptr = pt_multi_root(r->multi)->accum;
    return PT_OK;
  }
  *d_ptr = r->accum;
  return PT_OK;
}

int pt_get_stats(pt_renderer* r, pt_stats* out) {
  if (!r || !out) return PT_ERR_INVALID;
  if (r->multi) return pt_multi_get_stats(r->multi, out);
  *out = r->stats;
  return PT_OK;
}

int pt_release_caches(void) {
  pt_release_slab_cache();
  pt_bvh_release_arenas();
  return PT_OK;
}

int pt_trace_rays_device(pt_scene* scene, const void* d_rays, int64_t n, void* d_hits, void* cuda_stream) {
  if (!scene || !scene->committed) {
    pt_set_error("pt_trace_rays: scene not committed");
    return PT_ERR_STATE;
  }
  if (n <= 0) return PT_OK;
  if (!d_rays || !d_hits) {
    pt_set_error("pt_trace_rays_device: null ray or hit buffer");
    return PT_ERR_INVALID;
  }
  if (n > 0x7fffffff) {
    pt_set_error("pt_trace_rays: at most 2^31-1 rays per call");
    return PT_ERR_INVALID;
  }
  DeviceGuard dg(scene->device);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  unsigned long long* dc = nullptr;
  PT_CUDA(cudaMallocAsync((void**)&dc, 32, st));
  PT_CUDA(cudaMemsetAsync(dc, 0, 32, st));
  const bool special = scene->dev.special_mask != 0;
  const SchedParams sched = sched_params_rays(scene->dev);
  if (rays_deep(scene->dev, sched))
    PT_LAUNCH_TRACE(k_trace_rays_deep, false, special, sm_count() * trace_blocks_per_sm(special), st, scene->dev,
                    (const pt_ray*)d_rays, (int)n, (pt_hit*)d_hits, dc, sched);
  else
    PT_LAUNCH_TRACE(k_trace_rays, false, special, sm_count() * trace_blocks_per_sm(special), st, scene->dev,
                    (const pt_ray*)d_rays, (int)n, (pt_hit*)d_hits, dc, sched);
  PT_CUDA(cudaGetLastError());
  PT_CUDA(cudaFreeAsync(dc, st));
  return PT_OK;
}

int pt_trace_rays(pt_scene* scene, const pt_ray* rays, int64_t n, pt_hit* hits, uint64_t* counters) {
  if (!scene || !scene->committed) {
    pt_set_error("pt_trace_rays: scene not committed");
    return PT_ERR_STATE;
  }
  if (n <= 0) return PT_OK;
  if (!rays || !hits) return PT_ERR_INVALID;
  if (n > 0x7fffffff) {
    pt_set_error("pt_trace_rays: at most 2^31-1 rays per call");
    return PT_ERR_INVALID;
  }
  DeviceGuard dg(scene->device);
  DevBuf dr, dh, dc;
  PT_CUDA(dr.alloc((size_t)n * sizeof(pt_ray)));
  PT_CUDA(dh.alloc((size_t)n * sizeof(pt_hit)));
  PT_CUDA(dc.alloc(32));
  PT_CUDA(cudaMemset(dc.p, 0, 32));
  PT_CUDA(cudaMemcpy(dr.p, rays, (size_t)n * sizeof(pt_ray), cudaMemcpyHostToDevice));
  const bool special = scene->dev.special_mask != 0;
  const int grid = sm_count() * trace_blocks_per_sm(special);
  const SchedParams sched = sched_params_rays(scene->dev);
  if (rays_deep(scene->dev, sched))
    PT_LAUNCH_TRACE(k_trace_rays_deep, counters != nullptr, special, grid, (cudaStream_t)0, scene->dev, dr.as<pt_ray>(), (int)n,
                    dh.as<pt_hit>(), dc.as<unsigned long long>(), sched);
  else
    PT_LAUNCH_TRACE(k_trace_rays, counters != nullptr, special, grid, (cudaStream_t)0, scene->dev, dr.as<pt_ray>(), (int)n,
                    dh.as<pt_hit>(), dc.as<unsigned long long>(), sched);
  PT_CUDA(cudaGetLastError());
  PT_CUDA(cudaMemcpy(hits, dh.p, (size_t)n * sizeof(pt_hit), cudaMemcpyDeviceToHost));
  unsigned long long hc[4];
  PT_CUDA(cudaMemcpy(hc, dc.p, 32, cudaMemcpyDeviceToHost));
  if (counters) {
    counters[0] = hc[0];
    counters[1] = hc[1];
  }
  if (hc[3]) {
    pt_set_error("pt_trace_rays: BVH traversal stack overflow (tree deeper than %d node groups)", PT_STACK);
    return PT_ERR_INVALID;
  }
  return PT_OK;
}

}  // extern "C"
