// bvh_build.cu — GPU BVH builder driver: LBVH (Morton + radix sort + Karras
// hierarchy) -> bottom-up refit -> binned-SAH treelet refinement -> collapse
// to the compressed 8-wide BVH.  The per-thread bodies live in bvh_core.h.
//
// Replaces the OptiX "Trbvh" builds the reference requests
// (host_includes/scenes.hpp:205 etc., "OptiX Building Time" main.cpp:99).
#include <cub/cub.cuh>

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scene_host.h"

namespace {

constexpr int TB = 256;
inline int nblk(long long n, int tb = TB) { return (int)((n + tb - 1) / tb); }

__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  // valid for any sign: order-preserving int mapping
  if (v >= 0.f) atomicMin((int*)addr, __float_as_int(v));
  else atomicMax((unsigned*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v));
  else atomicMin((unsigned*)addr, __float_as_uint(v));
}

// scene_bounds: [0..2] centroid min, [3..5] centroid max, [6] largest |coordinate|
__global__ void k_prim_bounds(BuildIn in, float4* plo, float4* phi, float* scene_bounds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float3 c = f3(0.f);
  float amax = 0.f;
  bool ok = i < in.n;
  if (ok) {
    float3 lo, hi;
    prim_bounds(in, i, lo, hi);
    plo[i] = make_float4(lo.x, lo.y, lo.z, 0.f);
    phi[i] = make_float4(hi.x, hi.y, hi.z, 0.f);
    c = (lo + hi) * 0.5f;
    amax = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(lo.y)), fabsf(lo.z)), fmaxf(fmaxf(fabsf(hi.x), fabsf(hi.y)), fabsf(hi.z)));
    if (!(amax < 3.0e38f)) amax = 0.f;
  }
  // warp reduce then one atomic per warp
  float mn[3] = {ok ? c.x : FLT_MAX, ok ? c.y : FLT_MAX, ok ? c.z : FLT_MAX};
  float mx[3] = {ok ? c.x : -FLT_MAX, ok ? c.y : -FLT_MAX, ok ? c.z : -FLT_MAX};
  for (int o = 16; o > 0; o >>= 1)
    for (int a = 0; a < 3; a++) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) {
    for (int a = 0; a < 3; a++) {
      if (mn[a] != FLT_MAX) atomic_min_f(scene_bounds + a, mn[a]);
      if (mx[a] != -FLT_MAX) atomic_max_f(scene_bounds + 3 + a, mx[a]);
    }
    atomic_max_f(scene_bounds + 6, amax);
  }
}

__global__ void k_pad_bounds(int n, float4* plo, float4* phi, const float* scene_bounds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float3 lo = xyz(plo[i]), hi = xyz(phi[i]);
  pad_bounds(lo, hi, scene_bounds[6]);
  plo[i] = make_float4(lo.x, lo.y, lo.z, 0.f);
  phi[i] = make_float4(hi.x, hi.y, hi.z, 0.f);
}

// ---- top-down binned-SAH partition of the upper levels (bodies and rationale: bvh_core.h 2b) -------------
__global__ void k_td_root(TdNode* nodes, const float* sb, int* node_id, uint32_t* code, int* clen, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    TdNode r;
    for (int a = 0; a < 3; a++) {
      r.cmin[a] = sb[a];
      r.cmax[a] = sb[3 + a];
    }
    r.axis = -1;
    r.bin = 0;
    r.child = -1;
    r.pad = 0;
    nodes[0] = r;
  }
  if (i < n) {
    node_id[i] = 0;
    code[i] = 0u;
    clen[i] = 0;
  }
}
__global__ void k_td_clear(TdBin* bins, const int* n_nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_nodes * 3 * PT_TD_BINS) return;
  TdBin b;
  b.count = 0;
  for (int a = 0; a < 3; a++) {
    b.lo[a] = FLT_MAX;
    b.hi[a] = -FLT_MAX;
  }
  bins[i] = b;
}
// order-preserving float <-> uint mapping (warp reductions work on integers)
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
__device__ __forceinline__ void td_bin_merge(TdBin* b, int count, float3 lo, float3 hi) {
  atomicAdd(&b->count, count);
  atomic_min_f(&b->lo[0], lo.x);
  atomic_min_f(&b->lo[1], lo.y);
  atomic_min_f(&b->lo[2], lo.z);
  atomic_max_f(&b->hi[0], hi.x);
  atomic_max_f(&b->hi[1], hi.y);
  atomic_max_f(&b->hi[2], hi.z);
}
// Binning.  Every primitive adds its box to one bin per axis of its node: 21 atomics per primitive and level if done
// naively, with a million primitives aiming at the 48 bins of the root.  Two remedies:
//   SMEM   (level has <= PT_TD_SMEM_NODES nodes) bins are accumulated per block in shared memory and merged once;
//   else   lanes of a warp that aim at the same bin (meshes arrive spatially ordered) are reduced with
//          match.any + redux first, and one lane issues the atomics.
#define PT_TD_SMEM_NODES 16
template <bool SMEM>
__global__ void __launch_bounds__(256) k_td_bin(int n, const float4* __restrict__ plo, const float4* __restrict__ phi,
                                                const int* __restrict__ node_id, const TdNode* __restrict__ nodes,
                                                const int* n_nodes, TdBin* bins) {
  __shared__ TdBin sb[SMEM ? PT_TD_SMEM_NODES * 3 * PT_TD_BINS : 1];
  const int nb = SMEM ? *n_nodes * 3 * PT_TD_BINS : 0;
  if (SMEM) {
    for (int k = threadIdx.x; k < nb; k += blockDim.x) {
      sb[k].count = 0;
      for (int a = 0; a < 3; a++) {
        sb[k].lo[a] = FLT_MAX;
        sb[k].hi[a] = -FLT_MAX;
      }
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {  // warp-uniform trip count
    const int i = base + lane;
    const int id = i < n ? node_id[i] : -1;
    float3 lo = f3(0.f), hi = f3(0.f), c = f3(0.f);
    TdNode nd;
    if (id >= 0) {
      nd = nodes[id];
      lo = xyz(plo[i]);
      hi = xyz(phi[i]);
      c = (lo + hi) * 0.5f;
    }
    for (int a = 0; a < 3; a++) {
      const int slot = id >= 0 ? (id * 3 + a) * PT_TD_BINS + td_bin_of(comp(c, a), nd.cmin[a], nd.cmax[a]) : -1;
      if (SMEM) {
        if (slot >= 0) td_bin_merge(sb + slot, 1, lo, hi);
      } else {
        const unsigned peers = __match_any_sync(0xffffffffu, slot);
        if (slot >= 0) {
          const int cnt = __popc(peers);
          float3 l, h;
          l.x = ord2f(__reduce_min_sync(peers, f2ord(lo.x)));
          l.y = ord2f(__reduce_min_sync(peers, f2ord(lo.y)));
          l.z = ord2f(__reduce_min_sync(peers, f2ord(lo.z)));
          h.x = ord2f(__reduce_max_sync(peers, f2ord(hi.x)));
          h.y = ord2f(__reduce_max_sync(peers, f2ord(hi.y)));
          h.z = ord2f(__reduce_max_sync(peers, f2ord(hi.z)));
          if (lane == __ffs(peers) - 1) td_bin_merge(bins + slot, cnt, l, h);
        }
      }
    }
  }
  if (SMEM) {
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x)
      if (sb[k].count) td_bin_merge(bins + k, sb[k].count, f3(sb[k].lo[0], sb[k].lo[1], sb[k].lo[2]), f3(sb[k].hi[0], sb[k].hi[1], sb[k].hi[2]));
  }
}
__global__ void k_td_split(TdNode* nodes, const int* n_nodes, const TdBin* bins, int leaf_size, int depth, TdNode* next,
                           int* n_next) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= *n_nodes) return;
  TdNode nd = nodes[id];
  td_choose(bins + (size_t)id * 3 * PT_TD_BINS, nd, leaf_size, depth);
  if (nd.axis >= 0) {
    nd.child = atomicAdd(n_next, 2);
    TdNode ch;
    for (int a = 0; a < 3; a++) {
      ch.cmin[a] = FLT_MAX;
      ch.cmax[a] = -FLT_MAX;
    }
    ch.axis = -1;
    ch.bin = 0;
    ch.child = -1;
    ch.pad = 0;
    next[nd.child] = ch;
    next[nd.child + 1] = ch;
  }
  nodes[id] = nd;
}
__global__ void __launch_bounds__(256) k_td_assign(int n, const float4* __restrict__ plo, const float4* __restrict__ phi, int* node_id,
                                                   uint32_t* code, int* clen, const TdNode* __restrict__ nodes, TdNode* next) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // the grid covers n rounded up to whole warps
  int ch = -1;
  float3 c = f3(0.f);
  const int id = i < n ? node_id[i] : -1;
  if (id >= 0) {
    const TdNode nd = nodes[id];
    if (nd.axis < 0) {  // finished group: the path code stays as it is
      node_id[i] = -1;
    } else {
      c = (xyz(plo[i]) + xyz(phi[i])) * 0.5f;
      const int side = td_bin_of(comp(c, nd.axis), nd.cmin[nd.axis], nd.cmax[nd.axis]) <= nd.bin ? 0 : 1;
      ch = nd.child + side;
      node_id[i] = ch;
      code[i] = (code[i] << 1) | (uint32_t)side;
      clen[i] += 1;
    }
  }
  // centroid bounds of the children: lanes with the same child reduce first, one lane issues the six atomics
  const unsigned peers = __match_any_sync(0xffffffffu, ch);
  if (ch >= 0) {
    const float lx = ord2f(__reduce_min_sync(peers, f2ord(c.x))), ly = ord2f(__reduce_min_sync(peers, f2ord(c.y)));
    const float lz = ord2f(__reduce_min_sync(peers, f2ord(c.z))), hx = ord2f(__reduce_max_sync(peers, f2ord(c.x)));
    const float hy = ord2f(__reduce_max_sync(peers, f2ord(c.y))), hz = ord2f(__reduce_max_sync(peers, f2ord(c.z)));
    if (lane == __ffs(peers) - 1) {
      atomic_min_f(&next[ch].cmin[0], lx);
      atomic_min_f(&next[ch].cmin[1], ly);
      atomic_min_f(&next[ch].cmin[2], lz);
      atomic_max_f(&next[ch].cmax[0], hx);
      atomic_max_f(&next[ch].cmax[1], hy);
      atomic_max_f(&next[ch].cmax[2], hz);
    }
  }
}

__global__ void k_morton(int n, const float4* plo, const float4* phi, const float* sb, const uint32_t* code, const int* clen,
                         uint64_t* keys, uint32_t* vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float3 smin = f3(sb[0], sb[1], sb[2]);
  float3 ext = f3(sb[3], sb[4], sb[5]) - smin;
  float3 sinv = f3(ext.x > 0.f ? 1.f / ext.x : 0.f, ext.y > 0.f ? 1.f / ext.y : 0.f, ext.z > 0.f ? 1.f / ext.z : 0.f);
  float3 c = (xyz(plo[i]) + xyz(phi[i])) * 0.5f;
  keys[i] = td_compose_key(morton63(c, smin, sinv), code ? code[i] : 0u, code ? clen[i] : 0);
  vals[i] = (uint32_t)i;
}

__global__ void k_karras(Bvh2 t) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < t.n - 1) karras_node(t, i);
}

__device__ __forceinline__ void child_box_cg(const Bvh2& t, int ref, float3& lo, float3& hi, int& cnt) {
  if (ref >= 0) {  // written by another thread in this kernel: bypass L1
    const float4* p = (const float4*)(t.nodes + ref);
    float4 a = __ldcg(p), b = __ldcg(p + 1);
    lo = f3(a.x, a.y, a.z);
    hi = f3(b.x, b.y, b.z);
    cnt = __ldcg(t.count + ref);
  } else {
    uint32_t id = t.sorted[~ref];
    lo = xyz(t.plo[id]);
    hi = xyz(t.phi[id]);
    cnt = 1;
  }
}

__global__ void k_refit(Bvh2 t) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= t.n) return;
  int cur = t.leaf_parent[j];
  while (cur >= 0) {
    __threadfence();
    int old = atomicAdd(t.flag + cur, 1);
    if (old == 0) return;  // sibling subtree not finished yet
    Bvh2Node nd = t.nodes[cur];
    float3 llo, lhi, rlo, rhi;
    int lc, rc;
    child_box_cg(t, nd.left, llo, lhi, lc);
    child_box_cg(t, nd.right, rlo, rhi, rc);
    bvh2_set_box(nd, fmin3(llo, rlo), fmax3(lhi, rhi));
    t.nodes[cur] = nd;
    t.count[cur] = lc + rc;
    cur = t.parent[cur];
  }
}

__global__ void k_treelet_mark(Bvh2 t, int lo, long long hi, int* roots, int* n_roots) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.n - 1) return;
  if (treelet_is_root(t, i, lo, hi)) roots[atomicAdd(n_roots, 1)] = i;
}
// the number of roots stays on the device: the grid covers the upper bound, surplus threads leave
__global__ void k_treelet_sah(Bvh2 t, int lo, const int* roots, const int* n_roots) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *n_roots) treelet_sah(t, roots[i], lo);
}

// SAH cost: sum of inner-node areas (C_trav = 1) + leaf areas (C_isect = 1), over root area
__global__ void k_sah_cost(Bvh2 t, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double a = 0.0;
  if (i < t.n - 1) {
    const Bvh2Node& nd = t.nodes[i];
    a = (double)box_half_area(f3(nd.lo[0], nd.lo[1], nd.lo[2]), f3(nd.hi[0], nd.hi[1], nd.hi[2]));
  }
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0 && a != 0.0) atomicAdd(out, a);
}

// SAH-optimal collapse decisions (bvh_core.h 6a), bottom-up like the refit: the second thread to arrive at a node
// finds both children's cost tables complete
__global__ void k_collapse_dp(Bvh2 t, float* cost, unsigned char* dec) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= t.n) return;
  int cur = t.leaf_parent[j];
  while (cur >= 0) {
    __threadfence();
    int old = atomicAdd(t.flag + cur, 1);
    if (old == 0) return;
    const Bvh2Node nd = t.nodes[cur];
    float cl[7], cr[7];
    for (int side = 0; side < 2; side++) {
      const int rf = side ? nd.right : nd.left;
      float* dst = side ? cr : cl;
      if (rf >= 0) {
        const float4* p = (const float4*)(cost + (size_t)rf * 8);
        const float4 a = __ldcg(p), b = __ldcg(p + 1);
        dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
        dst[4] = b.x; dst[5] = b.y; dst[6] = b.z;
      } else {
        const uint32_t id = t.sorted[~rf];
        const float a = box_half_area(xyz(t.plo[id]), xyz(t.phi[id])) * PT_DP_CPRIM;
        for (int i = 0; i < 7; i++) dst[i] = a;
      }
    }
    float c8[8];
    unsigned char d9[9];
    collapse_dp_node(box_half_area(f3(nd.lo[0], nd.lo[1], nd.lo[2]), f3(nd.hi[0], nd.hi[1], nd.hi[2])), t.count[cur], cl, cr, c8, d9);
    c8[7] = 0.f;
    float4* q = (float4*)(cost + (size_t)cur * 8);
    q[0] = make_float4(c8[0], c8[1], c8[2], c8[3]);
    q[1] = make_float4(c8[4], c8[5], c8[6], c8[7]);
    for (int i = 0; i < 9; i++) dec[(size_t)cur * 9 + i] = d9[i];
    cur = t.parent[cur];
  }
}

// one level of the collapse; level counts live on the device (level_count[l] = work items of level l)
__global__ void k_collapse(Collapse c, const uint2* work, const int* n_work, uint2* out_work, int* out_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_work) return;
  uint2 w = work[i];
  collapse_node(c, (int)w.x, (int)w.y, out_work, out_count);
}

// Grow-only scratch arena of the build, one per device, kept for the life of the process: a 10 M-triangle build
// takes ~1.3 GB of scratch in ~20 pieces, and allocating them afresh per commit made build time depend on what the
// allocator had to map (collapse 7 ms or 74 ms between two runs of the same scene, round-1 finding).
// pt_release_caches() returns it.
struct Arena {
  DevBuf buf;
  size_t off = 0;
  void* take(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    void* p = (char*)buf.p + off;
    off += bytes;
    return p;
  }
};
Arena g_arena[PT_MAX_DEVICES];

}  // namespace

namespace {
struct StageTimer {  // PTB200_BUILD_TIMING=1 prints per-stage GPU time of the build
  bool on;
  cudaEvent_t last{};
  explicit StageTimer() : on(getenv("PTB200_BUILD_TIMING") != nullptr) {
    if (on) {
      cudaEventCreate(&last);
      cudaEventRecord(last);
    }
  }
  void mark(const char* name) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, last, e);
    fprintf(stderr, "[bvh build] %-16s %8.3f ms\n", name, ms);
    cudaEventDestroy(last);
    last = e;
  }
};
}  // namespace

void pt_bvh_release_arenas() {
  for (auto& a : g_arena) a.buf.release();
}

int pt_bvh_build(pt_scene* s) {
  const int n = (int)s->prims.size();
  StageTimer tm;
  if (n <= 0) {
    pt_set_error("scene has no primitives");
    return PT_ERR_INVALID;
  }
  int dev = 0;
  PT_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= PT_MAX_DEVICES) {
    pt_set_error("device ordinal %d outside the supported range", dev);
    return PT_ERR_INVALID;
  }
  cudaEvent_t ev0, ev1;
  PT_CUDA(cudaEventCreate(&ev0));
  PT_CUDA(cudaEventCreate(&ev1));
  PT_CUDA(cudaEventRecord(ev0));

  BuildIn in;
  in.prim_ref = s->d_prim_ref.as<pt_prim_ref>();
  in.spheres = s->d_spheres.as<pt_sphere>();
  in.rects = s->d_rects.as<pt_rect>();
  in.boxes = s->d_boxes.as<pt_box>();
  in.moving_spheres = s->d_moving_spheres.as<pt_moving_sphere>();
  in.volume_spheres = s->d_volume_spheres.as<pt_volume_sphere>();
  in.volume_boxes = s->d_volume_boxes.as<pt_volume_box>();
  in.cam_time0 = s->camera.time0;
  in.cam_time1 = s->camera.time1;
  in.tri_v = s->d_tri_v.as<float>();
  in.xforms = s->d_xforms.as<pt_xform>();
  in.materials = s->d_materials.as<pt_material>();
  in.n = n;

  // ---- scratch: one arena, carved once; nothing below allocates or talks to the host until the end ----
  const int ni = n > 1 ? n - 1 : 1;
  size_t cub_bytes = 0;
  PT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                          (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, 63));
  const size_t n8 = (size_t)(n > 1 ? n : 2);
  // nodes alive on one level of the top-down partition: two children per node that still holds > leaf_size primitives
  const int td_leaf = td_leaf_size(n);
  const size_t td_cap = std::min<size_t>(PT_TD_MAXNODES, 2 * ((size_t)n / td_leaf) + 2) + 2;
  const size_t td_bytes = (size_t)n * 12 + 2 * td_cap * sizeof(TdNode) + td_cap * 3 * PT_TD_BINS * sizeof(TdBin) + 4096;
  const size_t dp_bytes = (size_t)ni * (32 + 9) + 512;
  const size_t need = td_bytes + dp_bytes + (size_t)n * (16 + 16 + 8 + 8 + 4 + 4 + 4 + 8 + 8) + (size_t)ni * (sizeof(Bvh2Node) + 4 + 4 + 4 + 4) +
                      n8 * 80 + (size_t)n * sizeof(LeafRec) + cub_bytes + 64 * 256 + 4096;
  Arena& ar = g_arena[dev];
  if (ar.buf.bytes < need) PT_CUDA(ar.buf.alloc(need + need / 8));
  ar.off = 0;
  float4* plo = (float4*)ar.take((size_t)n * 16);
  float4* phi = (float4*)ar.take((size_t)n * 16);
  uint64_t* keys = (uint64_t*)ar.take((size_t)n * 8);
  uint64_t* keys2 = (uint64_t*)ar.take((size_t)n * 8);
  uint32_t* vals = (uint32_t*)ar.take((size_t)n * 4);
  uint32_t* vals2 = (uint32_t*)ar.take((size_t)n * 4);
  void* cubtmp = ar.take(cub_bytes);
  Bvh2Node* nodes2 = (Bvh2Node*)ar.take((size_t)ni * sizeof(Bvh2Node));
  int* parent = (int*)ar.take((size_t)ni * 4);
  int* leaf_parent = (int*)ar.take((size_t)n * 4);
  int* count = (int*)ar.take((size_t)ni * 4);
  int* flag = (int*)ar.take((size_t)ni * 4);
  int* roots = (int*)ar.take((size_t)ni * 4);
  uint2* wa = (uint2*)ar.take((size_t)n * 8 + 64);
  uint2* wb = (uint2*)ar.take((size_t)n * 8 + 64);
  uint4* nodes8_tmp = (uint4*)ar.take(n8 * 80);
  float* dp_cost = (float*)ar.take((size_t)ni * 32);
  unsigned char* dp_dec = (unsigned char*)ar.take((size_t)ni * 9);
  int* td_node_id = (int*)ar.take((size_t)n * 4);
  uint32_t* td_code = (uint32_t*)ar.take((size_t)n * 4);
  int* td_clen = (int*)ar.take((size_t)n * 4);
  TdNode* td_nodes[2] = {(TdNode*)ar.take(td_cap * sizeof(TdNode)), (TdNode*)ar.take(td_cap * sizeof(TdNode))};
  TdBin* td_bins = (TdBin*)ar.take(td_cap * 3 * PT_TD_BINS * sizeof(TdBin));
  LeafRec* leaves_tmp = (LeafRec*)ar.take((size_t)n * sizeof(LeafRec));
  // small device-resident state, zeroed once:
  //   [0..7] scene bounds (floats)   [16] BVH8 node counter   [17] leaf counter
  //   [32..32+MAXLEV] collapse work count per level   [200..200+MAXPASS) treelet root counts per (round, pass)
  //   [300..300+PT_TD_MAXDEPTH] active nodes per level of the top-down partition
  //   doubles at byte 1024: SAH sums
  constexpr int MAXLEV = 160, MAXPASS = 48;
  int* st = (int*)ar.take(2048);
  float* sbounds = (float*)st;
  int* pass_roots = st + 200;
  int* td_count = st + 300;
  int* d_cnt = st + 16;
  int* level_count = st + 32;
  double* cost = (double*)((char*)st + 1024);
  {
    int init[512];
    memset(init, 0, sizeof(init));
    const float fb[8] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f, 0.f};
    memcpy(init, fb, sizeof(fb));
    init[16] = 1;        // node 0 is the root
    init[32] = 1;        // level 0 holds one work item
    init[300] = 1;       // the top-down partition starts with the root
    PT_CUDA(cudaMemcpyAsync(st, init, sizeof(init), cudaMemcpyHostToDevice, 0));
  }

  k_prim_bounds<<<nblk(n), TB>>>(in, plo, phi, sbounds);
  PT_CUDA(cudaGetLastError());
  k_pad_bounds<<<nblk(n), TB>>>(n, plo, phi, sbounds);
  PT_CUDA(cudaGetLastError());
  tm.mark("bounds");
  // top-down binned-SAH partition of the upper levels: path bits per primitive (bvh_core.h 2b)
  static const int td_on = getenv("PTB200_NO_TOPDOWN") ? 0 : 1;
  static const int td_leaf_env = getenv("PTB200_TD_LEAF") ? atoi(getenv("PTB200_TD_LEAF")) : 0;  // dev
  const int leaf_size = std::max(td_leaf, td_leaf_env);
  const bool topdown = td_on && n > leaf_size;
  if (topdown) {
    k_td_root<<<nblk(n), TB>>>(td_nodes[0], sbounds, td_node_id, td_code, td_clen, n);
    PT_CUDA(cudaGetLastError());
    for (int level = 0; level < PT_TD_MAXDEPTH; level++) {
      TdNode* cur = td_nodes[level & 1];
      TdNode* nxt = td_nodes[(level + 1) & 1];
      const long long bound = std::min<long long>((long long)td_cap - 2, level < 30 ? (1ll << level) : (long long)td_cap);
      k_td_clear<<<nblk(bound * 3 * PT_TD_BINS), TB>>>(td_bins, td_count + level);
      if (bound <= PT_TD_SMEM_NODES) k_td_bin<true><<<std::min(nblk(n), 148 * 4), TB>>>(n, plo, phi, td_node_id, cur, td_count + level, td_bins);
      else k_td_bin<false><<<std::min(nblk(n), 148 * 16), TB>>>(n, plo, phi, td_node_id, cur, td_count + level, td_bins);
      k_td_split<<<nblk(bound, 64), 64>>>(cur, td_count + level, td_bins, leaf_size, level, nxt, td_count + level + 1);
      k_td_assign<<<nblk(n), TB>>>(n, plo, phi, td_node_id, td_code, td_clen, cur, nxt);
      PT_CUDA(cudaGetLastError());
      if ((level & 3) == 3) {  // a look at the counters every four levels: stop when no node is left to split
        int alive = 0;
        PT_CUDA(cudaMemcpy(&alive, td_count + level + 1, 4, cudaMemcpyDeviceToHost));
        if (alive == 0) break;
      }
    }
    tm.mark("top-down sah");
  }
  k_morton<<<nblk(n), TB>>>(n, plo, phi, sbounds, topdown ? td_code : nullptr, td_clen, keys, vals);
  PT_CUDA(cudaGetLastError());
  PT_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp, cub_bytes, keys, keys2, vals, vals2, n, 0, 63));
  tm.mark("morton+sort");

  PT_CUDA(cudaMemsetAsync(flag, 0, (size_t)ni * 4, 0));
  PT_CUDA(cudaMemsetAsync(leaf_parent, 0xff, (size_t)n * 4, 0));
  PT_CUDA(cudaMemsetAsync(parent, 0xff, (size_t)ni * 4, 0));
  PT_CUDA(cudaMemsetAsync(count, 0, (size_t)ni * 4, 0));

  Bvh2 t;
  t.nodes = nodes2;
  t.parent = parent;
  t.leaf_parent = leaf_parent;
  t.count = count;
  t.flag = flag;
  t.keys = keys2;
  t.sorted = vals2;
  t.plo = plo;
  t.phi = phi;
  t.n = n;

  if (n > 1) {
    k_karras<<<nblk(n - 1), TB>>>(t);
    PT_CUDA(cudaGetLastError());
    tm.mark("karras");
    k_refit<<<nblk(n), TB>>>(t);
    PT_CUDA(cudaGetLastError());
    tm.mark("refit");
    k_sah_cost<<<nblk(n - 1), TB>>>(t, cost);
    // hierarchical passes: leaves -> treelets of <= T leaves -> treelets of <= T treelets -> ..., one such
    // sweep per entry of PT_TREELET_ROUNDS.  A pass-k root has more than lo_k leaves and roots are disjoint
    // subtrees, so there are at most n / lo_k of them: the grid is sized for that bound and the actual count
    // never leaves the device.
    static const int rounds[] = PT_TREELET_ROUNDS;
    int slot = 0;
    for (int T : rounds) {
      long long lo = 1, hi = T;
      for (int pass = 0; pass < 12 && lo < (long long)n && slot < MAXPASS; pass++, slot++) {
        k_treelet_mark<<<nblk(n - 1), TB>>>(t, (int)lo, hi, roots, pass_roots + slot);
        PT_CUDA(cudaGetLastError());
        const long long bound = std::min<long long>(n - 1, (long long)n / lo + 1);
        k_treelet_sah<<<nblk(bound, 64), 64>>>(t, (int)lo, roots, pass_roots + slot);
        PT_CUDA(cudaGetLastError());
        lo = hi;
        hi *= T;
      }
    }
    tm.mark("treelet sah");
    k_sah_cost<<<nblk(n - 1), TB>>>(t, cost + 1);
  }

  // collapse: level-synchronous work lists; the per-level counts stay on the device.  An inner BVH8 child holds
  // more than 3 primitives, so a level has at most min(8^l, n / 4 + 1) items: launch for that bound.  Levels are
  // launched in batches of 16 and the host looks at the counters once per batch (usually once in total).
  PT_CUDA(s->d_prim_leaf.alloc((size_t)n * 4));
  Collapse c;
  c.t = t;
  c.in = in;
  c.nodes8 = nodes8_tmp;
  c.leaves = leaves_tmp;
  c.prim_leaf = s->d_prim_leaf.as<int>();
  // SAH-optimal choice of every wide node's children instead of the greedy expansion (PTB200_NO_DP=1: greedy, dev)
  static const int dp_on = getenv("PTB200_NO_DP") ? 0 : 1;
  c.dp = nullptr;
  if (dp_on && n > 1) {
    PT_CUDA(cudaMemsetAsync(flag, 0, (size_t)ni * 4, 0));
    k_collapse_dp<<<nblk(n), TB>>>(t, dp_cost, dp_dec);
    PT_CUDA(cudaGetLastError());
    c.dp = dp_dec;
    tm.mark("collapse dp");
  }
  c.node_counter = d_cnt;
  c.leaf_counter = d_cnt + 1;
  const uint2 first = make_uint2(n > 1 ? 0u : (uint32_t)~0, 0u);  // n == 1: leaf ref ~0
  PT_CUDA(cudaMemcpyAsync(wa, &first, sizeof(first), cudaMemcpyHostToDevice, 0));
  int level = 0;
  int h_state[2 + MAXLEV + 32];
  for (;;) {
    for (int k = 0; k < 16 && level < MAXLEV; k++, level++) {
      long long bound = (long long)n / 4 + 1;
      if (level < 10) bound = std::min<long long>(bound, 1ll << (3 * level));
      k_collapse<<<nblk(bound, 128), 128>>>(c, wa, level_count + level, wb, level_count + level + 1);
      PT_CUDA(cudaGetLastError());
      std::swap(wa, wb);
    }
    PT_CUDA(cudaMemcpy(h_state, d_cnt, sizeof(int) * (16 + level + 2), cudaMemcpyDeviceToHost));
    if (h_state[16 + level] == 0) break;  // the last level launched produced no work
    if (level >= MAXLEV) {
      pt_set_error("BVH collapse did not terminate");
      return PT_ERR_INVALID;
    }
  }
  tm.mark("collapse");
  const int hcnt[2] = {h_state[0], h_state[1]};
  if (hcnt[1] != n) {
    pt_set_error("BVH collapse wrote %d leaf records for %d primitives", hcnt[1], n);
    return PT_ERR_INVALID;
  }
  // nodes (at their final size) and leaf records into one block, nodes first
  s->bvh_nodes_bytes = (size_t)hcnt[0] * 80;
  s->bvh_leaves_off = (s->bvh_nodes_bytes + 255) & ~(size_t)255;
  PT_CUDA(s->d_bvh.alloc(s->bvh_leaves_off + (size_t)n * sizeof(LeafRec)));
  PT_CUDA(cudaMemcpyAsync(s->d_bvh.p, nodes8_tmp, s->bvh_nodes_bytes, cudaMemcpyDeviceToDevice, 0));
  PT_CUDA(cudaMemcpyAsync((char*)s->d_bvh.p + s->bvh_leaves_off, leaves_tmp, (size_t)n * sizeof(LeafRec),
                          cudaMemcpyDeviceToDevice, 0));
  double cost_lbvh = 0.0, cost_ref = 0.0;
  if (n > 1) {
    double hc[2];
    Bvh2Node root;
    PT_CUDA(cudaMemcpy(hc, cost, 16, cudaMemcpyDeviceToHost));
    PT_CUDA(cudaMemcpy(&root, nodes2, sizeof(root), cudaMemcpyDeviceToHost));
    double ra = (double)box_half_area(f3(root.lo[0], root.lo[1], root.lo[2]), f3(root.hi[0], root.hi[1], root.hi[2]));
    if (ra > 0.0) {
      cost_lbvh = hc[0] / ra;
      cost_ref = hc[1] / ra;
    }
  }
  PT_CUDA(cudaEventRecord(ev1));
  PT_CUDA(cudaEventSynchronize(ev1));
  float ms = 0.f;
  PT_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  s->info.n_prims = n;
  s->info.n_bvh2_nodes = n > 1 ? n - 1 : 0;
  s->info.n_bvh8_nodes = hcnt[0];
  s->info.n_leaf_records = hcnt[1];
  s->info.sah_cost_lbvh = (float)cost_lbvh;
  s->info.sah_cost_refined = (float)cost_ref;
  s->info.build_ms = ms;
  s->info.node_bytes = (int64_t)hcnt[0] * 80;
  s->info.leaf_bytes = (int64_t)hcnt[1] * (int64_t)sizeof(LeafRec);
  return PT_OK;
}
