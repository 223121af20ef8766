// bvh_core.h — per-thread bodies of the GPU BVH builder (PT_HD so that
// tests/hostsim can run the same code serially on the CPU).
//
// Replaces OptiX's closed "Trbvh" acceleration build that the reference
// requests at host_includes/scenes.hpp:32,112,205,278,393 and the
// Get_Bounds programs (programs/hitables/*.cu).  Pipeline (bvh_build.cu):
//   1 prim_bounds      world AABB per primitive (+padding), scene bounds
//   2 morton63         63-bit Morton code of the AABB centre, radix sort
//   3 karras_node      LBVH hierarchy (Karras 2012), one thread / inner node
//   4 refit_leaf       bottom-up AABB + leaf-count refit (atomic arrival flags)
//   5 treelet_sah      binned-SAH rebuild of every maximal subtree with
//                      <= PT_TREELET leaves, in place (subtrees are contiguous
//                      in Morton order, so node storage is reused)
//   6 collapse_node    BVH2 -> 8-wide: greedy surface-area expansion, octant
//                      slot assignment, child-box quantisation, leaf records
#pragma once
#include "scene_dev.h"

#ifndef PT_TREELET
#define PT_TREELET 64
#endif
#ifndef PT_SAH_BINS
#define PT_SAH_BINS 8
#endif
// Rounds of the hierarchical treelet refinement, each with its own treelet size (<= PT_TREELET).  One round fixes
// the grouping of leaves into treelets by the tree it starts from; a further round with another size regroups
// across the previous round's treelet boundaries.  On a plain LBVH that pays (1 M-triangle mesh, SAH cost: LBVH
// 54.5 -> {64} 45.8 -> {64, 32, 48} 40.2); on top of the top-down partition of the upper levels (2b), which fixes
// the grouping itself, further rounds change nothing (SAH 35.6 with {64}, 35.7 with {64, 32, 48}; nodes per ray
// equal) and cost ~5 ms of build per million triangles each: one round.
#ifndef PT_TREELET_ROUNDS
#define PT_TREELET_ROUNDS {64}
#endif

struct BuildIn {  // device copies of the flat scene arrays
  const pt_prim_ref* prim_ref;
  const pt_sphere* spheres;
  const pt_rect* rects;
  const pt_box* boxes;
  const pt_moving_sphere* moving_spheres;
  const pt_volume_sphere* volume_spheres;
  const pt_volume_box* volume_boxes;
  float cam_time0, cam_time1;  // shutter interval of the camera (bounds of moving spheres)
  const float* tri_v;
  const pt_xform* xforms;
  const pt_material* materials;
  int n;
};

struct Bvh2 {  // LBVH / refined binary BVH over Morton-sorted primitives
  Bvh2Node* nodes;    // n-1 inner nodes
  int* parent;        // parent of inner node (root: -1)
  int* leaf_parent;   // parent of sorted leaf j
  int* count;         // leaves under inner node
  int* flag;          // refit arrival counters
  const uint64_t* keys;     // sorted Morton codes
  const uint32_t* sorted;   // sorted position -> primitive id
  const float4* plo;        // per primitive id: padded AABB (xyz)
  const float4* phi;
  int n;
};

// ---- 1: primitive bounds ---------------------------------------------------------
// programs/hitables/moving_sphere.cu:44-46 `center(time)`, never-contracted arithmetic
PT_HD float3 moving_center(const float* c0, const float* c1, float t0, float t1, float time) {
  const float s = XDIV(XSUB(time, t0), XSUB(t1, t0));
  return f3(XADD(c0[0], XMUL(s, XSUB(c1[0], c0[0]))), XADD(c0[1], XMUL(s, XSUB(c1[1], c0[1]))),
            XADD(c0[2], XMUL(s, XSUB(c1[2], c0[2]))));
}

PT_HD void prim_bounds(const BuildIn& in, int id, float3& lo, float3& hi) {
  const pt_prim_ref p = in.prim_ref[id];
  if (p.type == PT_PRIM_TRIANGLE) {  // programs/hitables/triangle.cu:61-83
    const float* v = in.tri_v + 9 * (size_t)p.index;
    float3 a = f3(v[0], v[1], v[2]), b = f3(v[3], v[4], v[5]), c = f3(v[6], v[7], v[8]);
    lo = fmin3(fmin3(a, b), c);
    hi = fmax3(fmax3(a, b), c);
  } else if (p.type == PT_PRIM_SPHERE) {  // programs/hitables/sphere.cu:96-100
    const pt_sphere s = in.spheres[p.index];
    float3 c = f3(s.center[0], s.center[1], s.center[2]);
    lo = c - s.radius;
    hi = c + s.radius;
  } else if (p.type == PT_PRIM_RECT) {  // programs/hitables/aarect.cu:117-136
    const pt_rect q = in.rects[p.index];
    if (q.axis == PT_X_AXIS) {
      lo = f3(q.k, q.a0, q.b0);
      hi = f3(q.k, q.a1, q.b1);
    } else if (q.axis == PT_Y_AXIS) {
      lo = f3(q.a0, q.k, q.b0);
      hi = f3(q.a1, q.k, q.b1);
    } else {
      lo = f3(q.a0, q.b0, q.k);
      hi = f3(q.a1, q.b1, q.k);
    }
  } else if (p.type == PT_PRIM_MOVING_SPHERE) {  // programs/hitables/moving_sphere.cu:136-147
    // union of the sphere's boxes at its own time0/time1 (as the reference) and at the camera's
    // shutter times (a ray's time always lies in the camera interval; motion is linear)
    const pt_moving_sphere q = in.moving_spheres[p.index];
    const float ts[4] = {q.time0, q.time1, in.cam_time0, in.cam_time1};
    lo = f3(INFINITY);
    hi = f3(-INFINITY);
    for (int k = 0; k < 4; k++) {
      float3 c = moving_center(q.center0, q.center1, q.time0, q.time1, ts[k]);
      if (!(c.x == c.x && c.y == c.y && c.z == c.z)) continue;  // time0 == time1
      lo = fmin3(lo, c - q.radius);
      hi = fmax3(hi, c + q.radius);
    }
  } else if (p.type == PT_PRIM_VOLUME_SPHERE) {  // programs/hitables/volume_sphere.cu:111-115
    const pt_volume_sphere s = in.volume_spheres[p.index];
    float3 c = f3(s.center[0], s.center[1], s.center[2]);
    lo = c - s.radius;
    hi = c + s.radius;
  } else {  // programs/hitables/box.cu:84-88, volume_box.cu:86-90
    const float* q = p.type == PT_PRIM_VOLUME_BOX ? in.volume_boxes[p.index].bmin : in.boxes[p.index].bmin;
    lo = f3(q[0], q[1], q[2]);
    hi = f3(q[3], q[4], q[5]);
    float3 l2 = fmin3(lo, hi), h2 = fmax3(lo, hi);
    lo = l2;
    hi = h2;
  }
  if (p.xform >= 0) {  // Transform node: bound the 8 transformed corners
    const float* m = in.xforms[p.xform].o2w;
    float3 nlo = f3(INFINITY), nhi = f3(-INFINITY);
    for (int c = 0; c < 8; c++) {
      float3 q = xf_point(m, f3((c & 1) ? hi.x : lo.x, (c & 2) ? hi.y : lo.y, (c & 4) ? hi.z : lo.z));
      nlo = fmin3(nlo, q);
      nhi = fmax3(nhi, q);
    }
    lo = nlo;
    hi = nhi;
  }
}

// Padding.  The reference pads 1e-4 (rect/triangle) .. 1e-3 (box) on selected axes.  Here every
// axis gets a margin proportional to the primitive's own coordinates and to the scene scale S
// (largest |coordinate| in the scene), so the (FMA, quantised) box test can never cull a
// primitive the exact intersection test would accept: box-test rounding error grows with
// |coordinate| * 2^-23 and with the distance a ray travels (~S).
PT_HD void pad_bounds(float3& lo, float3& hi, float S) {
  float3 m = fmax3(f3(fabsf(lo.x), fabsf(lo.y), fabsf(lo.z)), f3(fabsf(hi.x), fabsf(hi.y), fabsf(hi.z)));
  float3 pad = m * 2e-5f + f3(2e-6f * S + 1e-7f);
  lo = lo - pad;
  hi = hi + pad;
}

// ---- 2: Morton code ---------------------------------------------------------------------
PT_HD uint64_t expand_bits21(uint64_t x) {
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
PT_HD uint64_t morton63(float3 c, float3 smin, float3 sinv) {
  float x = fminf(fmaxf((c.x - smin.x) * sinv.x * 2097152.0f, 0.0f), 2097151.0f);
  float y = fminf(fmaxf((c.y - smin.y) * sinv.y * 2097152.0f, 0.0f), 2097151.0f);
  float z = fminf(fmaxf((c.z - smin.z) * sinv.z * 2097152.0f, 0.0f), 2097151.0f);
  if (!(x == x)) x = 0.f;
  if (!(y == y)) y = 0.f;
  if (!(z == z)) z = 0.f;
  return (expand_bits21((uint64_t)x) << 2) | (expand_bits21((uint64_t)y) << 1) | expand_bits21((uint64_t)z);
}

PT_HD float box_half_area(float3 lo, float3 hi) {
  float3 e = hi - lo;
  return e.x * e.y + e.y * e.z + e.z * e.x;
}

// ---- 2b: top-down binned-SAH partition of the upper levels -----------------------------------------
// LBVH groups primitives by Morton cells, and the treelet refinement (5) can only re-arrange what those groups
// contain: on the 1 M-triangle mesh it stops at a SAH cost of 40.2 where a full top-down binned-SAH build reaches
// 35.3 (9 % fewer node visits per ray).  This pass decides the UPPER levels top-down instead: level by level every
// active node bins its primitives (16 bins per axis, counts and boxes by atomics), takes the cheapest split and
// hands each primitive one more PATH BIT (0 = left, 1 = right), until a node holds <= leaf_size primitives.  The
// path bits are then put IN FRONT of the Morton bits of the sort key, so the radix tree the Karras step builds over
// the sorted keys reproduces the top-down tree above the groups and is plain LBVH inside them - where the first
// treelet pass (<= 64 leaves) rebuilds everything anyway.  Nothing downstream changes.
// Measured in the host simulation (1 M triangles): SAH 40.2 -> 35.6 (a full top-down binned-SAH build: 35.3), BVH8
// nodes per camera / bounce ray 10.38 / 9.34 -> 9.67 / 8.66 (full build: 9.46 / 8.61).  On the B200: C4 +5.7 %,
// incoherent rays +6.1 %, C2 +3.5 %; the pass itself takes 2.3 ms per million triangles.
#ifndef PT_TD_BINS
#define PT_TD_BINS 16
#endif
#define PT_TD_MAXDEPTH 20
#ifndef PT_TD_LEAF
#define PT_TD_LEAF 64
#endif
#ifndef PT_TD_MAXNODES
#define PT_TD_MAXNODES 65536  // active nodes per level (leaf_size grows with the scene so that this holds)
#endif

struct TdBin {
  int count;
  float lo[3], hi[3];
};
struct TdNode {
  float cmin[3], cmax[3];  // centroid bounds of the node's primitives
  int axis, bin;           // chosen split: centroid bin <= bin goes left; axis < 0: the node is a finished group
  int child;               // index of the left child among the next level's nodes (right = child + 1)
  int pad;
};
PT_HD int td_leaf_size(int n) { return n / (PT_TD_MAXNODES / 2) > PT_TD_LEAF ? n / (PT_TD_MAXNODES / 2) : PT_TD_LEAF; }
PT_HD int td_bin_of(float c, float cmin, float cmax) {
  if (!(cmax > cmin)) return 0;
  const float scale = (float)PT_TD_BINS / (cmax - cmin);
  const int b = (int)((c - cmin) * scale);
  return b < 0 ? 0 : (b > PT_TD_BINS - 1 ? PT_TD_BINS - 1 : b);
}
// bins: [3][PT_TD_BINS] of one node, complete.  Chooses the split (or declares the node a finished group).
PT_HD void td_choose(const TdBin* bins, TdNode& nd, int leaf_size, int depth) {
  nd.axis = -1;
  nd.bin = 0;
  nd.child = -1;
  int total = 0;
  for (int b = 0; b < PT_TD_BINS; b++) total += bins[b].count;
  if (total <= leaf_size || depth >= PT_TD_MAXDEPTH) return;
  float best = INFINITY;
  for (int ax = 0; ax < 3; ax++) {
    if (!(nd.cmax[ax] > nd.cmin[ax])) continue;
    const TdBin* B = bins + ax * PT_TD_BINS;
    float rarea[PT_TD_BINS];
    int rcnt[PT_TD_BINS];
    float3 alo = f3(INFINITY), ahi = f3(-INFINITY);
    int c = 0;
    for (int b = PT_TD_BINS - 1; b > 0; b--) {
      if (B[b].count) {
        alo = fmin3(alo, f3(B[b].lo[0], B[b].lo[1], B[b].lo[2]));
        ahi = fmax3(ahi, f3(B[b].hi[0], B[b].hi[1], B[b].hi[2]));
      }
      c += B[b].count;
      rarea[b] = c ? box_half_area(alo, ahi) : 0.f;
      rcnt[b] = c;
    }
    alo = f3(INFINITY);
    ahi = f3(-INFINITY);
    c = 0;
    for (int b = 0; b < PT_TD_BINS - 1; b++) {
      if (B[b].count) {
        alo = fmin3(alo, f3(B[b].lo[0], B[b].lo[1], B[b].lo[2]));
        ahi = fmax3(ahi, f3(B[b].hi[0], B[b].hi[1], B[b].hi[2]));
      }
      c += B[b].count;
      if (!c || !rcnt[b + 1]) continue;
      const float cost = box_half_area(alo, ahi) * (float)c + rarea[b + 1] * (float)rcnt[b + 1];
      if (cost < best) {
        best = cost;
        nd.axis = ax;
        nd.bin = b;
      }
    }
  }
}
// sort key: `len` path bits in front of the Morton bits
PT_HD uint64_t td_compose_key(uint64_t morton, uint32_t code, int len) {
  return len ? (((uint64_t)code << (63 - len)) | (morton >> len)) : morton;
}

// ---- 3: Karras LBVH hierarchy --------------------------------------------------------
PT_HD int lbvh_delta(const uint64_t* keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + pt_clz32((uint32_t)i ^ (uint32_t)j);
  return pt_clz64(a ^ b);
}
// child refs: >= 0 inner node, < 0 sorted leaf ~ref
PT_HD void karras_node(const Bvh2& t, int i) {
  const int n = t.n;
  const uint64_t* k = t.keys;
  int d = (lbvh_delta(k, n, i, i + 1) - lbvh_delta(k, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = lbvh_delta(k, n, i, i - d);
  int lmax = 2;
  while (lbvh_delta(k, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int s = lmax / 2; s >= 1; s /= 2)
    if (lbvh_delta(k, n, i, i + (l + s) * d) > dmin) l += s;
  int j = i + l * d;
  int dnode = lbvh_delta(k, n, i, j);
  int s = 0;
  int tt = l;
  do {
    tt = (tt + 1) >> 1;
    if (lbvh_delta(k, n, i, i + (s + tt) * d) > dnode) s += tt;
  } while (tt > 1);
  int gamma = i + s * d + (d < 0 ? d : 0);
  int lo = i < j ? i : j, hi = i < j ? j : i;
  int left = (lo == gamma) ? ~gamma : gamma;
  int right = (hi == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
  t.nodes[i].left = left;
  t.nodes[i].right = right;
  if (left >= 0) t.parent[left] = i; else t.leaf_parent[~left] = i;
  if (right >= 0) t.parent[right] = i; else t.leaf_parent[~right] = i;
  if (i == 0) t.parent[0] = -1;
}

PT_HD void bvh2_child_box(const Bvh2& t, int ref, float3& lo, float3& hi, int& cnt) {
  if (ref >= 0) {
    const Bvh2Node& c = t.nodes[ref];
    lo = f3(c.lo[0], c.lo[1], c.lo[2]);
    hi = f3(c.hi[0], c.hi[1], c.hi[2]);
    cnt = t.count[ref];
  } else {
    uint32_t id = t.sorted[~ref];
    lo = xyz(t.plo[id]);
    hi = xyz(t.phi[id]);
    cnt = 1;
  }
}
PT_HD void bvh2_set_box(Bvh2Node& nd, float3 lo, float3 hi) {
  nd.lo[0] = lo.x; nd.lo[1] = lo.y; nd.lo[2] = lo.z;
  nd.hi[0] = hi.x; nd.hi[1] = hi.y; nd.hi[2] = hi.z;
}

// ---- 5: hierarchical binned-SAH treelet rebuild ------------------------------------------
// Pass k (k = 1,2,..) uses thresholds lo_k = PT_TREELET^(k-1), hi_k = PT_TREELET^k on the leaf
// count.  A node r is a pass-k root when count[r] <= hi_k < count[parent] (or r is the tree
// root).  One thread rebuilds the part of r's subtree ABOVE the pass-(k-1) roots: it expands r
// downwards (largest leaf count first) until every item has count <= lo_k or PT_TREELET items
// are collected, then rebuilds a binary tree over those items top-down with binned SAH (cost =
// half-area x leaf count), reusing exactly the inner-node slots it expanded.  Pass 1 therefore
// rebuilds all subtrees of <= PT_TREELET leaves down to the leaves (they are contiguous in Morton
// order); later passes fix the upper levels that plain LBVH leaves in Morton order.
PT_HD bool treelet_is_root(const Bvh2& t, int r, int lo, long long hi) {
  const int c = t.count[r];
  if (c < 3 || c <= lo || (long long)c > hi) return false;
  const int p = t.parent[r];
  return p < 0 || (long long)t.count[p] > hi;
}

PT_HD void treelet_sah(const Bvh2& t, int r, int lo_count) {
  int items[PT_TREELET];  // child refs: >= 0 inner node, < 0 sorted leaf ~ref
  int icnt[PT_TREELET];   // leaves below each item
  int slots[PT_TREELET];  // inner nodes expanded (owned by this treelet)
  int m = 0, ns = 0;
  items[0] = r;
  icnt[0] = t.count[r];
  m = 1;
  while (m < PT_TREELET) {  // expand the item with the most leaves that is still above the cut
    int bi = -1, bc = lo_count;
    for (int i = 0; i < m; i++)
      if (items[i] >= 0 && icnt[i] > bc) {
        bc = icnt[i];
        bi = i;
      }
    if (bi < 0) break;
    const int nd = items[bi];
    slots[ns++] = nd;
    const int l = t.nodes[nd].left, rr = t.nodes[nd].right;
    items[bi] = l;
    icnt[bi] = l >= 0 ? t.count[l] : 1;
    items[m] = rr;
    icnt[m] = rr >= 0 ? t.count[rr] : 1;
    m++;
  }
  if (ns == 0) return;
  // slots[0] == r: the treelet root keeps its index (its parent points at it)
  int wnode[PT_TREELET], wlo[PT_TREELET], whi[PT_TREELET];
  int wp = 1, next_slot = 1;
  wnode[0] = slots[0];
  wlo[0] = 0;
  whi[0] = m;
  while (wp) {
    --wp;
    const int node = wnode[wp], lo = wlo[wp], hi = whi[wp];
    float3 blo = f3(INFINITY), bhi = f3(-INFINITY), clo = blo, chi = bhi;
    int total = 0;
    for (int i = lo; i < hi; i++) {
      float3 a, b;
      int c;
      bvh2_child_box(t, items[i], a, b, c);
      blo = fmin3(blo, a);
      bhi = fmax3(bhi, b);
      float3 ce = (a + b) * 0.5f;
      clo = fmin3(clo, ce);
      chi = fmax3(chi, ce);
      total += icnt[i];
    }
    bvh2_set_box(t.nodes[node], blo, bhi);
    t.count[node] = total;
    int mid = -1;
    if (hi - lo > 2) {
      float best = INFINITY;
      int bestAxis = -1, bestBin = -1;
      for (int ax = 0; ax < 3; ax++) {
        float cmin = comp(clo, ax), cmax = comp(chi, ax);
        if (!(cmax > cmin)) continue;
        float scale = (float)PT_SAH_BINS / (cmax - cmin);
        float3 binlo[PT_SAH_BINS], binhi[PT_SAH_BINS];
        int bincnt[PT_SAH_BINS], binitems[PT_SAH_BINS];
        for (int bq = 0; bq < PT_SAH_BINS; bq++) {
          binlo[bq] = f3(INFINITY);
          binhi[bq] = f3(-INFINITY);
          bincnt[bq] = 0;
          binitems[bq] = 0;
        }
        for (int i = lo; i < hi; i++) {
          float3 a, b;
          int c;
          bvh2_child_box(t, items[i], a, b, c);
          float ce = (comp(a, ax) + comp(b, ax)) * 0.5f;
          int bi = (int)((ce - cmin) * scale);
          bi = bi < 0 ? 0 : (bi > PT_SAH_BINS - 1 ? PT_SAH_BINS - 1 : bi);
          binlo[bi] = fmin3(binlo[bi], a);
          binhi[bi] = fmax3(binhi[bi], b);
          bincnt[bi] += icnt[i];
          binitems[bi]++;
        }
        float rarea[PT_SAH_BINS];
        int rcnt[PT_SAH_BINS], ritems[PT_SAH_BINS];
        float3 alo = f3(INFINITY), ahi = f3(-INFINITY);
        int c = 0, ci = 0;
        for (int bq = PT_SAH_BINS - 1; bq > 0; bq--) {
          alo = fmin3(alo, binlo[bq]);
          ahi = fmax3(ahi, binhi[bq]);
          c += bincnt[bq];
          ci += binitems[bq];
          rarea[bq] = ci ? box_half_area(alo, ahi) : 0.f;
          rcnt[bq] = c;
          ritems[bq] = ci;
        }
        alo = f3(INFINITY);
        ahi = f3(-INFINITY);
        c = 0;
        ci = 0;
        for (int bq = 0; bq < PT_SAH_BINS - 1; bq++) {
          alo = fmin3(alo, binlo[bq]);
          ahi = fmax3(ahi, binhi[bq]);
          c += bincnt[bq];
          ci += binitems[bq];
          if (ci == 0 || ritems[bq + 1] == 0) continue;
          float cost = box_half_area(alo, ahi) * (float)c + rarea[bq + 1] * (float)rcnt[bq + 1];
          if (cost < best) {
            best = cost;
            bestAxis = ax;
            bestBin = bq;
          }
        }
      }
      if (bestAxis >= 0) {
        float cmin = comp(clo, bestAxis), cmax = comp(chi, bestAxis);
        float scale = (float)PT_SAH_BINS / (cmax - cmin);
        int i = lo, j = hi - 1;
        while (i <= j) {
          float3 a, b;
          int c;
          bvh2_child_box(t, items[i], a, b, c);
          float ce = (comp(a, bestAxis) + comp(b, bestAxis)) * 0.5f;
          int bi = (int)((ce - cmin) * scale);
          bi = bi < 0 ? 0 : (bi > PT_SAH_BINS - 1 ? PT_SAH_BINS - 1 : bi);
          if (bi <= bestBin) {
            i++;
          } else {
            int tmp = items[i]; items[i] = items[j]; items[j] = tmp;
            tmp = icnt[i]; icnt[i] = icnt[j]; icnt[j] = tmp;
            j--;
          }
        }
        mid = i;
      }
    }
    if (mid <= lo || mid >= hi) mid = (lo + hi) >> 1;  // degenerate: keep the current order, halve
    int refs[2];
    for (int side = 0; side < 2; side++) {
      int a = side ? mid : lo, b = side ? hi : mid;
      if (b - a == 1) {
        refs[side] = items[a];
        if (items[a] >= 0) t.parent[items[a]] = node;
        else t.leaf_parent[~items[a]] = node;
      } else {
        int child = slots[next_slot++];
        refs[side] = child;
        t.parent[child] = node;
        wnode[wp] = child;
        wlo[wp] = a;
        whi[wp] = b;
        wp++;
      }
    }
    t.nodes[node].left = refs[0];
    t.nodes[node].right = refs[1];
  }
}

// ---- 6: collapse to the compressed 8-wide BVH ----------------------------------------------
struct Collapse {
  Bvh2 t;
  BuildIn in;
  uint4* nodes8;
  LeafRec* leaves;
  int* prim_leaf;
  int* node_counter;   // next free BVH8 node
  int* leaf_counter;   // next free leaf record
  const unsigned char* dp;  // optional: per BVH2 inner node 9 decision bytes of the SAH-optimal collapse (collapse_dp)
};

// ---- 6a: SAH-optimal choice of the BVH8 children (dynamic programme of Ylitie, Karras, Laine 2017, section 3) ----
// C(n, i) = cheapest way to represent the subtree of BVH2 node n with at most i child slots of a wide node:
//   C(n, 1) = min( leaf: A(n) * P(n) * c_prim  if P(n) <= 3,   own wide node: D(n, 8) + A(n) * c_node )
//   C(n, i) = min( D(n, i), C(n, i - 1) ),   D(n, j) = min over 0 < k < j of C(left, k) + C(right, j - k)
// dec[1] = 1 when n becomes a wide node of its own (0: leaf child); dec[j], j = 2..8: the k of D(n, j), or 0 when
// fewer slots are as good (C(n, j) = C(n, j - 1)).  cost[0..7] = C(n, 1..7) + D(n, 8) in cost[7].
#ifndef PT_DP_CNODE
#define PT_DP_CNODE 1.0f
#endif
#ifndef PT_DP_CPRIM
#define PT_DP_CPRIM 0.6f
#endif
PT_HD void collapse_dp_node(float area, int count, const float* cl, const float* cr, float* cost, unsigned char* dec) {
  // cl / cr: C(child, 1..7) of the two children (a primitive: its area * c_prim for every i)
  float dist[9];
  unsigned char dk[9];
  for (int j = 2; j <= 8; j++) {
    float best = INFINITY;
    int bk = 1;
    for (int k = 1; k < j; k++) {
      if (k > 7 || j - k > 7) continue;
      const float c = cl[k - 1] + cr[j - k - 1];
      if (c < best) {
        best = c;
        bk = k;
      }
    }
    dist[j] = best;
    dk[j] = (unsigned char)bk;
  }
  const float c_leaf = count <= 3 ? area * (float)count * PT_DP_CPRIM : INFINITY;
  const float c_int = dist[8] + area * PT_DP_CNODE;
  cost[0] = c_leaf <= c_int ? c_leaf : c_int;
  dec[0] = 0;
  dec[1] = c_leaf <= c_int ? 0 : 1;
  for (int i = 2; i <= 7; i++) {
    if (dist[i] < cost[i - 2]) {
      cost[i - 1] = dist[i];
      dec[i] = dk[i];
    } else {
      cost[i - 1] = cost[i - 2];
      dec[i] = 0;
    }
  }
  dec[8] = dk[8];
}

PT_HD LeafRec make_leaf_rec(const BuildIn& in, int id) {
  const pt_prim_ref p = in.prim_ref[id];
  LeafRec L;
  uint32_t kindw = (uint32_t)p.type | ((uint32_t)(p.xform + 1) << 8);
  L.a = make_float4(0.f, 0.f, 0.f, pt_u2f((uint32_t)id));
  L.b = make_float4(0.f, 0.f, 0.f, pt_u2f(kindw));
  // material id in the low 24 bits, its PT_MAT_* type in the top 8 (lets the
  // trace kernel bin a path by material without touching the material table)
  L.c = make_float4(0.f, 0.f, 0.f,
                    pt_u2f(((uint32_t)p.material & 0xffffffu) | ((uint32_t)in.materials[p.material].type << 24)));
  if (p.type == PT_PRIM_TRIANGLE) {
    const float* v = in.tri_v + 9 * (size_t)p.index;
    L.a.x = v[0]; L.a.y = v[1]; L.a.z = v[2];
    L.b.x = v[3]; L.b.y = v[4]; L.b.z = v[5];
    L.c.x = v[6]; L.c.y = v[7]; L.c.z = v[8];
  } else if (p.type == PT_PRIM_SPHERE) {
    const pt_sphere s = in.spheres[p.index];
    L.a.x = s.center[0]; L.a.y = s.center[1]; L.a.z = s.center[2];
    L.b.x = s.radius;
  } else if (p.type == PT_PRIM_RECT) {
    const pt_rect q = in.rects[p.index];
    L.a.x = q.a0; L.a.y = q.a1; L.a.z = q.b0;
    L.b.x = q.b1; L.b.y = q.k;
    L.b.z = pt_u2f((uint32_t)(q.axis & 3) | (q.flip ? 4u : 0u));
  } else if (p.type == PT_PRIM_MOVING_SPHERE) {
    const pt_moving_sphere q = in.moving_spheres[p.index];
    L.a.x = q.center0[0]; L.a.y = q.center0[1]; L.a.z = q.center0[2];
    L.b.x = q.radius; L.b.y = q.time0; L.b.z = q.time1;
    L.c.x = q.center1[0]; L.c.y = q.center1[1]; L.c.z = q.center1[2];
  } else if (p.type == PT_PRIM_VOLUME_SPHERE) {
    const pt_volume_sphere q = in.volume_spheres[p.index];
    L.a.x = q.center[0]; L.a.y = q.center[1]; L.a.z = q.center[2];
    L.b.x = q.radius; L.b.y = q.density;
  } else if (p.type == PT_PRIM_VOLUME_BOX) {
    const pt_volume_box q = in.volume_boxes[p.index];
    L.a.x = q.bmin[0]; L.a.y = q.bmin[1]; L.a.z = q.bmin[2];
    L.b.x = q.bmax[0]; L.b.y = q.bmax[1]; L.b.z = q.bmax[2];
    L.c.x = q.density;
  } else {
    const pt_box q = in.boxes[p.index];
    L.a.x = q.bmin[0]; L.a.y = q.bmin[1]; L.a.z = q.bmin[2];
    L.b.x = q.bmax[0]; L.b.y = q.bmax[1]; L.b.z = q.bmax[2];
  }
  return L;
}

PT_HD uint32_t quant_exp(float ext) {  // biased exponent byte: 2^(eb-127)*255 >= ext
  float s = ext / 255.0f;
  uint32_t bits = pt_f2u(s);
  uint32_t eb = (bits >> 23) & 0xffu;
  if (bits & 0x7fffffu) eb += 1;
  if (eb < 1) eb = 1;
  if (eb > 254) eb = 254;
  while (eb < 254 && ceilf(ext / pt_u2f(eb << 23)) > 255.0f) eb++;
  return eb;
}

#ifdef __CUDA_ARCH__
#define PT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
PT_HD int pt_host_atomic_add(int* p, int v) { int o = *p; *p += v; return o; }
#define PT_ATOMIC_ADD(p, v) pt_host_atomic_add((p), (v))
#endif

// Builds BVH8 node `dst` from BVH2 subtree `ref`; appends inner children to the
// next-level work list (pairs ref, dst) via out_count.
PT_HD void collapse_node(const Collapse& c, int ref, int dst, uint2* out_work, int* out_count) {
  const Bvh2& t = c.t;
  int cref[8], ccnt[8];
  float3 clo[8], chi[8];
  int n = 0;
  if (ref >= 0) {
    cref[0] = t.nodes[ref].left;
    cref[1] = t.nodes[ref].right;
    n = 2;
  } else {
    cref[0] = ref;
    n = 1;
  }
  for (int i = 0; i < n; i++) bvh2_child_box(t, cref[i], clo[i], chi[i], ccnt[i]);
  // Quantisation guard: child boxes are stored on this node's 8-bit grid, whose step is the node's
  // extent / 255.  A subtree that spans only a few steps (small objects next to a huge one: the 1000-
  // unit ground sphere of the RTOW scenes, a fog ball around a whole scene, room walls) would have
  // all of its children rounded out to about the same coarse box, and every ray entering it would
  // visit all of them.  Such a subtree is not opened here: it stays ONE inner child and gets a node
  // (and a grid) of its own.  PT_QUANT_MIN_STEPS = 0 disables the rule.
#ifndef PT_QUANT_MIN_STEPS
#define PT_QUANT_MIN_STEPS 16
#endif
  float node_ext = 0.f;
  {
    float3 ulo = clo[0], uhi = chi[0];
    for (int i = 1; i < n; i++) {
      ulo = fmin3(ulo, clo[i]);
      uhi = fmax3(uhi, chi[i]);
    }
    node_ext = fmaxf(fmaxf(uhi.x - ulo.x, uhi.y - ulo.y), uhi.z - ulo.z);
  }
  const float min_child_ext = node_ext * ((float)PT_QUANT_MIN_STEPS / 255.0f);
  if (c.dp && ref >= 0) {
    // children from the dynamic programme: distribute 8 slots over the subtree as recorded
    int sref[16], sslots[16], sp2 = 0;
    n = 0;
    sref[0] = ref;
    sslots[0] = 8;
    sp2 = 1;
    bool top = true;
    while (sp2) {
      --sp2;
      const int rf = sref[sp2];
      int sl = sslots[sp2];
      bool open = false;
      int k = 0;
      if (rf >= 0) {
        const unsigned char* d = c.dp + (size_t)rf * 9;
        if (top) {
          open = true;
          k = d[8];
        } else {
          while (sl > 1 && d[sl] == 0) sl--;
          if (sl > 1) {
            open = true;
            k = d[sl];
            // quantisation guard (see above): a subtree spanning only a few steps of this node's grid stays one child
            const Bvh2Node& sn = t.nodes[rf];
            const float ext = fmaxf(fmaxf(sn.hi[0] - sn.lo[0], sn.hi[1] - sn.lo[1]), sn.hi[2] - sn.lo[2]);
            if (ext < min_child_ext && t.count[rf] > 3) open = false;
          }
        }
      }
      top = false;
      if (open) {
        sref[sp2] = t.nodes[rf].right;
        sslots[sp2] = sl - k;
        sp2++;
        sref[sp2] = t.nodes[rf].left;
        sslots[sp2] = k;
        sp2++;
      } else {
        cref[n] = rf;
        bvh2_child_box(t, rf, clo[n], chi[n], ccnt[n]);
        n++;
      }
    }
  } else
  // greedy expansion by surface area: first only subtrees too big for a leaf
  // (> 3 primitives), then any inner node while slots remain
  for (int pass = 0; pass < 2; pass++) {
    const int minc = pass == 0 ? 4 : 2;
    while (n < 8) {
      int bi = -1;
      float ba = -1.f;
      for (int i = 0; i < n; i++)
        if (cref[i] >= 0 && ccnt[i] >= minc) {
          const float3 e = chi[i] - clo[i];
          if (fmaxf(fmaxf(e.x, e.y), e.z) < min_child_ext && ccnt[i] > 3) continue;  // keep as one inner child
          float a = box_half_area(clo[i], chi[i]);
          // (tried: dissolving 4..6-primitive subtrees before flattening big ones - 15-25 % fewer BVH8 nodes but
          // the same node visits per ray within 1 %, the visits are at the upper levels; likewise treelets of
          // 32..256 leaves and 8 vs 16 SAH bins move nodes/ray by < 3 %)
          if (a > ba) {
            ba = a;
            bi = i;
          }
        }
      if (bi < 0) break;
      int l = t.nodes[cref[bi]].left, r = t.nodes[cref[bi]].right;
      cref[bi] = l;
      bvh2_child_box(t, l, clo[bi], chi[bi], ccnt[bi]);
      cref[n] = r;
      bvh2_child_box(t, r, clo[n], chi[n], ccnt[n]);
      n++;
    }
  }
  // node frame.  Child planes are quantised to 8 bits relative to (p, 2^e) per axis with a guard
  // band of G grid steps (covers the <= 2^-10-step rounding of the biased FMA in the box test,
  // traverse.h) and the origin is placed 1/64 step below the node minimum so the low side never
  // clamps: decoded boxes always contain the true boxes.
  float3 nlo = clo[0], nhi = chi[0];
  for (int i = 1; i < n; i++) {
    nlo = fmin3(nlo, clo[i]);
    nhi = fmax3(nhi, chi[i]);
  }
  const float G = 1.0f / 256.0f;
  uint32_t eb[3];
  float sc3[3], p3[3];
  for (int a = 0; a < 3; a++) {
    const float lo_a = comp(nlo, a), hi_a = comp(nhi, a);
    uint32_t e = quant_exp(hi_a - lo_a);
    for (;;) {
      const float sca = pt_u2f(e << 23);
      const float pa = lo_a - sca * (1.0f / 64.0f);
      if (e >= 254 || ceilf((hi_a - pa) / sca + G) <= 255.0f) {
        sc3[a] = sca;
        p3[a] = pa;
        break;
      }
      e++;
    }
    eb[a] = e;
  }
  const uint32_t ex = eb[0], ey = eb[1], ez = eb[2];
  const float sx = sc3[0], sy = sc3[1], sz = sc3[2];
  const float3 porg = f3(p3[0], p3[1], p3[2]);
  // octant slot assignment (greedy max of dot(child centre - node centre, slot dir))
  int slot_of[8], child_in[8];
  for (int i = 0; i < 8; i++) {
    slot_of[i] = -1;
    child_in[i] = -1;
  }
  float3 nc = (nlo + nhi) * 0.5f;
  for (int k = 0; k < n; k++) {
    float best = -INFINITY;
    int bc = -1, bs = -1;
    for (int i = 0; i < n; i++) {
      if (slot_of[i] >= 0) continue;
      float3 dc = (clo[i] + chi[i]) * 0.5f - nc;
      for (int s = 0; s < 8; s++) {
        if (child_in[s] >= 0) continue;
        float cost = ((s & 4) ? dc.x : -dc.x) + ((s & 2) ? dc.y : -dc.y) + ((s & 1) ? dc.z : -dc.z);
        if (cost > best) {
          best = cost;
          bc = i;
          bs = s;
        }
      }
    }
    slot_of[bc] = bs;
    child_in[bs] = bc;
  }
  int n_inner = 0, n_leafprims = 0;
  for (int i = 0; i < n; i++) {
    if (cref[i] >= 0 && ccnt[i] > 3) n_inner++;
    else n_leafprims += ccnt[i];
  }
  const int child_base = n_inner ? PT_ATOMIC_ADD(c.node_counter, n_inner) : 0;
  const int leaf_base = n_leafprims ? PT_ATOMIC_ADD(c.leaf_counter, n_leafprims) : 0;
  uint32_t imask = 0;
  uint32_t meta[8], qlo[3][8], qhi[3][8];
  int inner_seen = 0, leaf_off = 0;
  for (int s = 0; s < 8; s++) {
    meta[s] = 0;
    for (int a = 0; a < 3; a++) {
      qlo[a][s] = 255;
      qhi[a][s] = 0;
    }
    const int i = child_in[s];
    if (i < 0) continue;
    float q;
    q = floorf((clo[i].x - porg.x) / sx - G); qlo[0][s] = (uint32_t)fminf(fmaxf(q, 0.f), 255.f);
    q = floorf((clo[i].y - porg.y) / sy - G); qlo[1][s] = (uint32_t)fminf(fmaxf(q, 0.f), 255.f);
    q = floorf((clo[i].z - porg.z) / sz - G); qlo[2][s] = (uint32_t)fminf(fmaxf(q, 0.f), 255.f);
    q = ceilf((chi[i].x - porg.x) / sx + G); qhi[0][s] = (uint32_t)fminf(fmaxf(q, 0.f), 255.f);
    q = ceilf((chi[i].y - porg.y) / sy + G); qhi[1][s] = (uint32_t)fminf(fmaxf(q, 0.f), 255.f);
    q = ceilf((chi[i].z - porg.z) / sz + G); qhi[2][s] = (uint32_t)fminf(fmaxf(q, 0.f), 255.f);
    if (cref[i] >= 0 && ccnt[i] > 3) {
      meta[s] = (1u << 5) | (24u + (uint32_t)s);
      imask |= 1u << s;
      int w = PT_ATOMIC_ADD(out_count, 1);
      out_work[w] = make_uint2((uint32_t)cref[i], (uint32_t)(child_base + inner_seen));
      inner_seen++;
    } else {
      // leaf child: 1..3 primitives, enumerate the tiny subtree
      int prims[3], np = 0;
      int st[4], sp = 0;
      st[sp++] = cref[i];
      while (sp) {
        int rf = st[--sp];
        if (rf < 0) prims[np++] = (int)t.sorted[~rf];
        else {
          st[sp++] = t.nodes[rf].right;
          st[sp++] = t.nodes[rf].left;
        }
      }
      const uint32_t unary = np == 1 ? 1u : (np == 2 ? 3u : 7u);
      meta[s] = (unary << 5) | (uint32_t)leaf_off;
      for (int k = 0; k < np; k++) {
        c.leaves[leaf_base + leaf_off + k] = make_leaf_rec(c.in, prims[k]);
        c.prim_leaf[prims[k]] = leaf_base + leaf_off + k;
      }
      leaf_off += np;
    }
  }
  auto pack4 = [](const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); };
  uint4* out = c.nodes8 + (size_t)dst * 5;
  out[0] = make_uint4(pt_f2u(porg.x), pt_f2u(porg.y), pt_f2u(porg.z), ex | (ey << 8) | (ez << 16) | (imask << 24));
  out[1] = make_uint4((uint32_t)child_base, (uint32_t)leaf_base, pack4(meta), pack4(meta + 4));
  out[2] = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
  out[3] = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
  out[4] = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
}
