// tests/hostsim/sim.cpp — TEST-ONLY host simulation of the kernel cores.
//
// Compiles optix-path-tracer_b200/csrc/{bvh_core,traverse,shading}.h with g++
// (the PT_HD bodies are written to be host-compilable) and drives them with
// serial loops that mirror the CUDA drivers (bvh_build.cu, integrator.cu).
// Purpose: check the builder / traversal / shading LOGIC against the oracle on
// a machine without a GPU.  This is not a product path: nothing in the package
// loads it, and the shipped library has no CPU execution mode.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../optix-path-tracer_b200/csrc/bvh_core.h"
#include "../../optix-path-tracer_b200/csrc/sample_order.h"
#include "../../optix-path-tracer_b200/csrc/shading.h"
#include "../../optix-path-tracer_b200/csrc/traverse.h"

struct Sim {
  pt_scene_desc d;
  std::vector<uint4> nodes8;
  std::vector<LeafRec> leaves;
  std::vector<int> prim_leaf;
  std::vector<ImageDev> images;
  SceneDev sc;
  int n_nodes8 = 0;
  double sah_before = 0, sah_after = 0;
  int n_treelets = 0;
};

static double sah_sum(const Bvh2& t) {
  double a = 0;
  for (int i = 0; i < t.n - 1; i++) {
    const Bvh2Node& nd = t.nodes[i];
    a += box_half_area(f3(nd.lo[0], nd.lo[1], nd.lo[2]), f3(nd.hi[0], nd.hi[1], nd.hi[2]));
  }
  const Bvh2Node& r = t.nodes[0];
  double ra = box_half_area(f3(r.lo[0], r.lo[1], r.lo[2]), f3(r.hi[0], r.hi[1], r.hi[2]));
  return ra > 0 ? a / ra : 0;
}


// ---- EXPERIMENT (dev): full top-down binned-SAH build into the same Bvh2 arrays, as the quality yardstick for the
// GPU builder's LBVH + treelet refinement (scripts/bvh_quality_probe.py, PTSIM_BUILDER=sah) -----------------
static void full_sah_build(Bvh2& t, int nbins) {
  const int n = t.n;
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);  // sorted positions
  struct Job { int node, lo, hi; };
  std::vector<Job> stack;
  int next_node = 1;
  stack.push_back({0, 0, n});
  t.parent[0] = -1;
  while (!stack.empty()) {
    Job j = stack.back();
    stack.pop_back();
    float3 blo = f3(INFINITY), bhi = f3(-INFINITY), clo = blo, chi = bhi;
    for (int i = j.lo; i < j.hi; i++) {
      uint32_t id = t.sorted[idx[i]];
      float3 a = xyz(t.plo[id]), b = xyz(t.phi[id]);
      blo = fmin3(blo, a); bhi = fmax3(bhi, b);
      float3 c = (a + b) * 0.5f;
      clo = fmin3(clo, c); chi = fmax3(chi, c);
    }
    bvh2_set_box(t.nodes[j.node], blo, bhi);
    t.count[j.node] = j.hi - j.lo;
    int mid = -1;
    if (j.hi - j.lo > 2) {
      float best = INFINITY; int bax = -1, bbin = -1;
      for (int ax = 0; ax < 3; ax++) {
        float cmin = comp(clo, ax), cmax = comp(chi, ax);
        if (!(cmax > cmin)) continue;
        float scale = nbins / (cmax - cmin);
        std::vector<float3> binlo(nbins, f3(INFINITY)), binhi(nbins, f3(-INFINITY));
        std::vector<int> cnt(nbins, 0);
        for (int i = j.lo; i < j.hi; i++) {
          uint32_t id = t.sorted[idx[i]];
          float3 a = xyz(t.plo[id]), b = xyz(t.phi[id]);
          int bi = std::min(nbins - 1, std::max(0, (int)(((comp(a, ax) + comp(b, ax)) * 0.5f - cmin) * scale)));
          binlo[bi] = fmin3(binlo[bi], a); binhi[bi] = fmax3(binhi[bi], b); cnt[bi]++;
        }
        std::vector<float> rarea(nbins); std::vector<int> rcnt(nbins);
        float3 alo = f3(INFINITY), ahi = f3(-INFINITY); int c = 0;
        for (int b = nbins - 1; b > 0; b--) { alo = fmin3(alo, binlo[b]); ahi = fmax3(ahi, binhi[b]); c += cnt[b]; rarea[b] = c ? box_half_area(alo, ahi) : 0.f; rcnt[b] = c; }
        alo = f3(INFINITY); ahi = f3(-INFINITY); c = 0;
        for (int b = 0; b < nbins - 1; b++) {
          alo = fmin3(alo, binlo[b]); ahi = fmax3(ahi, binhi[b]); c += cnt[b];
          if (!c || !rcnt[b + 1]) continue;
          float cost = box_half_area(alo, ahi) * c + rarea[b + 1] * rcnt[b + 1];
          if (cost < best) { best = cost; bax = ax; bbin = b; }
        }
      }
      if (bax >= 0) {
        float cmin = comp(clo, bax), cmax = comp(chi, bax);
        float scale = nbins / (cmax - cmin);
        auto it = std::partition(idx.begin() + j.lo, idx.begin() + j.hi, [&](int k) {
          uint32_t id = t.sorted[k];
          float3 a = xyz(t.plo[id]), b = xyz(t.phi[id]);
          int bi = std::min(nbins - 1, std::max(0, (int)(((comp(a, bax) + comp(b, bax)) * 0.5f - cmin) * scale)));
          return bi <= bbin;
        });
        mid = (int)(it - idx.begin());
      }
    }
    if (mid <= j.lo || mid >= j.hi) mid = (j.lo + j.hi) >> 1;
    int refs[2];
    for (int side = 0; side < 2; side++) {
      int a = side ? mid : j.lo, b = side ? j.hi : mid;
      if (b - a == 1) { refs[side] = ~idx[a]; t.leaf_parent[idx[a]] = j.node; }
      else { int ch = next_node++; refs[side] = ch; t.parent[ch] = j.node; stack.push_back({ch, a, b}); }
    }
    t.nodes[j.node].left = refs[0];
    t.nodes[j.node].right = refs[1];
  }
}

extern "C" {

Sim* sim_create(const pt_scene_desc* desc, int refine) {
  Sim* s = new Sim();
  s->d = *desc;
  const int n = desc->n_prims;
  BuildIn in;
  in.prim_ref = desc->prims;
  in.spheres = desc->spheres;
  in.rects = desc->rects;
  in.boxes = desc->boxes;
  in.moving_spheres = desc->moving_spheres;
  in.volume_spheres = desc->volume_spheres;
  in.volume_boxes = desc->volume_boxes;
  in.cam_time0 = desc->camera.time0;
  in.cam_time1 = desc->camera.time1;
  in.tri_v = desc->tri_v;
  in.xforms = desc->xforms;
  in.materials = desc->materials;
  in.n = n;
  std::vector<float4> plo(n), phi(n);
  float3 cmin = f3(INFINITY), cmax = f3(-INFINITY);
  float S = 0.f;
  for (int i = 0; i < n; i++) {
    float3 lo, hi;
    prim_bounds(in, i, lo, hi);
    plo[i] = make_float4(lo.x, lo.y, lo.z, 0.f);
    phi[i] = make_float4(hi.x, hi.y, hi.z, 0.f);
    float3 c = (lo + hi) * 0.5f;
    cmin = fmin3(cmin, c);
    cmax = fmax3(cmax, c);
    S = fmaxf(S, fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(lo.y)), fabsf(lo.z)), fmaxf(fmaxf(fabsf(hi.x), fabsf(hi.y)), fabsf(hi.z))));
  }
  for (int i = 0; i < n; i++) {
    float3 lo = xyz(plo[i]), hi = xyz(phi[i]);
    pad_bounds(lo, hi, S);
    plo[i] = make_float4(lo.x, lo.y, lo.z, 0.f);
    phi[i] = make_float4(hi.x, hi.y, hi.z, 0.f);
  }
  float3 ext = cmax - cmin;
  float3 sinv = f3(ext.x > 0 ? 1.f / ext.x : 0.f, ext.y > 0 ? 1.f / ext.y : 0.f, ext.z > 0 ? 1.f / ext.z : 0.f);
  std::vector<uint64_t> keys(n);
  std::vector<uint32_t> order(n);
  for (int i = 0; i < n; i++) keys[i] = morton63((xyz(plo[i]) + xyz(phi[i])) * 0.5f, cmin, sinv);
  // top-down binned-SAH partition of the upper levels (mirrors bvh_build.cu: same bodies, serial loops)
  {
    const int leaf_size = td_leaf_size(n);
    std::vector<uint32_t> code(n, 0u);
    std::vector<int> clen(n, 0), node_id(n, 0);
    if (n > leaf_size && refine && !getenv("PTSIM_NO_TOPDOWN")) {
      std::vector<TdNode> cur(1), nxt;
      for (int a = 0; a < 3; a++) {
        cur[0].cmin[a] = comp(cmin, a);
        cur[0].cmax[a] = comp(cmax, a);
      }
      cur[0].axis = -1;
      cur[0].bin = 0;
      cur[0].child = -1;
      for (int level = 0; level < PT_TD_MAXDEPTH && !cur.empty(); level++) {
        std::vector<TdBin> bins(cur.size() * 3 * PT_TD_BINS);
        for (auto& bq : bins) {
          bq.count = 0;
          for (int a = 0; a < 3; a++) {
            bq.lo[a] = FLT_MAX;
            bq.hi[a] = -FLT_MAX;
          }
        }
        for (int i = 0; i < n; i++) {
          const int id = node_id[i];
          if (id < 0) continue;
          const float3 lo = xyz(plo[i]), hi = xyz(phi[i]);
          const float3 c = (lo + hi) * 0.5f;
          for (int a = 0; a < 3; a++) {
            TdBin& bq = bins[(size_t)(id * 3 + a) * PT_TD_BINS + td_bin_of(comp(c, a), cur[id].cmin[a], cur[id].cmax[a])];
            bq.count++;
            for (int k = 0; k < 3; k++) {
              bq.lo[k] = fminf(bq.lo[k], comp(lo, k));
              bq.hi[k] = fmaxf(bq.hi[k], comp(hi, k));
            }
          }
        }
        nxt.clear();
        for (size_t id = 0; id < cur.size(); id++) {
          td_choose(bins.data() + id * 3 * PT_TD_BINS, cur[id], leaf_size, level);
          if (cur[id].axis >= 0) {
            cur[id].child = (int)nxt.size();
            TdNode ch;
            for (int a = 0; a < 3; a++) {
              ch.cmin[a] = FLT_MAX;
              ch.cmax[a] = -FLT_MAX;
            }
            ch.axis = -1;
            ch.bin = 0;
            ch.child = -1;
            nxt.push_back(ch);
            nxt.push_back(ch);
          }
        }
        for (int i = 0; i < n; i++) {
          const int id = node_id[i];
          if (id < 0) continue;
          const TdNode& nd = cur[id];
          if (nd.axis < 0) {
            node_id[i] = -1;
            continue;
          }
          const float3 c = (xyz(plo[i]) + xyz(phi[i])) * 0.5f;
          const int side = td_bin_of(comp(c, nd.axis), nd.cmin[nd.axis], nd.cmax[nd.axis]) <= nd.bin ? 0 : 1;
          const int ch = nd.child + side;
          node_id[i] = ch;
          code[i] = (code[i] << 1) | (uint32_t)side;
          clen[i]++;
          for (int a = 0; a < 3; a++) {
            nxt[ch].cmin[a] = fminf(nxt[ch].cmin[a], comp(c, a));
            nxt[ch].cmax[a] = fmaxf(nxt[ch].cmax[a], comp(c, a));
          }
        }
        cur.swap(nxt);
      }
    }
    for (int i = 0; i < n; i++) keys[i] = td_compose_key(keys[i], code[i], clen[i]);
  }
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  std::vector<uint64_t> skeys(n);
  for (int i = 0; i < n; i++) skeys[i] = keys[order[i]];

  const int ni = n > 1 ? n - 1 : 1;
  std::vector<Bvh2Node> nodes2(ni);
  std::vector<int> parent(ni, -1), leaf_parent(n, -1), count(ni, 0), flag(ni, 0);
  Bvh2 t;
  t.nodes = nodes2.data();
  t.parent = parent.data();
  t.leaf_parent = leaf_parent.data();
  t.count = count.data();
  t.flag = flag.data();
  t.keys = skeys.data();
  t.sorted = order.data();
  t.plo = plo.data();
  t.phi = phi.data();
  t.n = n;
  if (n > 1) {
    for (int i = 0; i < n - 1; i++) karras_node(t, i);
    // refit (mirrors k_refit: second arrival computes the node)
    for (int j = 0; j < n; j++) {
      int cur = t.leaf_parent[j];
      while (cur >= 0) {
        if (t.flag[cur]++ == 0) break;
        Bvh2Node& nd = t.nodes[cur];
        float3 llo, lhi, rlo, rhi;
        int lc, rc;
        bvh2_child_box(t, nd.left, llo, lhi, lc);
        bvh2_child_box(t, nd.right, rlo, rhi, rc);
        bvh2_set_box(nd, fmin3(llo, rlo), fmax3(lhi, rhi));
        t.count[cur] = lc + rc;
        cur = t.parent[cur];
      }
    }
    s->sah_before = sah_sum(t);
    if (refine) {
      static const int rounds[] = PT_TREELET_ROUNDS;  // mirrors bvh_build.cu
      for (int T : rounds) {
        long long lo = 1, hi = T;
        for (int pass = 0; pass < 12 && lo < (long long)n; pass++) {
          std::vector<int> roots;
          for (int i = 0; i < n - 1; i++)
            if (treelet_is_root(t, i, (int)lo, hi)) roots.push_back(i);
          for (int r : roots) treelet_sah(t, r, (int)lo);
          s->n_treelets += (int)roots.size();
          lo = hi;
          hi *= T;
        }
      }
    }
    if (const char* e = getenv("PTSIM_BUILDER")) {
      if (!strcmp(e, "sah")) full_sah_build(t, getenv("PTSIM_BINS") ? atoi(getenv("PTSIM_BINS")) : 32);
    }
    s->sah_after = sah_sum(t);
  }
  s->nodes8.resize((size_t)(n > 1 ? n : 2) * 5);
  s->leaves.resize(n);
  s->prim_leaf.assign(n, -1);
  int counters[2] = {1, 0};
  // SAH-optimal collapse: decisions per BVH2 inner node, bottom-up
  std::vector<unsigned char> dp;
  if (n > 1 && refine && !getenv("PTSIM_NO_DP")) {  // mirrors k_collapse_dp
    dp.assign((size_t)(n - 1) * 9, 0);
    std::vector<float> cost((size_t)(n - 1) * 8, 0.f);
    // post-order over inner nodes
    std::vector<int> order2, stk = {0};
    while (!stk.empty()) {
      int v = stk.back();
      stk.pop_back();
      order2.push_back(v);
      if (t.nodes[v].left >= 0) stk.push_back(t.nodes[v].left);
      if (t.nodes[v].right >= 0) stk.push_back(t.nodes[v].right);
    }
    for (size_t q = order2.size(); q-- > 0;) {
      const int v = order2[q];
      float cl[7], cr[7];
      for (int side = 0; side < 2; side++) {
        const int rf = side ? t.nodes[v].right : t.nodes[v].left;
        float* dst = side ? cr : cl;
        if (rf >= 0) {
          for (int i = 0; i < 7; i++) dst[i] = cost[(size_t)rf * 8 + i];
        } else {
          float3 a, b;
          int cn;
          bvh2_child_box(t, rf, a, b, cn);
          for (int i = 0; i < 7; i++) dst[i] = box_half_area(a, b) * PT_DP_CPRIM;
        }
      }
      const Bvh2Node& nd = t.nodes[v];
      collapse_dp_node(box_half_area(f3(nd.lo[0], nd.lo[1], nd.lo[2]), f3(nd.hi[0], nd.hi[1], nd.hi[2])), t.count[v], cl, cr,
                       &cost[(size_t)v * 8], &dp[(size_t)v * 9]);
    }
  }
  Collapse c;
  c.dp = dp.empty() ? nullptr : dp.data();
  c.t = t;
  c.in = in;
  c.nodes8 = s->nodes8.data();
  c.leaves = s->leaves.data();
  c.prim_leaf = s->prim_leaf.data();
  c.node_counter = &counters[0];
  c.leaf_counter = &counters[1];
  std::vector<uint2> wa(n + 1), wb(n + 1);
  wa[0] = make_uint2(n > 1 ? 0u : (uint32_t)~0, 0u);
  int nw = 1;
  while (nw > 0) {
    int out = 0;
    for (int i = 0; i < nw; i++) collapse_node(c, (int)wa[i].x, (int)wa[i].y, wb.data(), &out);
    std::swap(wa, wb);
    nw = out;
  }
  s->n_nodes8 = counters[0];
  if (counters[1] != n) fprintf(stderr, "hostsim: leaf count %d != %d\n", counters[1], n);

  for (int i = 0; i < desc->n_images; i++) {
    ImageDev im;
    im.texels = (const float4*)desc->images[i].rgba;
    im.width = desc->images[i].width;
    im.height = desc->images[i].height;
    s->images.push_back(im);
  }
  SceneDev& d = s->sc;
  d.nodes8 = s->nodes8.data();
  d.leaves = s->leaves.data();
  d.prim_ref = desc->prims;
  d.prim_leaf = s->prim_leaf.data();
  d.xforms = desc->xforms;
  d.tri_n = desc->tri_n;
  d.tri_uv = desc->tri_uv;
  d.tri_tex = desc->tri_tex;
  d.materials = desc->materials;
  d.textures = desc->textures;
  d.texvec = desc->texvec;
  d.texdata = desc->texdata;
  d.n_texdata = desc->n_texdata;
  d.images = s->images.data();
  d.lights = desc->lights;
  d.env_alias_prob = desc->env_alias_prob;
  d.env_alias_idx = desc->env_alias_idx;
  d.env_texel_pdf = desc->env_texel_pdf;
  d.env_n = desc->env_n;
  d.env_w = desc->env_n ? desc->images[desc->miss.image].width : 0;
  d.env_h = desc->env_n ? desc->images[desc->miss.image].height : 0;
  d.tex_need_uv = 0;
  for (int i = 0; i < desc->n_textures; i++) d.tex_need_uv |= desc->textures[i].type == PT_TEX_IMAGE;
  d.qbias = 0x47000000u;
  d.special_mask = (desc->n_moving_spheres ? 1 : 0) | (desc->n_volume_spheres || desc->n_volume_boxes ? 2 : 0);
  d.n_prims = n;
  d.n_materials = desc->n_materials;
  d.n_textures = desc->n_textures;
  d.n_texvec = desc->n_texvec;
  d.n_images = desc->n_images;
  d.n_lights = desc->n_lights;
  d.camera = desc->camera;
  d.miss = desc->miss;
  return s;
}

void sim_destroy(Sim* s) { delete s; }

void sim_info(const Sim* s, int* n_nodes8, double* sah_before, double* sah_after, int* n_treelets) {
  *n_nodes8 = s->n_nodes8;
  *sah_before = s->sah_before;
  *sah_after = s->sah_after;
  *n_treelets = s->n_treelets;
}

// every primitive must be referenced by exactly one leaf slot reachable from the root,
// and every child box must contain its subtree's primitive boxes (checked via leaves)
int sim_validate(const Sim* s) {
  const int n = s->d.n_prims;
  std::vector<int> seen(n, 0);
  std::vector<int> stack = {0};
  int visited = 0;
  while (!stack.empty()) {
    int ni = stack.back();
    stack.pop_back();
    visited++;
    const uint4* np = s->nodes8.data() + (size_t)ni * 5;
    uint32_t imask = np[0].w >> 24;
    uint32_t meta[8];
    for (int k = 0; k < 4; k++) {
      meta[k] = (np[1].z >> (8 * k)) & 0xff;
      meta[4 + k] = (np[1].w >> (8 * k)) & 0xff;
    }
    int rel = 0;
    for (int slot = 0; slot < 8; slot++) {
      uint32_t m = meta[slot];
      if (!m) continue;
      if (imask & (1u << slot)) {
        if ((m >> 5) != 1 || (m & 31) != 24u + slot) return -2;
        stack.push_back((int)np[1].x + rel);
        rel++;
      } else {
        int cnt = pt_popc(m >> 5);
        int off = m & 31;
        for (int k = 0; k < cnt; k++) {
          int li = (int)np[1].y + off + k;
          if (li < 0 || li >= n) return -3;
          int prim = (int)pt_f2u(s->leaves[li].a.w);
          if (prim < 0 || prim >= n) return -4;
          seen[prim]++;
          if (s->prim_leaf[prim] != li) return -5;
        }
      }
    }
  }
  for (int i = 0; i < n; i++)
    if (seen[i] != 1) return -1;
  if (visited != s->n_nodes8) return -6;
  return 0;
}

// children per BVH8 node, primitives per leaf child, and how many nodes have an empty upper half (slots 4..7)
void sim_child_hist(const Sim* s, int* children /*9*/, int* leaf_prims /*4*/, int* upper_empty) {
  for (int i = 0; i < 9; i++) children[i] = 0;
  for (int i = 0; i < 4; i++) leaf_prims[i] = 0;
  *upper_empty = 0;
  for (int ni = 0; ni < s->n_nodes8; ni++) {
    const uint4* np = s->nodes8.data() + (size_t)ni * 5;
    int c = 0;
    for (int k = 0; k < 8; k++) {
      uint32_t m = ((k < 4 ? np[1].z : np[1].w) >> (8 * (k & 3))) & 0xff;
      if (!m) continue;
      c++;
      if (!((np[0].w >> 24) & (1u << k))) leaf_prims[pt_popc(m >> 5)]++;
    }
    children[c]++;
    if (np[1].w == 0) ++*upper_empty;
  }
}

// the node step's octant table (traverse.h pt_perm8): bit s -> bit s ^ octinv
unsigned sim_perm8(unsigned m, unsigned octinv) { return pt_perm8(m & 0xffu, octinv & 7u); }

// Invariants the node step's hit-mask assembly relies on, counted over all BVH8 nodes: an empty child slot has
// meta 0, no imask bit and the inverted box (qlo 255, qhi 0 on every axis) so that it can never pass the slab test;
// an inner slot s has meta (1 << 5) | (24 + s); a leaf slot's records stay inside the 24-bit leaf field.
int sim_node_invariant_violations(const Sim* s) {
  int bad = 0;
  for (int ni = 0; ni < s->n_nodes8; ni++) {
    const uint4* np = s->nodes8.data() + (size_t)ni * 5;
    const uint32_t imask = np[0].w >> 24;
    const uint32_t q[12] = {np[2].x, np[2].y, np[2].z, np[2].w, np[3].x, np[3].y, np[3].z, np[3].w, np[4].x, np[4].y, np[4].z, np[4].w};
    for (int k = 0; k < 8; k++) {
      const uint32_t m = ((k < 4 ? np[1].z : np[1].w) >> (8 * (k & 3))) & 0xff;
      const bool inner = (imask >> k) & 1u;
      uint32_t lo[3], hi[3];
      for (int a = 0; a < 3; a++) {
        lo[a] = (q[2 * a + (k >> 2)] >> (8 * (k & 3))) & 0xff;
        hi[a] = (q[6 + 2 * a + (k >> 2)] >> (8 * (k & 3))) & 0xff;
      }
      if (m == 0) {
        if (inner) bad++;
        for (int a = 0; a < 3; a++)
          if (lo[a] != 255 || hi[a] != 0) bad++;
      } else if (inner) {
        if (m != ((1u << 5) | (24u + (uint32_t)k))) bad++;
      } else {
        const uint32_t unary = m >> 5, off = m & 31u;
        if (!(unary == 1 || unary == 3 || unary == 7) || off + (uint32_t)pt_popc(unary) > 24u) bad++;
      }
    }
  }
  return bad;
}

int sim_trace(const Sim* s, const pt_ray* rays, long long n, pt_hit* hits, unsigned long long* counters) {
  unsigned long long cn = 0, cl = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : cn, cl)
  for (long long i = 0; i < n; i++) {
    const pt_ray& r = rays[i];
    TraceHit h;
    uint32_t a = 0, b = 0;
    if (r.tmax > r.tmin) {
      bvh8_closest<true>(s->sc, f3(r.origin[0], r.origin[1], r.origin[2]), f3(r.dir[0], r.dir[1], r.dir[2]), r.tmin,
                         r.tmax, h, a, b, RayAux{s->sc.camera.time0, (uint32_t)i});
    } else {
      h.t = r.tmax;
      h.prim = -1;
      h.u = h.v = 0.f;
    }
    hits[i].t = h.t;
    hits[i].prim = h.prim;
    hits[i].u = h.u;
    hits[i].v = h.v;
    cn += a;
    cl += b;
  }
  if (counters) {
    counters[0] = cn;
    counters[1] = cl;
  }
  return 0;
}

// serial mirror of the wavefront loop (integrator.cu) for one path at a time
int sim_render(const Sim* s, int W, int H, int first_frame, int nframes, int max_depth, int russian, float* accum) {
  const SceneDev& sc = s->sc;
#pragma omp parallel for schedule(dynamic, 1)
  for (int y = 0; y < H; y++)
    for (int f = first_frame; f < first_frame + nframes; f++)
      for (int x = 0; x < W; x++) {
        uint32_t seed;
        float3 o, d;
        camera_ray(sc.camera, W, H, x, y, f, seed, o, d);
        const float time = sc.camera.time0 + pt_rnd(seed) * (sc.camera.time1 - sc.camera.time0);
        float3 thr = f3(1.f), rad = f3(0.f);
        float last_pdf = 0.f;
        bool spec = false;
        int depth = 0;
        float3 col = f3(0.f);
        bool alive = max_depth > 0;
        while (alive) {
          TraceHit h;
          uint32_t a = 0, b = 0;
          bvh8_closest<false>(sc, o, d, depth == 0 ? 1e-6f : 1e-3f, PT_RAY_MAX, h, a, b, RayAux{time, seed});
          PathIn in;
          in.o = o; in.d = d; in.thr = thr; in.rad = rad; in.seed = seed; in.depth = depth; in.prev_specular = spec; in.last_pdf = last_pdf; in.time = time;
          if (h.leaf < 0) {
            col = shade_miss(sc, in);
            break;
          }
          int mt = (int)(pt_f2u(sc.leaves[h.leaf].c.w) >> 24);
          ShadeOut out;
          shade_hit_any(sc, mt, in, h.t, h.leaf, h.u, h.v, out);
          if (out.kind == SH_TERMINATE) {
            col = out.col;
            break;
          }
          float3 nd, nthr, nrad;
          uint32_t nseed;
          bool nspec = out.specular;
          if (out.kind == SH_SHADOW) {
            TraceHit sh;
            bvh8_closest<false>(sc, out.P, out.WiL, 1e-3f, PT_RAY_MAX, sh, a, b, RayAux{time, out.seedShadow});
            bool vis = out.env_shadow ? sh.leaf < 0
                                      : (sh.leaf >= 0 && (pt_f2u(sc.leaves[sh.leaf].c.w) >> 24) == PT_MAT_DIFFUSE_LIGHT);
            if (vis) { nd = out.dir; nthr = out.thr; nrad = out.rad; nseed = out.seed; }
            else { nd = out.dirS; nthr = out.thrS; nrad = out.radS; nseed = out.seedS; }
            nspec = false;
            if (out.terminal) {
              col = nrad;
              break;
            }
          } else {
            nd = out.dir; nthr = out.thr; nrad = out.rad; nseed = out.seed;
          }
          // finish_bounce
          if (russian) {
            float prob = pt_ffmax(pt_ffmax(nthr.x, nthr.y), nthr.z);
            if (depth > 10) {
              if (pt_rnd(nseed) >= prob) {
                col = nrad + nthr;
                break;
              }
              nthr *= 1.f / prob;
            }
          }
          depth++;
          if (depth >= max_depth) {
            col = f3(0.f);
            break;
          }
          o = out.P; d = nd; thr = nthr; rad = nrad; seed = nseed; spec = nspec; last_pdf = out.pdf;
        }
        col = pt_de_nan(col);
        float* px = accum + 4 * ((size_t)(H - 1 - y) * W + x);
        px[0] += col.x;
        px[1] += col.y;
        px[2] += col.z;
        px[3] += 1.f;
      }
  return 0;
}

// k_generate's enumeration of the camera samples (csrc/sample_order.h), ids [sbase, sbase + n) of a render of
// `nframes` frames of a W x H band, exactly as the kernel walks them (32-bit offsets from the launch's first id):
// out[3 * i] = x, y, frame (relative to the first frame)
int sim_sample_order(int W, int H, int tiled, int kmax, int nframes, unsigned long long sbase, int n, int* out) {
  const uint32_t npix = (uint32_t)W * (uint32_t)H;
  const uint32_t fb0 = (uint32_t)(sbase / npix);
  const uint32_t rem0 = (uint32_t)(sbase - (unsigned long long)fb0 * npix);
  const SampleOrder so = sample_order_make(W, H, (uint32_t)nframes, kmax, tiled);
  for (int i = 0; i < n; i++) {
    const uint32_t t = rem0 + (uint32_t)i;
    const uint32_t q = t / npix;
    int x, y;
    uint32_t fr;
    sample_order_decode(so, fb0 + q, t - q * npix, x, y, fr);
    out[3 * i] = x;
    out[3 * i + 1] = y;
    out[3 * i + 2] = (int)fr;
  }
  return 0;
}

}  // extern "C"
