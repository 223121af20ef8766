// env_alias_core.h — per-element bodies of the alias-table construction for HDRI environment importance
// sampling (pt_miss.env_nee; base program programs/miss.cu:64-95, whose radiance 2 x texel is the target).
// PT_HD so that the device kernels (env_alias.cu) and the host description path (pt_scene_get_desc of a scene
// that was never committed: CPU tests feeding the oracle) run the same arithmetic and produce the same tables
// bit for bit.
//
// Walker/Vose tables are usually built with two stacks, which is sequential.  Here the construction is the
// "sweep" formulation, which has a closed form per item and is therefore one independent thread per texel:
//   q_i    = pdf_i * n in 32.32 fixed point (ONE = 2^32).  Integer arithmetic: prefix sums are exact,
//            associative and reproducible (a floating-point scan is neither across launches)
//   light  = items with q <= ONE, in index order l_0..l_{a-1};  heavy = the others, h_0..h_{b-1}
//   DF[r]  = sum of the deficits (ONE - q) of the lights before l_r   (exclusive prefix, r = 0..a)
//   EX[j]  = sum of the excesses (q - ONE) of the heavies up to and including h_j (inclusive prefix)
// The sweep fills light buckets in order from the current heavy item and moves on to the next heavy item once
// the current one's residual drops to <= ONE.  After filling r lights with j heavies used up the residual of
// h_j is ONE + EX[j] - DF[r], so
//   light l_r  keeps q, aliases to h_j with   j  = min { j : EX[j] >  DF[r] }      (upper bound in EX)
//   heavy h_j  keeps ONE + EX[j] - DF[r*],    r* = min { r : DF[r] >= EX[j] }      (lower bound in DF)
//              and aliases to h_{j+1}
// Items the sums do not reach (rounding residue of the float pdfs) keep probability 1.
#pragma once
#include "hd.h"

#define PT_ENV_ONE 4294967296.0  // 2^32

// luminance weight of one texel: the miss program returns 2 x texel (miss.cu:93); rows shrink with sin(theta)
PT_HD double env_texel_weight(const float* rgba, double row_sin) {
  double lum = 2.0 * (0.2126 * (double)rgba[0] + 0.7152 * (double)rgba[1] + 0.0722 * (double)rgba[2]);
  if (!(lum > 0.0) || !(lum < 1.0e300)) lum = 0.0;
  return lum * row_sin;
}

// Sum of one image row.  The order of the additions is part of the result (device and host mirror must agree bit
// for bit), and it is chosen so that a warp can do it with coalesced loads: 32 partial sums, partial l over the
// texels x = l, l + 32, l + 64, ... in increasing x, then the partials added in order 0..31.
#define PT_ENV_LANES 32
PT_HD double env_row_partial(const float* rgba_row, int w, int lane, double row_sin, double floor_w) {
  double s = 0.0;
  for (int x = lane; x < w; x += PT_ENV_LANES) s += env_texel_weight(rgba_row + 4 * (size_t)x, row_sin) + floor_w * row_sin;
  return s;
}
PT_HD double env_row_sum(const float* rgba_row, int w, double row_sin, double floor_w) {  // serial form (host mirror)
  double s = 0.0;
  for (int l = 0; l < PT_ENV_LANES; l++) s += env_row_partial(rgba_row, w, l, row_sin, floor_w);
  return s;
}

PT_HD float env_texel_pdf(const float* rgba, double row_sin, double floor_w, double total) {
  return (float)((env_texel_weight(rgba, row_sin) + floor_w * row_sin) / total);
}

// pdf (float) -> q in 32.32 fixed point.  pdf * n is exact in double (24 + 31 bits), the scaling is a power of two
PT_HD unsigned long long env_q_fixed(float pdf, int n) {
  const double q = (double)pdf * (double)n * PT_ENV_ONE;
  return q >= 0.0 ? (unsigned long long)q : 0ull;
}

PT_HD int env_upper_bound(const unsigned long long* a, int n, unsigned long long key) {  // first i with a[i] > key
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] > key) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}
PT_HD int env_lower_bound(const unsigned long long* a, int n, unsigned long long key) {  // first i with a[i] >= key
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] >= key) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

struct EnvAliasIn {
  const unsigned long long* q;        // n: fixed-point q per item
  const int* rank;                    // n: rank of the item among the lights / among the heavies
  const int* light_idx;               // a: item index of light r
  const int* heavy_idx;               // b: item index of heavy j
  const unsigned long long* light_df; // a + 1: exclusive deficit prefix (entry a = total deficit)
  const unsigned long long* heavy_ex; // b: inclusive excess prefix
  int n, a, b;
};

PT_HD void env_alias_item(const EnvAliasIn& in, int i, float& prob, int& alias) {
  const unsigned long long ONE = 1ull << 32;
  const unsigned long long q = in.q[i];
  prob = 1.f;
  alias = i;
  if (q <= ONE) {
    const int r = in.rank[i];
    const int j = env_upper_bound(in.heavy_ex, in.b, in.light_df[r]);
    if (j < in.b) {
      prob = (float)((double)q / PT_ENV_ONE);
      alias = in.heavy_idx[j];
    }
  } else {
    const int j = in.rank[i];
    const unsigned long long ex = in.heavy_ex[j];
    const int r = env_lower_bound(in.light_df, in.a + 1, ex);
    if (r <= in.a && j + 1 < in.b) {
      prob = (float)((double)(ONE + ex - in.light_df[r]) / PT_ENV_ONE);
      alias = in.heavy_idx[j + 1];
    }
  }
}
