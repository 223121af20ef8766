// env_alias.cu — alias tables for HDRI environment importance sampling, built on the GPU at commit.
//
// NEW relative to the reference (its miss program, programs/miss.cu:64-95, only evaluates the map); the
// estimator that consumes the tables is in shading.h (env_sample_dir / env_pdf_dir).  The per-element
// bodies are in env_alias_core.h; this file holds the kernels, the device driver and the serial host
// mirror that answers pt_scene_get_desc for scenes that were never committed.
//
// Pipeline (n = W * H texels):
//   k_env_row_sum    one warp per image row: 32 strided partial sums added in lane order (the summation order is
//                    part of the result and is the same in the host mirror)
//   k_env_total      one thread: total of the row sums in row order, 1 % floor, grand total
//   k_env_pdf        per texel: pdf (float), q = pdf * n in 32.32 fixed point, light flag, deficit / excess
//   cub scans        rank of each item among lights / heavies, deficit and excess prefix sums (uint64: exact)
//   k_env_scatter    compacted light / heavy index and prefix arrays
//   k_env_alias      per texel: closed-form bucket (two binary searches)
#include <cub/cub.cuh>

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "env_alias_core.h"
#include "scene_host.h"

namespace {

struct EnvTotals {
  double total0, floor_w, total;
};

// one warp per image row: lane l sums the texels l, l + 32, ... (coalesced), lane 0 adds the 32 partials in order
__global__ void k_env_row_sum(const float4* __restrict__ texels, int w, int h, const double* __restrict__ row_sin,
                              const EnvTotals* tot, int with_floor, double* __restrict__ row_sums) {
  const int y = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (y >= h) return;
  const double fl = with_floor ? tot->floor_w : 0.0;
  const double part = env_row_partial((const float*)(texels + (size_t)y * w), w, lane, row_sin[y], fl);
  double s = 0.0;
  for (int l = 0; l < PT_ENV_LANES; l++) s += __shfl_sync(0xffffffffu, part, l);
  if (lane == 0) row_sums[y] = s;
}

__global__ void k_env_total(const double* __restrict__ row_sums, int h, int n, int second, EnvTotals* tot) {
  double s = 0.0;
  for (int y = 0; y < h; y++) s += row_sums[y];
  if (!second) {
    tot->total0 = s;
    tot->floor_w = s > 0.0 ? 0.01 * s / (double)n : 1.0;
  } else {
    tot->total = s;
  }
}

__global__ void k_env_pdf(const float4* __restrict__ texels, int w, int n, const double* __restrict__ row_sin,
                          const EnvTotals* tot, float* __restrict__ pdf, unsigned long long* __restrict__ q,
                          int* __restrict__ is_light, unsigned long long* __restrict__ deficit,
                          unsigned long long* __restrict__ excess) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long ONE = 1ull << 32;
  const float p = env_texel_pdf((const float*)(texels + i), row_sin[i / w], tot->floor_w, tot->total);
  const unsigned long long qi = env_q_fixed(p, n);
  pdf[i] = p;
  q[i] = qi;
  is_light[i] = qi <= ONE;
  deficit[i] = qi <= ONE ? ONE - qi : 0ull;
  excess[i] = qi <= ONE ? 0ull : qi - ONE;
}

__global__ void k_env_scatter(int n, const unsigned long long* __restrict__ q, const int* __restrict__ light_rank,
                              const unsigned long long* __restrict__ df_excl, const unsigned long long* __restrict__ ex_incl,
                              int* __restrict__ rank, int* __restrict__ light_idx, int* __restrict__ heavy_idx,
                              unsigned long long* __restrict__ light_df, unsigned long long* __restrict__ heavy_ex,
                              int* __restrict__ n_light) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long ONE = 1ull << 32;
  const int lr = light_rank[i];
  if (q[i] <= ONE) {
    rank[i] = lr;
    light_idx[lr] = i;
    light_df[lr] = df_excl[i];
  } else {
    const int hr = i - lr;
    rank[i] = hr;
    heavy_idx[hr] = i;
    heavy_ex[hr] = ex_incl[i];
  }
  if (i == n - 1) {
    const int a = lr + (q[i] <= ONE ? 1 : 0);
    *n_light = a;
    light_df[a] = df_excl[i] + (q[i] <= ONE ? ONE - q[i] : 0ull);  // total deficit
  }
}

__global__ void k_env_alias(EnvAliasIn in, const int* n_light, float* __restrict__ prob, int* __restrict__ alias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in.n) return;
  in.a = *n_light;
  in.b = in.n - in.a;
  float p;
  int al;
  env_alias_item(in, i, p, al);
  prob[i] = p;
  alias[i] = al;
}

bool env_wanted(const pt_scene* s) {
  const pt_miss& m = s->miss;
  return m.env_nee && m.type == PT_MISS_HDR && m.spherical && m.image >= 0 && m.image < (int)s->images.size();
}

std::vector<double> row_sines(int h) {
  std::vector<double> rs((size_t)h);
  for (int y = 0; y < h; y++) rs[y] = sin(3.14159265358979323846 * (y + 0.5) / h);
  return rs;
}

// serial host mirror of the device pipeline: same bodies, same order of operations
void env_tables_host(pt_scene* s) {
  s->env_alias_prob.clear();
  s->env_alias_idx.clear();
  s->env_texel_pdf.clear();
  if (!env_wanted(s)) return;
  const HostImage& im = s->images[s->miss.image];
  const int w = im.width, h = im.height, n = w * h;
  const std::vector<double> rs = row_sines(h);
  double total0 = 0.0;
  for (int y = 0; y < h; y++) total0 += env_row_sum(im.rgba.data() + 4 * (size_t)y * w, w, rs[y], 0.0);
  const double floor_w = total0 > 0.0 ? 0.01 * total0 / (double)n : 1.0;
  double total = 0.0;
  for (int y = 0; y < h; y++) total += env_row_sum(im.rgba.data() + 4 * (size_t)y * w, w, rs[y], floor_w);
  const unsigned long long ONE = 1ull << 32;
  std::vector<unsigned long long> q((size_t)n), ldf, hex;
  std::vector<int> rank((size_t)n), lidx, hidx;
  s->env_texel_pdf.resize((size_t)n);
  unsigned long long df = 0, ex = 0;
  for (int i = 0; i < n; i++) {
    const float p = env_texel_pdf(im.rgba.data() + 4 * (size_t)i, rs[i / w], floor_w, total);
    s->env_texel_pdf[i] = p;
    q[i] = env_q_fixed(p, n);
    if (q[i] <= ONE) {
      rank[i] = (int)lidx.size();
      lidx.push_back(i);
      ldf.push_back(df);
      df += ONE - q[i];
    } else {
      ex += q[i] - ONE;
      rank[i] = (int)hidx.size();
      hidx.push_back(i);
      hex.push_back(ex);
    }
  }
  ldf.push_back(df);
  EnvAliasIn in{q.data(), rank.data(), lidx.data(), hidx.data(), ldf.data(), hex.data(), n, (int)lidx.size(), (int)hidx.size()};
  s->env_alias_prob.resize((size_t)n);
  s->env_alias_idx.resize((size_t)n);
  for (int i = 0; i < n; i++) env_alias_item(in, i, s->env_alias_prob[i], s->env_alias_idx[i]);
}

}  // namespace

int pt_env_tables_build(pt_scene* s) {
  s->env_tables_valid = false;
  s->env_alias_prob.clear();
  s->env_alias_idx.clear();
  s->env_texel_pdf.clear();
  SceneDev& d = s->dev;
  d.env_n = d.env_w = d.env_h = 0;
  if (!env_wanted(s)) {
    PT_CUDA(s->d_env_prob.alloc(16));
    PT_CUDA(s->d_env_idx.alloc(16));
    PT_CUDA(s->d_env_pdf.alloc(16));
    return PT_OK;
  }
  const HostImage& im = s->images[s->miss.image];
  const int w = im.width, h = im.height;
  if ((long long)w * h > (1ll << 30)) {
    pt_set_error("environment image too large for importance sampling");
    return PT_ERR_INVALID;
  }
  const int n = w * h;
  const float4* texels = s->d_image_texels[s->miss.image]->as<float4>();  // uploaded by commit just before
  const std::vector<double> rs = row_sines(h);
  DevBuf d_rs, d_rows, d_tot, d_q, d_flag, d_def, d_exc, d_lrank, d_dfx, d_exi, d_rank, d_lidx, d_hidx, d_ldf, d_hex, d_na, d_tmp;
  PT_CUDA(d_rs.alloc((size_t)h * 8));
  PT_CUDA(cudaMemcpy(d_rs.p, rs.data(), (size_t)h * 8, cudaMemcpyHostToDevice));
  PT_CUDA(d_rows.alloc((size_t)h * 8));
  PT_CUDA(d_tot.alloc(sizeof(EnvTotals)));
  PT_CUDA(cudaMemset(d_tot.p, 0, sizeof(EnvTotals)));
  PT_CUDA(s->d_env_prob.alloc((size_t)n * 4));
  PT_CUDA(s->d_env_idx.alloc((size_t)n * 4));
  PT_CUDA(s->d_env_pdf.alloc((size_t)n * 4));
  PT_CUDA(d_q.alloc((size_t)n * 8));
  PT_CUDA(d_flag.alloc((size_t)n * 4));
  PT_CUDA(d_def.alloc((size_t)n * 8));
  PT_CUDA(d_exc.alloc((size_t)n * 8));
  PT_CUDA(d_lrank.alloc((size_t)n * 4));
  PT_CUDA(d_dfx.alloc((size_t)n * 8));
  PT_CUDA(d_exi.alloc((size_t)n * 8));
  PT_CUDA(d_rank.alloc((size_t)n * 4));
  PT_CUDA(d_lidx.alloc((size_t)n * 4));
  PT_CUDA(d_hidx.alloc((size_t)n * 4));
  PT_CUDA(d_ldf.alloc(((size_t)n + 1) * 8));
  PT_CUDA(d_hex.alloc((size_t)n * 8));
  PT_CUDA(d_na.alloc(16));
  const bool timing = getenv("PTB200_BUILD_TIMING") != nullptr;
  cudaEvent_t te[4] = {};
  if (timing)
    for (auto& e : te) cudaEventCreate(&e);
  if (timing) cudaEventRecord(te[0]);
  const int TB = 256;
  const int gb = (n + TB - 1) / TB, gr = (h * 32 + 127) / 128;  // row sums: one warp per row, 4 rows per block
  EnvTotals* tot = d_tot.as<EnvTotals>();
  k_env_row_sum<<<gr, 128>>>(texels, w, h, d_rs.as<double>(), tot, 0, d_rows.as<double>());
  k_env_total<<<1, 1>>>(d_rows.as<double>(), h, n, 0, tot);
  k_env_row_sum<<<gr, 128>>>(texels, w, h, d_rs.as<double>(), tot, 1, d_rows.as<double>());
  k_env_total<<<1, 1>>>(d_rows.as<double>(), h, n, 1, tot);
  k_env_pdf<<<gb, TB>>>(texels, w, n, d_rs.as<double>(), tot, s->d_env_pdf.as<float>(), d_q.as<unsigned long long>(),
                        d_flag.as<int>(), d_def.as<unsigned long long>(), d_exc.as<unsigned long long>());
  PT_CUDA(cudaGetLastError());
  if (timing) cudaEventRecord(te[1]);
  size_t t1 = 0, t2 = 0, t3 = 0;
  PT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t1, d_flag.as<int>(), d_lrank.as<int>(), n));
  PT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, d_def.as<unsigned long long>(), d_dfx.as<unsigned long long>(), n));
  PT_CUDA(cub::DeviceScan::InclusiveSum(nullptr, t3, d_exc.as<unsigned long long>(), d_exi.as<unsigned long long>(), n));
  size_t tmp = std::max(t1, std::max(t2, t3));
  PT_CUDA(d_tmp.alloc(tmp));
  PT_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp, d_flag.as<int>(), d_lrank.as<int>(), n));
  PT_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp, d_def.as<unsigned long long>(), d_dfx.as<unsigned long long>(), n));
  PT_CUDA(cub::DeviceScan::InclusiveSum(d_tmp.p, tmp, d_exc.as<unsigned long long>(), d_exi.as<unsigned long long>(), n));
  if (timing) cudaEventRecord(te[2]);
  k_env_scatter<<<gb, TB>>>(n, d_q.as<unsigned long long>(), d_lrank.as<int>(), d_dfx.as<unsigned long long>(),
                            d_exi.as<unsigned long long>(), d_rank.as<int>(), d_lidx.as<int>(), d_hidx.as<int>(),
                            d_ldf.as<unsigned long long>(), d_hex.as<unsigned long long>(), d_na.as<int>());
  EnvAliasIn in{d_q.as<unsigned long long>(), d_rank.as<int>(), d_lidx.as<int>(), d_hidx.as<int>(),
                d_ldf.as<unsigned long long>(), d_hex.as<unsigned long long>(), n, 0, 0};
  k_env_alias<<<gb, TB>>>(in, d_na.as<int>(), s->d_env_prob.as<float>(), s->d_env_idx.as<int>());
  PT_CUDA(cudaGetLastError());
  if (timing) {
    cudaEventRecord(te[3]);
    cudaEventSynchronize(te[3]);
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, te[0], te[1]);
    cudaEventElapsedTime(&b, te[1], te[2]);
    cudaEventElapsedTime(&c, te[2], te[3]);
    fprintf(stderr, "[env tables] weights + pdf %.3f ms, scans %.3f ms, scatter + alias %.3f ms\n", a, b, c);
    for (auto& e : te) cudaEventDestroy(e);
  }
  PT_CUDA(cudaStreamSynchronize(0));  // the scratch buffers above are released when this function returns
  d.env_n = n;
  d.env_w = w;
  d.env_h = h;
  return PT_OK;
}

int pt_env_tables_fetch(pt_scene* s) {
  if (s->committed && s->device >= 0) {
    DeviceGuard g(s->device);
    const size_t n = (size_t)s->dev.env_n;
    s->env_alias_prob.resize(n);
    s->env_alias_idx.resize(n);
    s->env_texel_pdf.resize(n);
    if (n) {
      PT_CUDA(cudaMemcpy(s->env_alias_prob.data(), s->d_env_prob.p, n * 4, cudaMemcpyDeviceToHost));
      PT_CUDA(cudaMemcpy(s->env_alias_idx.data(), s->d_env_idx.p, n * 4, cudaMemcpyDeviceToHost));
      PT_CUDA(cudaMemcpy(s->env_texel_pdf.data(), s->d_env_pdf.p, n * 4, cudaMemcpyDeviceToHost));
    }
  } else {
    // never committed (no device involved): the serial mirror of the same bodies, for the scene description the
    // oracle and the host simulation read.  Nothing on the render path reads these host copies.
    env_tables_host(s);
  }
  s->env_tables_valid = true;
  return PT_OK;
}
