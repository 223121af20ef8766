// multi.cu — multi-GPU below the C ABI: sample split, scene replicated per device, ONE collective.
//
// The reference drives one OptiX context on one GPU (main.cpp:30-33) with one launch per sample
// (main.cpp:264-293).  Samples are independent (seed = tea<64>(W*y+x, frame), programs/raygen.cu:150) and the
// accumulation is a sum (raygen.cu:166): device g of G renders a contiguous share of the frames into its own
// float4 buffer and the buffers are merged by one ncclReduce (sum) over NVLink.  Nothing else crosses devices.
//
// Two forms (include/ptb200.h):
//   pt_renderer_create_multi   one process, one renderer + one host thread per device behind ONE handle
//   pt_comm_*                  one process per device (torchrun / MPI): an NCCL communicator from a unique id
// NCCL is bound at run time with dlopen("libnccl.so.2") - inside a PyTorch process that is the copy torch has
// already loaded - so libptb200.so itself loads on machines without NCCL (every single-GPU path, the CPU-side
// ABI tests).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "renderer_internal.h"
#include "scene_host.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

const NcclApi* nccl() {
  static NcclApi api = [] {
    NcclApi a;
    const char* names[] = {getenv("PTB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      a.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);  // the copy this process already has (torch's)
      if (!a.lib) a.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) return a;
#define PT_NCCL_SYM(field, name) *(void**)(&a.field) = dlsym(a.lib, name)
    PT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    PT_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    PT_NCCL_SYM(CommInitAll, "ncclCommInitAll");
    PT_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    PT_NCCL_SYM(Reduce, "ncclReduce");
    PT_NCCL_SYM(GroupStart, "ncclGroupStart");
    PT_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    PT_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef PT_NCCL_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.Reduce && a.GroupStart && a.GroupEnd &&
           a.GetErrorString;
    return a;
  }();
  if (!api.ok) {
    pt_set_error("NCCL is not available (libnccl.so.2 could not be loaded): multi-GPU needs it");
    return nullptr;
  }
  return &api;
}

#define PT_NCCL(api, call)                                                                          \
  do {                                                                                              \
    ncclResult_t _r = (call);                                                                       \
    if (_r != ncclSuccess) {                                                                        \
      pt_set_error("%s failed: %s (%s:%d)", #call, (api)->GetErrorString(_r), __FILE__, __LINE__);  \
      return PT_ERR_CUDA;                                                                           \
    }                                                                                               \
  } while (0)

// runs f(i) for i in [0, n) on one host thread each (the library's calls are synchronous, so concurrency across
// devices needs one thread per device); returns the first failure with its thread-local message carried over
template <class F>
int for_each_device(int n, F f) {
  std::vector<int> rc((size_t)n, PT_OK);
  std::vector<std::string> msg((size_t)n);
  std::vector<std::thread> th;
  for (int i = 1; i < n; i++)
    th.emplace_back([&, i] {
      rc[i] = f(i);
      if (rc[i]) msg[i] = pt_last_error();
    });
  rc[0] = f(0);
  if (rc[0]) msg[0] = pt_last_error();
  for (auto& t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rc[i]) {
      pt_set_error("device slot %d: %s", i, msg[i].c_str());
      return rc[i];
    }
  return PT_OK;
}

}  // namespace

struct pt_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

struct pt_multi {
  pt_scene* scene = nullptr;  // the caller's scene: replica 0, not owned
  int W = 0, H = 0;
  std::vector<int> devices;
  std::vector<pt_scene*> replicas;  // [0] == scene; the others are owned copies committed on their device
  std::vector<pt_renderer*> subs;
  std::vector<ncclComm_t> comms;
  unsigned long long seen_commit_gen = 0;
  bool merged = true;  // the root buffer holds the sum, the others are zero
  std::vector<float> rate;  // frames per ms of every device in the previous render (0 = unknown): the next split follows it
  float ms_reduce = 0.f;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // on the root device
};

namespace {

int sync_replica(pt_multi* m, int i) {  // host data of the caller's scene -> replica i, committed on device i
  PT_CUDA(cudaSetDevice(m->devices[i]));
  pt_scene* rep = m->replicas[i];
  static_cast<SceneHostData&>(*rep) = static_cast<const SceneHostData&>(*m->scene);
  rep->env_tables_valid = false;
  return pt_scene_commit(rep);
}

// the replicas follow the caller's scene: a new commit rebuilds them, camera / miss changes are copied
int refresh_replicas(pt_multi* m) {
  pt_scene* s = m->scene;
  if (!s->committed) {
    pt_set_error("pt_render: the scene was modified after the last successful pt_scene_commit");
    return PT_ERR_STATE;
  }
  const int n = (int)m->devices.size();
  if (s->commit_gen != m->seen_commit_gen) {
    int rc = for_each_device(n, [&](int i) { return i == 0 ? PT_OK : sync_replica(m, i); });
    if (rc) return rc;
    m->seen_commit_gen = s->commit_gen;
  }
  for (int i = 1; i < n; i++) {
    pt_scene* rep = m->replicas[i];
    if (memcmp(&rep->camera, &s->camera, sizeof(pt_camera)) != 0) {
      int rc = pt_scene_set_camera(rep, &s->camera);
      if (rc) return rc;
    }
    if (memcmp(&rep->miss, &s->miss, sizeof(pt_miss)) != 0) {
      int rc = pt_scene_set_miss(rep, &s->miss);
      if (rc) return rc;
    }
    if (!rep->committed) {  // a camera / miss change that needs a rebuild (shutter interval, env tables)
      DeviceGuard g(m->devices[i]);
      int rc = pt_scene_commit(rep);
      if (rc) return rc;
    }
  }
  return PT_OK;
}

}  // namespace

int pt_multi_destroy(pt_multi* m) {
  if (!m) return PT_OK;
  const NcclApi* api = m->comms.empty() ? nullptr : nccl();
  for (size_t i = 0; i < m->subs.size(); i++)
    if (m->subs[i]) pt_renderer_destroy(m->subs[i]);
  for (size_t i = 0; i < m->comms.size(); i++)
    if (m->comms[i] && api) api->CommDestroy(m->comms[i]);
  for (size_t i = 1; i < m->replicas.size(); i++) {
    if (!m->replicas[i]) continue;
    DeviceGuard g(m->devices[i]);
    delete m->replicas[i];
  }
  {
    DeviceGuard g(m->devices.empty() ? -1 : m->devices[0]);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
  }
  delete m;
  return PT_OK;
}

pt_renderer* pt_multi_root(pt_multi* m) { return m->subs[0]; }

int pt_multi_set_flags(pt_multi* m, int32_t flags) {
  for (pt_renderer* r : m->subs) pt_renderer_set_flags(r, flags);
  return PT_OK;
}
int pt_multi_set_pool(pt_multi* m, int64_t paths) {
  for (pt_renderer* r : m->subs) pt_renderer_set_pool(r, paths);
  return PT_OK;
}
int pt_multi_sync(pt_multi* m) {
  for (pt_renderer* r : m->subs) {
    int rc = pt_renderer_sync(r);
    if (rc) return rc;
  }
  return PT_OK;
}

int pt_multi_render(pt_multi* m, int32_t first_frame, int32_t nframes, int32_t max_depth, int32_t russian,
                    int32_t clear_accum) {
  int rc = refresh_replicas(m);
  if (rc) return rc;
  const int n = (int)m->devices.size();
  // GPUs of one box differ by 2-3 % in speed and a split ends when its slowest device does: shares follow the rates
  // measured in the previous render (equal shares the first time)
  std::vector<int32_t> share_n((size_t)n, 0);
  rc = for_each_device(n, [&](int i) {
    int32_t f0 = 0, fn = 0;
    pt_split_frames_weighted(nframes, i, n, m->rate.size() == (size_t)n ? m->rate.data() : nullptr, first_frame, &f0, &fn);
    share_n[i] = fn;
    // every buffer starts from zero (creation, clear, or the zeroing that follows a merge); the reference's
    // "frame 0 clears" rule (raygen.cu:154) is applied by the renderer that owns frame 0
    return pt_render(m->subs[i], f0, fn, max_depth, russian, clear_accum);
  });
  if (rc) return rc;
  m->rate.assign((size_t)n, 0.f);
  for (int i = 0; i < n; i++) {
    pt_stats st{};
    pt_get_stats(m->subs[i], &st);
    m->rate[i] = share_n[i] > 0 && st.ms_total > 0.f ? (float)share_n[i] / st.ms_total : 0.f;
  }
  m->merged = false;
  return PT_OK;
}

int pt_multi_merge(pt_multi* m) {
  if (m->merged) return PT_OK;
  if (m->devices.size() == 1) {  // nothing to merge, and no NCCL needed
    m->merged = true;
    m->ms_reduce = 0.f;
    return pt_multi_sync(m);
  }
  const NcclApi* api = nccl();
  if (!api) return PT_ERR_STATE;
  const int n = (int)m->devices.size();
  const size_t count = (size_t)m->W * m->H * 4;
  {
    DeviceGuard g(m->devices[0]);
    PT_CUDA(cudaEventRecord(m->ev0, pt_r_stream(m->subs[0])));
  }
  PT_NCCL(api, api->GroupStart());
  for (int i = 0; i < n; i++) {
    float4* buf = pt_r_accum(m->subs[i]);
    ncclResult_t r = api->Reduce(buf, buf, count, ncclFloat, ncclSum, 0, m->comms[i], pt_r_stream(m->subs[i]));
    if (r != ncclSuccess) {
      api->GroupEnd();
      pt_set_error("ncclReduce failed: %s", api->GetErrorString(r));
      return PT_ERR_CUDA;
    }
  }
  PT_NCCL(api, api->GroupEnd());
  {
    DeviceGuard g(m->devices[0]);
    PT_CUDA(cudaEventRecord(m->ev1, pt_r_stream(m->subs[0])));
  }
  for (int i = 1; i < n; i++) {  // their samples now live in the root buffer
    DeviceGuard g(m->devices[i]);
    PT_CUDA(cudaMemsetAsync(pt_r_accum(m->subs[i]), 0, count * 4, pt_r_stream(m->subs[i])));
  }
  int rc = pt_multi_sync(m);
  if (rc) return rc;
  {
    DeviceGuard g(m->devices[0]);
    PT_CUDA(cudaEventElapsedTime(&m->ms_reduce, m->ev0, m->ev1));
  }
  m->merged = true;
  return PT_OK;
}

int pt_multi_write_accum(pt_multi* m, const float* host_rgba) {
  int rc = pt_write_accum(m->subs[0], host_rgba);
  if (rc) return rc;
  const size_t bytes = (size_t)m->W * m->H * 16;
  for (size_t i = 1; i < m->subs.size(); i++) {
    DeviceGuard g(m->devices[i]);
    PT_CUDA(cudaMemsetAsync(pt_r_accum(m->subs[i]), 0, bytes, pt_r_stream(m->subs[i])));
    PT_CUDA(cudaStreamSynchronize(pt_r_stream(m->subs[i])));
    pt_r_mark_started(m->subs[i]);
  }
  m->merged = true;
  return PT_OK;
}

int pt_multi_get_stats(pt_multi* m, pt_stats* out) {
  pt_stats t{};
  for (pt_renderer* r : m->subs) {
    pt_stats s{};
    pt_get_stats(r, &s);
    t.samples += s.samples;
    t.closest_rays += s.closest_rays;
    t.shadow_rays += s.shadow_rays;
    t.nodes_fetched += s.nodes_fetched;
    t.leaves_fetched += s.leaves_fetched;
    t.kernel_launches += s.kernel_launches;
    t.iterations = std::max(t.iterations, s.iterations);
    // the devices run concurrently: the render took as long as the slowest of them
    t.ms_total = std::max(t.ms_total, s.ms_total);
    t.ms_trace_closest = std::max(t.ms_trace_closest, s.ms_trace_closest);
    t.ms_trace_shadow = std::max(t.ms_trace_shadow, s.ms_trace_shadow);
    t.ms_shade = std::max(t.ms_shade, s.ms_shade);
    t.ms_generate = std::max(t.ms_generate, s.ms_generate);
    t.drained_paths |= s.drained_paths;
  }
  t.ms_reduce = m->ms_reduce;
  *out = t;
  return PT_OK;
}

extern "C" {

int pt_split_frames(int32_t total_frames, int32_t rank, int32_t world, int32_t first_frame, int32_t* out_first,
                    int32_t* out_n) {
  if (total_frames < 0 || world <= 0 || rank < 0 || rank >= world || !out_first || !out_n) {
    pt_set_error("pt_split_frames: bad arguments");
    return PT_ERR_INVALID;
  }
  const int32_t base = total_frames / world, rem = total_frames % world;
  *out_n = base + (rank < rem ? 1 : 0);
  *out_first = first_frame + rank * base + (rank < rem ? rank : rem);
  return PT_OK;
}

int pt_split_frames_weighted(int32_t total_frames, int32_t rank, int32_t world, const float* weights, int32_t first_frame,
                             int32_t* out_first, int32_t* out_n) {
  if (total_frames < 0 || world <= 0 || rank < 0 || rank >= world || !out_first || !out_n) {
    pt_set_error("pt_split_frames_weighted: bad arguments");
    return PT_ERR_INVALID;
  }
  double sum = 0.0;
  bool usable = weights != nullptr;
  for (int i = 0; usable && i < world; i++) {
    if (!(weights[i] > 0.f) || !(weights[i] < 1e30f)) usable = false;
    else sum += (double)weights[i];
  }
  if (!usable) return pt_split_frames(total_frames, rank, world, first_frame, out_first, out_n);
  // largest-remainder apportionment: floor of the exact share, the leftover frames to the largest fractions
  // (ties to the lower rank)
  std::vector<int32_t> n((size_t)world);
  std::vector<double> frac((size_t)world);
  int32_t given = 0;
  for (int i = 0; i < world; i++) {
    const double share = (double)total_frames * (double)weights[i] / sum;
    n[i] = (int32_t)share;
    frac[i] = share - (double)n[i];
    given += n[i];
  }
  for (int32_t left = total_frames - given; left > 0; left--) {
    int best = 0;
    for (int i = 1; i < world; i++)
      if (frac[i] > frac[best]) best = i;
    n[best]++;
    frac[best] = -1.0;
  }
  int32_t start = first_frame;
  for (int i = 0; i < rank; i++) start += n[i];
  *out_first = start;
  *out_n = n[rank];
  return PT_OK;
}

int pt_renderer_create_multi(pt_scene* scene, int32_t width, int32_t height, const int32_t* device_ids, int32_t n_devices,
                             pt_renderer** out) {
  if (!scene || !out || !device_ids || n_devices <= 0 || n_devices > PT_MAX_DEVICES || width <= 0 || height <= 0) {
    pt_set_error("pt_renderer_create_multi: bad arguments");
    return PT_ERR_INVALID;
  }
  if (!scene->committed) {
    pt_set_error("pt_renderer_create_multi: scene not committed");
    return PT_ERR_STATE;
  }
  int ndev = 0;
  PT_CUDA(cudaGetDeviceCount(&ndev));
  for (int i = 0; i < n_devices; i++) {
    if (device_ids[i] < 0 || device_ids[i] >= ndev) {
      pt_set_error("pt_renderer_create_multi: device %d does not exist (%d devices visible)", device_ids[i], ndev);
      return PT_ERR_INVALID;
    }
    for (int j = 0; j < i; j++)
      if (device_ids[j] == device_ids[i]) {
        pt_set_error("pt_renderer_create_multi: device %d listed twice", device_ids[i]);
        return PT_ERR_INVALID;
      }
  }
  int root_pos = -1;
  for (int i = 0; i < n_devices; i++)
    if (device_ids[i] == scene->device) root_pos = i;
  if (root_pos < 0) {
    pt_set_error("pt_renderer_create_multi: the scene was committed on device %d, which is not in the device list",
                 scene->device);
    return PT_ERR_INVALID;
  }
  const NcclApi* api = n_devices > 1 ? nccl() : nullptr;
  if (n_devices > 1 && !api) return PT_ERR_STATE;

  pt_multi* m = new pt_multi();
  m->scene = scene;
  m->W = width;
  m->H = height;
  m->devices.assign(device_ids, device_ids + n_devices);
  std::swap(m->devices[0], m->devices[root_pos]);  // the scene's own device is the root of the merge
  m->replicas.assign((size_t)n_devices, nullptr);
  m->subs.assign((size_t)n_devices, nullptr);
  m->replicas[0] = scene;
  for (int i = 1; i < n_devices; i++) m->replicas[i] = new pt_scene();
  int rc = for_each_device(n_devices, [&](int i) {
    if (i > 0) {
      int r = sync_replica(m, i);
      if (r) return r;
    }
    return pt_renderer_create(m->replicas[i], width, height, nullptr, &m->subs[i]);
  });
  m->seen_commit_gen = scene->commit_gen;
  if (!rc && n_devices > 1) {
    m->comms.assign((size_t)n_devices, nullptr);
    ncclResult_t r = api->CommInitAll(m->comms.data(), n_devices, m->devices.data());
    if (r != ncclSuccess) {
      pt_set_error("ncclCommInitAll failed: %s", api->GetErrorString(r));
      rc = PT_ERR_CUDA;
    }
  }
  if (!rc) {
    DeviceGuard g(m->devices[0]);
    if (cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess) {
      pt_set_error("pt_renderer_create_multi: event creation failed");
      rc = PT_ERR_CUDA;
    }
  }
  if (rc) {
    std::string keep = pt_last_error();
    pt_multi_destroy(m);
    pt_set_error("%s", keep.c_str());
    return rc;
  }
  *out = pt_r_new_multi_handle(scene, width, height, m);
  return PT_OK;
}

int pt_comm_unique_id(uint8_t id[PT_COMM_ID_BYTES]) {
  const NcclApi* api = nccl();
  if (!api) return PT_ERR_STATE;
  if (!id) return PT_ERR_INVALID;
  static_assert(sizeof(ncclUniqueId) == PT_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  PT_NCCL(api, api->GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return PT_OK;
}

int pt_comm_create(const uint8_t id[PT_COMM_ID_BYTES], int32_t rank, int32_t world, pt_comm** out) {
  if (!id || !out || world <= 0 || rank < 0 || rank >= world) {
    pt_set_error("pt_comm_create: bad arguments");
    return PT_ERR_INVALID;
  }
  const NcclApi* api = nccl();
  if (!api) return PT_ERR_STATE;
  pt_comm* c = new pt_comm();
  c->rank = rank;
  c->world = world;
  if (cudaGetDevice(&c->device) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess ||
      cudaEventCreate(&c->ev1) != cudaSuccess) {
    pt_set_error("pt_comm_create: no usable CUDA device");
    delete c;
    return PT_ERR_CUDA;
  }
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
  if (r != ncclSuccess) {
    pt_set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    delete c;
    return PT_ERR_CUDA;
  }
  *out = c;
  return PT_OK;
}

int pt_comm_destroy(pt_comm* c) {
  if (!c) return PT_OK;
  const NcclApi* api = nccl();
  DeviceGuard g(c->device);
  if (api && c->comm) api->CommDestroy(c->comm);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  delete c;
  return PT_OK;
}

int pt_comm_reduce(pt_comm* c, pt_renderer* r, int32_t root) {
  if (!c || !r || root < 0 || root >= c->world) {
    pt_set_error("pt_comm_reduce: bad arguments");
    return PT_ERR_INVALID;
  }
  if (pt_r_multi(r)) {
    pt_set_error("pt_comm_reduce: a multi-device renderer merges by itself (pt_read_accum)");
    return PT_ERR_INVALID;
  }
  if (pt_r_device(r) != c->device) {
    pt_set_error("pt_comm_reduce: the renderer lives on device %d, the communicator on device %d", pt_r_device(r), c->device);
    return PT_ERR_INVALID;
  }
  const NcclApi* api = nccl();
  if (!api) return PT_ERR_STATE;
  DeviceGuard g(c->device);
  cudaStream_t st = pt_r_stream(r);
  float4* buf = pt_r_accum(r);
  const size_t count = (size_t)pt_r_width(r) * pt_r_height(r) * 4;
  PT_CUDA(cudaEventRecord(c->ev0, st));
  PT_NCCL(api, api->Reduce(buf, buf, count, ncclFloat, ncclSum, root, c->comm, st));
  PT_CUDA(cudaEventRecord(c->ev1, st));
  PT_CUDA(cudaEventSynchronize(c->ev1));
  float ms = 0.f;
  PT_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  pt_r_set_ms_reduce(r, ms);
  return PT_OK;
}

}  // extern "C"
