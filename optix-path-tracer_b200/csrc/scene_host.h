// scene_host.h — host-side scene object behind the opaque pt_scene handle,
// plus the device allocations commit() creates.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include "bvh_core.h"
#include "scene_dev.h"

void pt_set_error(const char* fmt, ...);

#define PT_MAX_DEVICES 64  // per-device caches (sm counts, slabs, build arenas) are indexed by device ordinal

#define PT_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      pt_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PT_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

// RAII switch of the calling thread's current device (multi-GPU renderers drive several devices)
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

struct DevBuf {  // owning device allocation; remembers the device it lives on
  void* p = nullptr;
  size_t bytes = 0;
  int dev = -1;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  // stream-ordered allocations from the device's default pool (kept warm: repeated scene
  // commits reuse memory instead of paying cudaMalloc/cudaFree every time)
  static void warm_pool(int dev) {
    static std::atomic<unsigned long long> done_mask{0};  // bit per device ordinal (< 64); commits may run on several threads
    if (dev < 0 || dev >= 64 || (done_mask.fetch_or(1ull << dev) >> dev) & 1ull) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  void release() {
    if (p) {
      DeviceGuard g(dev);  // the legacy stream of the owning device
      cudaFreeAsync(p, 0);
    }
    p = nullptr;
    bytes = 0;
    dev = -1;
  }
  cudaError_t alloc(size_t n) {
    release();
    int cur = 0;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return e;
    warm_pool(cur);
    if (n == 0) n = 16;
    e = cudaMallocAsync(&p, n, 0);
    if (e == cudaSuccess) {
      bytes = n;
      dev = cur;
    } else {
      p = nullptr;
    }
    return e;
  }
  template <typename T>
  T* as() const { return (T*)p; }
};

struct HostImage {
  int width, height;
  std::vector<float> rgba;
};

// Everything the host knows about a scene: plain copyable data.  A multi-GPU renderer replicates a scene
// by copying this part and committing the copy on another device (multi.cu).
struct SceneHostData {
  std::vector<pt_prim_ref> prims;
  std::vector<pt_sphere> spheres;
  std::vector<pt_rect> rects;
  std::vector<pt_box> boxes;
  std::vector<pt_moving_sphere> moving_spheres;
  std::vector<pt_volume_sphere> volume_spheres;
  std::vector<pt_volume_box> volume_boxes;
  std::vector<float> tri_v, tri_n, tri_uv;
  std::vector<int32_t> tri_tex;
  bool any_n = false, any_uv = false, any_tex = false;
  std::vector<pt_xform> xforms;
  std::vector<pt_material> materials;
  std::vector<pt_texture> textures;
  std::vector<int32_t> texvec;
  std::vector<float> texdata;
  std::vector<HostImage> images;
  std::vector<pt_light> lights;
  pt_camera camera{};
  pt_miss miss{};
};

struct pt_scene : SceneHostData {
  std::vector<pt_image> image_descs;
  bool committed = false;
  unsigned long long commit_gen = 0;  // bumped by every successful commit (renderers key cached graphs on it)
  // shutter interval the bounds of moving spheres were built for (pt_scene_set_camera re-commits when it changes)
  float built_time0 = 0.f, built_time1 = 0.f;
  // environment importance sampling tables: built on the GPU at commit (env_alias.cu); the host copies are
  // fetched lazily for pt_scene_get_desc (what the oracle reads)
  std::vector<float> env_alias_prob, env_texel_pdf;
  std::vector<int32_t> env_alias_idx;
  bool env_tables_valid = false;  // host copies match the device tables

  // device side
  int device = -1;  // device the scene was committed on
  DevBuf d_moving_spheres, d_volume_spheres, d_volume_boxes;
  DevBuf d_prim_ref, d_spheres, d_rects, d_boxes, d_tri_v, d_tri_n, d_tri_uv, d_tri_tex, d_xforms;
  DevBuf d_materials, d_textures, d_texvec, d_texdata, d_images, d_lights;
  std::vector<DevBuf*> d_image_texels;
  // BVH8 nodes and leaf records live in ONE block, nodes first (one allocation per commit; an L2 access-policy
  // window over it was measured and not kept: profiles/r02_l2_persist_ab.md)
  DevBuf d_bvh;
  size_t bvh_nodes_bytes = 0, bvh_leaves_off = 0;
  DevBuf d_prim_leaf, d_env_prob, d_env_idx, d_env_pdf;
  DevBuf build_arena;  // scratch of the BVH build, kept across commits (grow-only)
  SceneDev dev{};
  pt_bvh_info info{};
  uint32_t material_type_mask = 0;  // which PT_MAT_* occur (host decides which shade kernels to launch)

  ~pt_scene() {
    for (auto* b : d_image_texels) delete b;
  }
};

// bvh_build.cu
int pt_bvh_build(pt_scene* s);
void pt_bvh_release_arenas();
// env_alias.cu: alias tables of the HDR environment image, built on the device (s->d_env_*, s->dev.env_*)
int pt_env_tables_build(pt_scene* s);
int pt_env_tables_fetch(pt_scene* s);  // device tables -> host copies (pt_scene_get_desc)
