// scene_host.h — host-side scene object behind the opaque pt_scene handle,
// plus the device allocations commit() creates.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "bvh_core.h"
#include "scene_dev.h"

void pt_set_error(const char* fmt, ...);

#define PT_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      pt_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PT_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

struct DevBuf {  // owning device allocation
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { release(); }
  // stream-ordered allocations from the device's default pool (kept warm: repeated scene
  // commits reuse memory instead of paying cudaMalloc/cudaFree every time)
  static void warm_pool() {
    static bool done = false;
    if (done) return;
    done = true;
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  void release() {
    if (p) cudaFreeAsync(p, 0);
    p = nullptr;
    bytes = 0;
  }
  cudaError_t alloc(size_t n) {
    release();
    warm_pool();
    if (n == 0) n = 16;
    cudaError_t e = cudaMallocAsync(&p, n, 0);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <typename T>
  T* as() const { return (T*)p; }
};

struct HostImage {
  int width, height;
  std::vector<float> rgba;
};

struct pt_scene {
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
  std::vector<pt_image> image_descs;
  std::vector<pt_light> lights;
  pt_camera camera{};
  pt_miss miss{};
  bool committed = false;
  unsigned long long commit_gen = 0;  // bumped by every successful commit (renderers key cached graphs on it)
  // environment importance sampling tables (built on the host at commit / get_desc)
  std::vector<float> env_alias_prob, env_texel_pdf;
  std::vector<int32_t> env_alias_idx;
  bool env_tables_valid = false;

  // device side
  DevBuf d_moving_spheres, d_volume_spheres, d_volume_boxes;
  DevBuf d_prim_ref, d_spheres, d_rects, d_boxes, d_tri_v, d_tri_n, d_tri_uv, d_tri_tex, d_xforms;
  DevBuf d_materials, d_textures, d_texvec, d_texdata, d_images, d_lights;
  std::vector<DevBuf*> d_image_texels;
  DevBuf d_nodes8, d_leaves, d_prim_leaf, d_env_prob, d_env_idx, d_env_pdf;
  SceneDev dev{};
  pt_bvh_info info{};
  uint32_t material_type_mask = 0;  // which PT_MAT_* occur (host decides which shade kernels to launch)

  ~pt_scene() {
    for (auto* b : d_image_texels) delete b;
  }
};

// bvh_build.cu
int pt_bvh_build(pt_scene* s);
