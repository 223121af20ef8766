/*
 * ptb200.h — C ABI of the B200-native path tracer (libptb200.so).
 *
 * This is the drop-in boundary for the reference's hot path
 *   ray generation -> BVH traversal (closest hit + shadow) -> material
 *   shading / MIS / Russian roulette -> float4 accumulation.
 * The reference (joaovbs96/OptiX-Path-Tracer) has no FFI layer; its scene code
 * talks to OptiX 6 objects directly.  Every entry point below names the
 * reference interface it replaces (paths relative to
 * OptiX-Path-Tracer/ in the reference tree).
 *
 * Conventions
 *   - plain C, no C++ or torch types; every call returns 0 on success or a
 *     negative pt_status; pt_last_error() gives a thread-local message.
 *   - the caller owns every input array; the library copies before returning.
 *   - the library owns device memory until the matching *_destroy.
 *   - primitive ids are assigned in the order primitives are added
 *     (first id of every batch is returned through `first_prim_id`).
 *   - all arithmetic is FP32; RNG is uint32 (programs/random.cuh).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     PT_ERR_CUDA when no CUDA device is usable.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB200_ABI_VERSION 3

/* fixed limits of the renderer.  The 64-byte path record packs pixel index | depth | specular flag into 32 bits
 * (23 + 8 + 1): larger depths are rejected by pt_render (PT_ERR_INVALID); the reference's default maxDepth is 50
 * (host_includes/host_common.hpp:30).  An image of more than PT_BAND_PIXELS pixels (3840x2160 is 8.3 M, just below)
 * is rendered as horizontal bands of at most that many pixels, one after the other inside pt_render, with the same
 * per-pixel sample set (the seed is tea<64>(W*y + x, frame), raygen.cu:150): up to PT_MAX_PIXELS pixels per image,
 * and at most PT_BAND_PIXELS per row; pt_renderer_create rejects anything larger. */
#define PT_MAX_PIXELS (1 << 28)
#define PT_BAND_PIXELS (1 << 23)
#define PT_MAX_DEPTH 255

typedef enum pt_status {
  PT_OK = 0,
  PT_ERR_INVALID = -1, /* bad argument / bad scene state                     */
  PT_ERR_CUDA = -2,    /* CUDA runtime error or no device                    */
  PT_ERR_OOM = -3,     /* host or device allocation failed                   */
  PT_ERR_STATE = -4    /* call order violated (e.g. render before commit)    */
} pt_status;

/* ---- primitive / material / texture vocabulary (reference names) -------- */

/* programs/common.hpp:26 `typedef enum { X_AXIS, Y_AXIS, Z_AXIS } AXIS;` */
enum { PT_X_AXIS = 0, PT_Y_AXIS = 1, PT_Z_AXIS = 2 };

/* primitive kinds: programs/hitables/{triangle,sphere,aarect,box,moving_sphere,volume_sphere,
 * volume_box}.cu.  Kinds >= PT_PRIM_MOVING_SPHERE read per-ray data during traversal (the
 * sample's shutter time `prd.time`, the path's RNG state `prd.seed`). */
enum {
  PT_PRIM_TRIANGLE = 0,
  PT_PRIM_SPHERE = 1,
  PT_PRIM_RECT = 2,
  PT_PRIM_BOX = 3,
  PT_PRIM_MOVING_SPHERE = 4,
  PT_PRIM_VOLUME_SPHERE = 5,
  PT_PRIM_VOLUME_BOX = 6,
  PT_PRIM_KIND_COUNT = 7
};

/* host_includes/hitables.hpp:73-128 `Sphere(center, radius, BRDF*)` */
typedef struct pt_sphere {
  float center[3];
  float radius;
} pt_sphere;

/* host_includes/hitables.hpp:216-270 `AARect(a0,a1,b0,b1,k,flip,axis,BRDF*)`;
 * axis -> (a,b) mapping: X:(y,z) Y:(x,z) Z:(x,y) (programs/hitables/aarect.cu:20-45) */
typedef struct pt_rect {
  float a0, a1, b0, b1, k;
  int32_t axis; /* PT_*_AXIS */
  int32_t flip; /* flip normal */
} pt_rect;

/* host_includes/hitables.hpp:272-311 `Box(p0, p1, BRDF*)` */
typedef struct pt_box {
  float bmin[3];
  float bmax[3];
} pt_box;

/* host_includes/hitables.hpp:131-173 `Moving_Sphere(c0, c1, t0, t1, r, BRDF*)`,
 * programs/hitables/moving_sphere.cu: centre moves linearly from center0 (time0) to center1
 * (time1); a ray sees it at its sample's shutter time (programs/raygen.cu:69). */
typedef struct pt_moving_sphere {
  float center0[3];
  float center1[3];
  float radius;
  float time0, time1;
} pt_moving_sphere;

/* host_includes/hitables.hpp:175-217 `Volumetric_Sphere(center, radius, density, BRDF*)`,
 * programs/hitables/volume_sphere.cu: constant-density medium bounded by a sphere; the hit is a
 * scattering event at an exponentially distributed distance after the entry point. */
typedef struct pt_volume_sphere {
  float center[3];
  float radius;
  float density;
} pt_volume_sphere;

/* host_includes/hitables.hpp:316-350 `Volumetric_Box(p0, p1, density, BRDF*)`,
 * programs/hitables/volume_box.cu */
typedef struct pt_volume_box {
  float bmin[3];
  float bmax[3];
  float density;
} pt_volume_box;

/* host_includes/materials.hpp:42-242 */
enum {
  PT_MAT_LAMBERTIAN = 0,        /* Lambertian(tex)                              */
  PT_MAT_METAL = 1,             /* Metal(tex, fuzz)             p0=fuzz         */
  PT_MAT_DIELECTRIC = 2,        /* Dielectric(base, ext, ref_idx, density)      */
                                /*   tex0=base tex1=extinction p0=ref_idx p1=density */
  PT_MAT_DIFFUSE_LIGHT = 3,     /* Diffuse_Light(tex)                           */
  PT_MAT_OREN_NAYAR = 4,        /* Oren_Nayar(tex, R)  p0=rA p1=rB (already converted,
                                   host_includes/materials.hpp:192-198)         */
  PT_MAT_TORRANCE_SPARROW = 5,  /* Torrance_Sparrow(tex,nu,nv) p0,p1 = alpha =
                                   max(R,1e-3)^2 (materials.hpp:235-238)        */
  PT_MAT_ASHIKHMIN_SHIRLEY = 6, /* Ashikhmin_Shirley(diff,spec,nu,nv) tex0=diffuse
                                   tex1=specular p0=max(1,nu) p1=max(1,nv)      */
  PT_MAT_NORMAL = 7,            /* Normal_Shader(useShadingNormal) flags bit0   */
  PT_MAT_ISOTROPIC = 8,         /* Isotropic(tex): phase function of the volumes
                                   (materials.hpp:125-141, programs/materials/isotropic.*) */
  PT_MAT_TYPE_COUNT = 9
};

typedef struct pt_material {
  int32_t type;
  int32_t tex0, tex1; /* texture ids, -1 if unused */
  float p0, p1;
  int32_t flags;
} pt_material;

/* host_includes/textures.hpp:16-270, programs/textures/ (all .cu files) */
enum {
  PT_TEX_CONSTANT = 0, /* color                                               */
  PT_TEX_CHECKER = 1,  /* a=odd texture id, b=even texture id                 */
  PT_TEX_IMAGE = 2,    /* a=image id (bilinear, repeat, normalised coords)    */
  PT_TEX_VECTOR = 3,   /* a=first entry in the texture-vector table, b=count  */
  PT_TEX_NOISE = 4,    /* Perlin turbulence (noise_texture.cu): color[0]=scale, a=offset into the
                          texture-data table of 768 gradient floats (ranvec) followed by
                          3x256 permutation entries (perm_x, perm_y, perm_z), b=axis (recorded;
                          the reference's switch falls through, so Z is always used)          */
  PT_TEX_GRADIENT = 5  /* gradient_texture.cu: a=offset of 9 floats (colorA,colorB,colorC)   */
};

typedef struct pt_texture {
  int32_t type;
  float color[3];
  int32_t a, b;
} pt_texture;

/* host_includes/pdfs.hpp:18-80 `Rectangle_PDF(a0,a1,b0,b1,k,axis)` +
 * Light_Sampler::emissions (host_includes/host_common.hpp:65-68) */
/* PT_LIGHT_SPHERE replaces Sphere_PDF (host_includes/pdfs.hpp:82-109, programs/pdfs/sphere_pdf.cu):
 * centre = (a0, a1, b0), radius = k; axis, b1 unused. */
enum { PT_LIGHT_RECT = 0, PT_LIGHT_SPHERE = 1 };

typedef struct pt_light {
  int32_t type;
  int32_t axis;
  float a0, a1, b0, b1, k;
  float emission[3];
} pt_light;

/* host_includes/camera.hpp:13-51: the seven float3 + three floats the raygen
 * program reads (programs/raygen.cu:39-48) */
typedef struct pt_camera {
  float origin[3];
  float lower_left_corner[3];
  float horizontal[3];
  float vertical[3];
  float u[3], v[3], w[3];
  float lens_radius;
  float time0, time1;
} pt_camera;

/* host_includes/programs.hpp:32 `enum Miss_Programs {GRADIENT, CONSTANT, IMG, HDR}`,
 * programs/miss.cu:30-95 */
enum { PT_MISS_GRADIENT = 0, PT_MISS_CONSTANT = 1, PT_MISS_IMG = 2, PT_MISS_HDR = 3 };

typedef struct pt_miss {
  int32_t type;
  float color1[3]; /* GRADIENT: bottom colour; CONSTANT: the colour           */
  float color2[3]; /* GRADIENT: top colour                                    */
  int32_t image;   /* IMG/HDR: image id                                       */
  int32_t spherical; /* HDR: 1 = equirect, 0 = cylindrical (miss.cu:64-95)    */
  int32_t env_nee; /* 0 = reference semantics (no env importance sampling);
                      1 = NEW: alias-table env NEE + MIS (not in reference)   */
} pt_miss;

/* one traversal query; programs/raygen.cu:58-62 `make_Ray(o,d,type,tmin,tmax)` */
typedef struct pt_ray {
  float origin[3];
  float tmin;
  float dir[3];
  float tmax;
} pt_ray;

/* closest-hit answer: rtIntersectionDistance + `geo_index`/`bc` attributes
 * (programs/hitables/triangle.cu:52-56).  prim = -1 on miss. */
typedef struct pt_hit {
  float t;
  int32_t prim;
  float u, v; /* triangle barycentrics, 0 for analytic primitives */
} pt_hit;

/* ---- flat scene description (what commit uploads; what the oracle reads) -- */

typedef struct pt_prim_ref {
  int32_t type;     /* PT_PRIM_*                                             */
  int32_t index;    /* index into the typed array                            */
  int32_t material; /* material id                                           */
  int32_t xform;    /* -1 = identity, else index into xforms                 */
} pt_prim_ref;

/* object->world and world->object, row-major 3x4 (Transform node matrices,
 * host_includes/transforms.hpp:52-64; composed per primitive) */
typedef struct pt_xform {
  float o2w[12];
  float w2o[12];
} pt_xform;

typedef struct pt_image {
  int32_t width, height;
  const float* rgba; /* width*height*4 floats, row 0 first                    */
} pt_image;

typedef struct pt_scene_desc {
  int32_t n_prims;
  const pt_prim_ref* prims;
  int32_t n_spheres;
  const pt_sphere* spheres;
  int32_t n_rects;
  const pt_rect* rects;
  int32_t n_boxes;
  const pt_box* boxes;
  int32_t n_tris;
  const float* tri_v;        /* 9 floats / triangle (a,b,c), world space      */
  const float* tri_n;        /* 9 floats / triangle or NULL                   */
  const float* tri_uv;       /* 6 floats / triangle or NULL                   */
  const int32_t* tri_tex;    /* per-triangle texture index (material_buffer,
                                programs/hitables/triangle.cu:146) or NULL=0  */
  int32_t n_xforms;
  const pt_xform* xforms;
  int32_t n_materials;
  const pt_material* materials;
  int32_t n_textures;
  const pt_texture* textures;
  int32_t n_texvec;
  const int32_t* texvec;     /* texture-vector table (programs/textures/vector.cu) */
  int32_t n_texdata;
  const float* texdata;      /* float table for noise / gradient textures          */
  int32_t n_images;
  const pt_image* images;
  int32_t n_lights;
  const pt_light* lights;
  pt_camera camera;
  pt_miss miss;
  /* environment importance sampling tables (only when miss.env_nee != 0 and miss.type ==
   * PT_MISS_HDR with a spherical map; NEW, not in the reference): one entry per texel of the
   * environment image, row-major.  Alias method: pick texel i uniformly, keep it with
   * probability env_alias_prob[i], otherwise take env_alias_idx[i].  env_texel_pdf[i] is the
   * probability of texel i (sums to 1). */
  int32_t env_n;
  const float* env_alias_prob;
  const int32_t* env_alias_idx;
  const float* env_texel_pdf;
  /* ABI 2: time-dependent and volumetric primitives */
  int32_t n_moving_spheres;
  const pt_moving_sphere* moving_spheres;
  int32_t n_volume_spheres;
  const pt_volume_sphere* volume_spheres;
  int32_t n_volume_boxes;
  const pt_volume_box* volume_boxes;
} pt_scene_desc;

/* ---- scene construction --------------------------------------------------- */

typedef struct pt_scene pt_scene;
typedef struct pt_renderer pt_renderer;

const char* pt_last_error(void);
int pt_abi_version(void);

/* replaces Context::create() + createGroup()/createAcceleration("Trbvh")
 * (host_includes/host_common.hpp:24, host_includes/scenes.hpp:31-32) */
int pt_scene_create(pt_scene** out);
int pt_scene_destroy(pt_scene* scene);

/* replaces Sphere/AARect/Box::getGeometryInstance + addAndTransform
 * (host_includes/hitables.hpp:81-127,233-266,279-307; transforms.hpp:400-419).
 * xform_o2w: n*12 floats (row-major 3x4 object->world) or NULL for identity;
 * material: n material ids. */
int pt_scene_add_spheres(pt_scene* scene, const pt_sphere* prims, const float* xform_o2w,
                         const int32_t* material, int32_t n, int32_t* first_prim_id);
int pt_scene_add_rects(pt_scene* scene, const pt_rect* prims, const float* xform_o2w,
                       const int32_t* material, int32_t n, int32_t* first_prim_id);
int pt_scene_add_boxes(pt_scene* scene, const pt_box* prims, const float* xform_o2w,
                       const int32_t* material, int32_t n, int32_t* first_prim_id);

/* replaces Moving_Sphere / Volumetric_Sphere / Volumetric_Box ::getGeometryInstance
 * (host_includes/hitables.hpp:131-217,316-350).  Volumes need a PT_MAT_ISOTROPIC material to
 * scatter like the reference's scenes do, but any material is accepted (as in the reference). */
int pt_scene_add_moving_spheres(pt_scene* scene, const pt_moving_sphere* prims, const float* xform_o2w,
                                const int32_t* material, int32_t n, int32_t* first_prim_id);
int pt_scene_add_volume_spheres(pt_scene* scene, const pt_volume_sphere* prims, const float* xform_o2w,
                                const int32_t* material, int32_t n, int32_t* first_prim_id);
int pt_scene_add_volume_boxes(pt_scene* scene, const pt_volume_box* prims, const float* xform_o2w,
                              const int32_t* material, int32_t n, int32_t* first_prim_id);

/* replaces Mesh::getGeometryInstance buffers (host_includes/mesh.hpp:158-201):
 * unindexed per-face arrays.  v: ntris*9; n: ntris*9 or NULL; uv: ntris*6 or
 * NULL; tex_index: ntris or NULL.  xform_o2w (12 floats or NULL) is applied to
 * the vertices/normals on the host (the BVH holds world-space triangles). */
int pt_scene_add_mesh(pt_scene* scene, const float* v, const float* n, const float* uv,
                      const int32_t* tex_index, int32_t ntris, const float* xform_o2w,
                      int32_t material, int32_t* first_prim_id);

/* replaces BRDF::assignTo / Texture::assignTo program creation
 * (host_includes/materials.hpp:47-242, textures.hpp:28-270) */
int pt_scene_set_materials(pt_scene* scene, const pt_material* m, int32_t n);
int pt_scene_set_textures(pt_scene* scene, const pt_texture* t, int32_t n);
int pt_scene_set_texture_vector(pt_scene* scene, const int32_t* texture_ids, int32_t n);
/* replaces Noise_Texture's ranvec/perm_x/perm_y/perm_z buffers and Gradient_Texture's colours
 * (host_includes/textures.hpp:64-108,236-252): one float table, textures index into it */
int pt_scene_set_texture_data(pt_scene* scene, const float* data, int32_t n);
/* replaces Image_Texture::loadTexture / HDR_Texture::loadHDRTexture buffers
 * (host_includes/textures.hpp:115-165,178-217): RGBA float, row 0 first. */
int pt_scene_add_image(pt_scene* scene, const float* rgba, int32_t width, int32_t height,
                       int32_t* image_id);
/* replaces setRayGenerationProgram's Light_Sample/Light_PDF/Light_Emissions
 * buffers (host_includes/programs.hpp:17-30) */
int pt_scene_set_lights(pt_scene* scene, const pt_light* lights, int32_t n);
/* replaces Camera::set (host_includes/camera.hpp:40-51) */
int pt_scene_set_camera(pt_scene* scene, const pt_camera* cam);
/* replaces setMissProgram (host_includes/programs.hpp:35-92) */
int pt_scene_set_miss(pt_scene* scene, const pt_miss* miss);

/* replaces the first `renderFrame(ctx,0,0)` ("OptiX Building Time",
 * main.cpp:99): uploads the scene and builds the BVH on the GPU. */
int pt_scene_commit(pt_scene* scene);

/* host-side flat description (valid until the scene is modified/destroyed) */
int pt_scene_get_desc(const pt_scene* scene, pt_scene_desc* out);

typedef struct pt_bvh_info {
  int32_t n_prims;
  int32_t n_bvh2_nodes;
  int32_t n_bvh8_nodes;
  int32_t n_leaf_records;
  float sah_cost_lbvh;    /* SAH cost of the BVH2 before refinement          */
  float sah_cost_refined; /* ... after binned-SAH treelet refinement         */
  float build_ms;         /* GPU time of the whole build                     */
  int64_t node_bytes;
  int64_t leaf_bytes;
} pt_bvh_info;
int pt_scene_bvh_info(const pt_scene* scene, pt_bvh_info* out);

/* ---- traversal entry point (hit-ID parity) -------------------------------- */

/* replaces rtTrace(world, ray, prd) for a host ray set (programs/raygen.cu:77):
 * closest hit in (tmin, tmax), ties -> lowest primitive id.
 * `counters` (nullable): [0]=BVH8 nodes fetched, [1]=leaf records fetched. */
int pt_trace_rays(pt_scene* scene, const pt_ray* rays, int64_t n, pt_hit* hits,
                  uint64_t* counters);
/* same with device-resident buffers (d_rays/d_hits are CUDA device pointers) */
int pt_trace_rays_device(pt_scene* scene, const void* d_rays, int64_t n, void* d_hits,
                         void* cuda_stream);

/* ---- renderer -------------------------------------------------------------- */

/* replaces createFrameBuffer + context["acc_buffer"] (main.cpp:92-93,
 * host_includes/buffers.hpp:13-18).  d_accum: optional caller-owned device
 * buffer of width*height float4 (e.g. a torch tensor used as the NCCL reduce
 * buffer); NULL = library allocates. */
int pt_renderer_create(pt_scene* scene, int32_t width, int32_t height, void* d_accum,
                       pt_renderer** out);
int pt_renderer_destroy(pt_renderer* r);

/* replaces the host loop `context["frame"]->setInt(s); launch(0,W,H)` for
 * s in [first_frame, first_frame+nframes) (main.cpp:264-293).  frame 0 of a
 * renderer's life clears the accumulation buffer like raygen.cu:154 does
 * (clear_accum != 0 forces it).  Asynchronous w.r.t. the host only through
 * pt_renderer_sync.  Every (pixel, frame) pair of the range is rendered exactly once with
 * the seed tea<64>(W*y + x, frame) (raygen.cu:150); the ORDER inside one call is the
 * library's (frames in groups, pixels in tiles, images above PT_BAND_PIXELS in bands of
 * rows), so the accumulation buffer equals the reference's frame-by-frame loop up to the
 * order of the fp32 additions into a pixel. */
int pt_render(pt_renderer* r, int32_t first_frame, int32_t nframes, int32_t max_depth,
              int32_t russian, int32_t clear_accum);
int pt_renderer_sync(pt_renderer* r);

/* replaces Buffer::map() on accBuffer (host_includes/image_save.hpp:11):
 * width*height float4, row 0 = top image row, .w = number of samples. */
int pt_read_accum(pt_renderer* r, float* host_rgba);
int pt_accum_device_ptr(pt_renderer* r, void** d_ptr);
/* replaces display_buffer + make_Color (programs/raygen.cu:136-146,167) and its map() in the
 * progressive loop (main.cpp:268-272): width*height RGBA8 preview of the accumulation buffer,
 * (int)(255.99 * clamp(sqrt(rgb / samples), 0, 1)), alpha 255, converted on the device (4 B per
 * pixel cross PCIe instead of 16).  The divisor is the pixel's own sample count (.w), which is
 * `frame + 1` in the reference's one-launch-per-sample loop. */
int pt_read_display(pt_renderer* r, uint8_t* host_rgba8);
/* checkpoint / resume of a progressive render: overwrites the accumulation buffer with one saved
 * by pt_read_accum (same layout).  Continue with pt_render(first_frame = samples already in the
 * buffer, ..., clear_accum = 0).  The reference's loop cannot resume (main.cpp:264-293 restarts at
 * sample 0 and raygen.cu:154 clears on frame 0); SURVEY.md 8f rank 4. */
int pt_write_accum(pt_renderer* r, const float* host_rgba);

typedef struct pt_stats {
  uint64_t samples;        /* camera paths finished                          */
  uint64_t closest_rays;   /* closest-hit traversal queries                  */
  uint64_t shadow_rays;    /* shadow traversal queries                       */
  uint64_t nodes_fetched;  /* BVH8 nodes fetched (instrumented mode only)    */
  uint64_t leaves_fetched; /* leaf records fetched (instrumented mode only)  */
  uint64_t iterations;     /* wavefront iterations                           */
  uint64_t kernel_launches;
  float ms_total;          /* CUDA-event time of the last pt_render          */
  float ms_trace_closest;  /* per-kernel-class time (timing mode only)       */
  float ms_trace_shadow;
  float ms_shade;
  float ms_generate;
  float ms_reduce;         /* ABI 3: device time of the last NCCL merge of the accumulation buffers */
  uint32_t drained_paths;  /* ABI 3: 1 when the render ended in the single-launch drain kernel */
} pt_stats;
int pt_get_stats(pt_renderer* r, pt_stats* out);
/* bit0: count node/leaf fetches (instrumented kernels); bit1: per-kernel-class
 * CUDA-event timing (adds synchronisation; off for benchmarking) */
int pt_renderer_set_flags(pt_renderer* r, int32_t flags);
/* path-pool capacity (paths in flight); 0 = default */
int pt_renderer_set_pool(pt_renderer* r, int64_t paths);
/* frees the process-level caches (parked path-pool slabs, BVH build arenas) of every device */
int pt_release_caches(void);

/* ---- multi-GPU (ABI 3) --------------------------------------------------------
 * The reference renders on one GPU (one OptiX context, main.cpp:30-33, one launch per sample,
 * main.cpp:264-293).  Samples are independent - seed = tea<64>(W*y+x, frame), programs/raygen.cu:150 -
 * and the accumulation is a sum (raygen.cu:166), so the frames of a render can be split over devices with the
 * scene + BVH replicated and the float4 buffers merged by ONE sum-reduce (NCCL over NVLink).  The merged
 * buffer equals the single-device buffer up to fp32 summation order; .w sums to the sample count.
 *
 * (1) One process, several devices - the reference's usage model (a C++ host program):
 *     pt_renderer_create_multi replicates the committed scene (host data copied, uploaded and its BVH built
 *     on every other device), creates one renderer per device and returns ONE handle.  pt_render on it
 *     renders frames [first, first+n) split as pt_split_frames does, all devices concurrently;
 *     pt_read_accum / pt_read_display / pt_accum_device_ptr first merge the buffers into the first device's
 *     with ncclReduce.  The list must contain the device the scene was committed on (the root of the merge).
 * (2) One process per device (torchrun / MPI): every rank builds and commits the scene on its own device,
 *     renders its share (pt_split_frames) with an ordinary renderer, and the ranks merge with pt_comm_reduce.
 *     Rank 0 creates the 128-byte id with pt_comm_unique_id and hands it to the others by any means
 *     (bench.py broadcasts it through torch.distributed; a file or MPI_Bcast do as well).
 * NCCL is loaded at run time (libnccl.so.2); without it these calls fail with PT_ERR_STATE. */
int pt_renderer_create_multi(pt_scene* scene, int32_t width, int32_t height, const int32_t* device_ids,
                             int32_t n_devices, pt_renderer** out);
/* the share of `total_frames` that participant `rank` of `world` renders: contiguous, sizes differ by at most 1 */
int pt_split_frames(int32_t total_frames, int32_t rank, int32_t world, int32_t first_frame,
                    int32_t* out_first, int32_t* out_n);

/* the same with shares proportional to `weights` (e.g. the ranks' measured rates: GPUs of one box differ by 2-3 %
 * and a sample split ends when its slowest rank does); largest-remainder rounding, deterministic, every frame assigned
 * exactly once; weights NULL or all equal = pt_split_frames */
int pt_split_frames_weighted(int32_t total_frames, int32_t rank, int32_t world, const float* weights,
                             int32_t first_frame, int32_t* out_first, int32_t* out_n);

typedef struct pt_comm pt_comm;
#define PT_COMM_ID_BYTES 128
int pt_comm_unique_id(uint8_t id[PT_COMM_ID_BYTES]);
/* joins the communicator on the calling thread's current device; collective over all `world` ranks */
int pt_comm_create(const uint8_t id[PT_COMM_ID_BYTES], int32_t rank, int32_t world, pt_comm** out);
int pt_comm_destroy(pt_comm* comm);
/* sum-reduce of the renderers' accumulation buffers to rank `root` (in place), on the renderer's stream;
 * synchronous; the device time lands in pt_stats.ms_reduce.  The only collective of the path. */
int pt_comm_reduce(pt_comm* comm, pt_renderer* r, int32_t root);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
