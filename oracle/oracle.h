/*
 * oracle.h — C API of the CPU oracle (liboracle.so).
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a from-scratch host C++/OpenMP
 * restatement of the reference's device programs
 * (/root/reference/OptiX-Path-Tracer/programs/, cited per function in
 * oracle.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product
 * (optix-path-tracer_b200/, libptb200.so) never links, imports or calls it.
 *
 * Parity pinning: the reference has no tests and cannot run here (needs the
 * proprietary OptiX 6 SDK + a GPU), but its device code compiles UNMODIFIED as
 * host C++ once the OptiX device-program environment is emulated
 * (oracle/ref_shim/).  oracle/Makefile builds, from the sources where they lie
 * under /root/reference, oracle/_ref/libref_{probe,cu,mat,ray,tex,host,render}.so: the BSDF /
 * RNG headers, the intersection + bounds + hit-record programs of all seven
 * primitive kinds, both light programs, all nine closest-hit programs with
 * Direct_Light, the ray-generation program (camera, color() loop, accumulation,
 * display pixel), the miss programs and the texture programs — and
 * libref_render.so (oracle/ref_render.cpp), all of those programs wired into an
 * emulated OptiX pipeline, i.e. the reference itself rendering whole images on the
 * CPU (tests/test_reference_render.py: bit-identical to this oracle; it is also
 * what bench.py --impl reference times, kind "reference").  This restatement
 * reproduces every one of them bit for bit (tests/test_oracle_*golden.py; vectors
 * recorded by tests/golden/make_golden*.py so the tests also run without the tree).
 * Pinned by decision, not by reference output (OptiX runtime behaviour): the open
 * ray interval, what a shadow ray answers (hit.cu leaves it undefined on a miss),
 * traversal order / tie-breaking, the per-(ray, primitive) uniform of the volumes,
 * hit records of transformed primitives, hardware texture filtering.
 *
 * It consumes the same flat scene description as the product
 * (include/ptb200.h: pt_scene_desc), so both sides see identical inputs.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include "../include/ptb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_stats {
  uint64_t samples;
  uint64_t closest_rays;
  uint64_t shadow_rays;
  uint64_t bounces; /* sum over samples of path length (closest-hit events) */
  uint64_t node_tests;
  uint64_t prim_tests;
  double seconds;
  int32_t threads;
} orc_stats;

/* acceleration handle: brute force for tiny scenes, BVH2 otherwise */
typedef struct orc_scene orc_scene;
/* mode: 0 = auto (brute force when n_prims <= 64), 1 = brute force, 2 = BVH2 */
orc_scene* orc_scene_create(const pt_scene_desc* desc, int mode);
void orc_scene_destroy(orc_scene* s);

/* closest hit in (tmin,tmax), ties -> lowest primitive id */
int orc_trace(const orc_scene* s, const pt_ray* rays, int64_t n, pt_hit* hits, orc_stats* stats);

/* camera rays of frame `frame` (programs/raygen.cu:148-162), one per pixel,
 * index = y*W + x with y = pixelID.y (0 = bottom image row). */
int orc_camera_rays(const orc_scene* s, int W, int H, int frame, pt_ray* out);
/* incoherent secondary rays: for each input ray that hits, a cosine-sampled
 * bounce about the shading normal (origin P, tmin 1e-3); misses get tmax<0 */
int orc_bounce_rays(const orc_scene* s, const pt_ray* in, int64_t n, uint32_t seed, pt_ray* out);

/* frames [first_frame, first_frame+nframes) ADDED into accum (W*H float4,
 * row 0 = top image row, .w += 1 per sample), programs/raygen.cu:148-167 */
int orc_render(const orc_scene* s, int W, int H, int first_frame, int nframes, int max_depth,
               int russian, float* accum, orc_stats* stats);

/* single pixel-sample with a trace of the path for debugging/parity */
int orc_render_pixel(const orc_scene* s, int W, int H, int x, int y, int frame, int max_depth,
                     int russian, float* rgb);

/* host_includes/image_save.hpp:13-29: 8-bit RGB, W*H*3 */
int orc_tonemap_png8(const float* accum, int W, int H, int samples, uint8_t* rgb);

/* programs/raygen.cu:136-146 make_Color: RGBA8 preview of an accumulation buffer */
int orc_display_rgba8(const float* accum, int W, int H, uint8_t* rgba);

/* ---- unit entry points (golden vectors / reference-header probes) -------- */
uint32_t orc_tea64(uint32_t v0, uint32_t v1);            /* programs/random.cuh:33-47 */
float orc_rnd(uint32_t* seed);                           /* programs/random.cuh:63-65 */
void orc_random_in_unit_disk(uint32_t* seed, float* out3);   /* sampling.cuh:22-29 */
void orc_random_in_unit_sphere(uint32_t* seed, float* out3); /* sampling.cuh:31-44 */
void orc_random_on_unit_sphere(uint32_t* seed, float* out3); /* sampling.cuh:46-59 */

/* BSDF probes.  mat: PT_MAT_*; color/color2: texture values; p0,p1 as pt_material.
 * u1,u2 are the two uniform draws in DEVICE order (u1 drawn first). */
void orc_bsdf_sample(int mat, const float* color, const float* color2, float p0, float p1,
                     const float* P, const float* Wo, const float* N, float u1, float u2,
                     float* Wi_out);
void orc_bsdf_eval(int mat, const float* color, const float* color2, float p0, float p1,
                   const float* P, const float* Wo, const float* Wi, const float* N, float* f_out,
                   float* pdf_out);
float orc_schlick(float cosine, float ref_idx);          /* material.cuh:30-34 */
int orc_refract(const float* v, const float* n, float ni_over_nt, float* out3); /* :50-63 */
void orc_ggx_sample(const float* origin, float u1, float u2, float nu, float nv, float* out3);
float orc_ggx_d(const float* H, float nu, float nv);
float orc_ggx_g(const float* Wo, const float* Wi, float nu, float nv);
/* rect light: programs/pdfs/rect_pdf.cu */
void orc_rect_sample(const pt_light* l, const float* P, float u1, float u2, float* Wi_out);
float orc_rect_pdf(const pt_light* l, const float* P, const float* Wi);
/* single primitive intersection: returns 1 and *t when hit in (tmin,tmax) */
int orc_intersect_prim(const orc_scene* s, int prim, const pt_ray* ray, float* t, float* u, float* v);
/* ... with prd.time / ray seed, returning hit3 = (t, u, v) and the hit record
 * rec14 = P, geometric normal, shading normal, u, v, Wo (compared with the reference's own .cu programs) */
int orc_intersect_prim_ex(const orc_scene* s, int prim, const pt_ray* ray, float time, uint32_t seed,
                          float* hit3, float* rec14);
/* one closest-hit program invocation (programs/materials/<x>.cu closest_hit); see oracle.cpp for the layout */
void orc_closest_hit_probe(const orc_scene* s, const pt_ray* ray, int prim, float t, float u, float v,
                           float* prd_io16, float* shadow8);
/* one rtTrace of a radiance ray (closest hit + its closest-hit program, or the miss program) on a payload */
void orc_trace_radiance(const orc_scene* s, const float* ray8, float* prd16);
void orc_miss_probe(const orc_scene* s, const float* dir, const float* thr_in, float* thr_out, float* uv2);
void orc_sample_texture(const orc_scene* s, int tex, float u, float v, const float* P, int index, float* rgb);
void orc_prim_bounds(const orc_scene* s, int prim, float* lo_hi6);
void orc_light_sample(const pt_light* l, const float* P, uint32_t* seed, float* Wi_out);
float orc_light_pdf(const pt_light* l, const float* P, const float* Wi);
uint32_t orc_volume_seed(uint32_t ray_seed, uint32_t prim); /* tea<4>: the pinned volume uniform's seed */

#ifdef __cplusplus
}
#endif
#endif
