// ref_render.cpp — THE REFERENCE RUN ON THE CPU: an emulation of the part of the OptiX 6 pipeline the reference
// uses (acceleration structure + traversal, program dispatch, callable programs) around the reference's OWN
// programs, which are compiled unmodified from /root/reference (ref_cu_probe.cpp, ref_mat_probe.cpp,
// ref_ray_probe.cpp, ref_tex_probe.cpp).  Everything that computes a ray, a hit, a colour or a random number is
// reference code; this file only decides which program runs next, the way OptiX would:
//   ray-generation program (raygen.cu renderPixel)  --rtTrace-->  traversal over the primitives' intersection
//   programs --> closest-hit program of the hit primitive's material (or the miss program) --rtTrace (shadow)-->
//   traversal --> any-hit rule.
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_render.so): an end-to-end check of the oracle against the reference
// (tests/test_reference_render.py) and the `--impl reference` arm of bench.py.
//
// Emulation choices (the same pins as the oracle's, see oracle.cpp): open ray interval, closest hit with ties to the
// lowest primitive id, a shadow ray asks the any-hit rule of its CLOSEST hit, Transform nodes = ray to object space /
// record back to world space, software bilinear image lookup, volumes seeded per (ray, primitive).  Acceleration:
// a median-split BVH over the reference's own bounding-box programs (OptiX's "Trbvh" is closed source).
// One rendering thread per process (the reference's program variables are process globals, like in one OptiX
// launch index); bench.py runs one process per host core on disjoint rows.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../include/ptb200.h"

extern "C" {
// reference programs (one translation unit each, see the *_probe.cpp files)
#define DECL_PRIM(kind) \
  int refcu_##kind(const float* prm, const float* ray8, float time, unsigned seed, float* hit3, float* rec14, float* bounds6);
DECL_PRIM(sphere) DECL_PRIM(rect) DECL_PRIM(box) DECL_PRIM(triangle) DECL_PRIM(moving_sphere) DECL_PRIM(volume_sphere)
DECL_PRIM(volume_box)
int refcu_any_hit(int light, const float* N, const float* d);
#define DECL_MAT(name)                                                                                              \
  void refmat_##name(const float* rec14, const float* ray8, float t, const float* tex6, const float* mp3, const float* lights, \
                     int in_shadow, float* prd16, float* shadow8);
DECL_MAT(lambertian) DECL_MAT(metal) DECL_MAT(dielectric) DECL_MAT(diffuse_light) DECL_MAT(oren_nayar) DECL_MAT(torrance_sparrow)
DECL_MAT(ashikhmin_shirley) DECL_MAT(normal_shader) DECL_MAT(isotropic)
void refmiss(int type, int spherical, const float* color1, const float* color2, const float* texel, const float* dir,
             const float* thr_in, float* thr_out, float* uv_out, int* event_out);
typedef void (*radiance_hook_t)(const float* ray8, float* prd16);
void refray_render_rows(const float* cam, int W, int H, int y0, int y1, int first, int n, int max_depth, int russian_,
                        radiance_hook_t hook, float* acc, unsigned char* disp);
#define DECL_TEX(name) void reftex_##name(const float* prm, int n, float u, float v, const float* p, int index, float* rgb);
DECL_TEX(constant) DECL_TEX(checker) DECL_TEX(gradient) DECL_TEX(vector) DECL_TEX(noise)
typedef int (*refshim_shadow_hook_t)(const float* ray8, const float* normal3);
typedef void (*refshim_texture_hook_t)(int tex, float u, float v, const float* p3, int index, float* rgb3);
extern refshim_shadow_hook_t refshim_shadow_hook;
extern refshim_texture_hook_t refshim_texture_hook;
extern int refshim_hit_index;
}

namespace {

struct Prim {
  int kind, material, xform;
  int tex_index;  // HitRecord::index: the mesh's per-triangle texture index, 0 for everything else
  std::vector<float> prm;
  float lo[3], hi[3];  // world bounds from the reference's bounds program
};
struct Node {
  float lo[3], hi[3];
  int left, first, count;
};
struct Scene {
  std::vector<Prim> prims;
  std::vector<pt_xform> xforms;
  std::vector<pt_material> materials;
  std::vector<pt_texture> textures;
  std::vector<int> texvec;
  std::vector<float> texdata;
  struct Img { int w, h; std::vector<float> rgba; };
  std::vector<Img> images;
  std::vector<float> lights;  // count, then 10 floats per light (pt_light)
  pt_camera cam;
  pt_miss miss;
  std::vector<Node> nodes;
  std::vector<int> order;
};
Scene* g = nullptr;       // the scene of the render in progress (program variables are process globals anyway)
float g_time = 0.f;       // prd.time of the path in flight (read by moving spheres on shadow rays too)
uint32_t g_shadow_seed = 0;  // prd.seed at the moment Direct_Light traces its shadow ray (volumes are seeded with it)

void xf_point(const float* m, const float* p, float* o) {
  for (int r = 0; r < 3; r++) o[r] = m[4 * r] * p[0] + m[4 * r + 1] * p[1] + m[4 * r + 2] * p[2] + m[4 * r + 3];
}
void xf_vector(const float* m, const float* v, float* o) {
  for (int r = 0; r < 3; r++) o[r] = m[4 * r] * v[0] + m[4 * r + 1] * v[1] + m[4 * r + 2] * v[2];
}
void xf_normal(const float* w2o, const float* n, float* o) {  // transpose(world->object) * n
  for (int c = 0; c < 3; c++) o[c] = w2o[c] * n[0] + w2o[4 + c] * n[1] + w2o[8 + c] * n[2];
}
void normalize3(float* v) {  // optixu normalize: v * (1 / sqrt(dot))
  float inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  v[0] *= inv; v[1] *= inv; v[2] *= inv;
}
uint32_t tea4(uint32_t v0, uint32_t v1) {  // the pinned per-(ray, primitive) volume seed (oracle.cpp header)
  uint32_t s0 = 0;
  for (int n = 0; n < 4; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}

typedef int (*prim_fn)(const float*, const float*, float, unsigned, float*, float*, float*);
prim_fn prim_program(int kind) {
  switch (kind) {
    case PT_PRIM_TRIANGLE: return refcu_triangle;
    case PT_PRIM_SPHERE: return refcu_sphere;
    case PT_PRIM_RECT: return refcu_rect;
    case PT_PRIM_BOX: return refcu_box;
    case PT_PRIM_MOVING_SPHERE: return refcu_moving_sphere;
    case PT_PRIM_VOLUME_SPHERE: return refcu_volume_sphere;
    default: return refcu_volume_box;
  }
}

// one primitive's intersection program on a world ray; fills the world-space hit record.  The program always sees
// the ray's OWN interval (tmin, tmax) - the volume programs read ray.tmax - and the closest hit is selected outside
bool intersect(const Scene& s, int id, const float* ray8, unsigned seed, float* hit3, float* rec14) {
  const Prim& p = s.prims[id];
  float r8[8];
  memcpy(r8, ray8, 32);
  if (p.xform >= 0) {  // Transform node: the program sees the ray in object space (t is shared)
    xf_point(s.xforms[p.xform].w2o, ray8, r8);
    xf_vector(s.xforms[p.xform].w2o, ray8 + 4, r8 + 4);
  }
  float bnd[6];
  if (!prim_program(p.kind)(p.prm.data(), r8, g_time, tea4(seed, (uint32_t)id), hit3, rec14, bnd)) return false;
  if (p.xform >= 0) {  // record to world space (intended rtTransform* semantics; the oracle's pin)
    float q[3];
    xf_point(s.xforms[p.xform].o2w, rec14, q);
    memcpy(rec14, q, 12);
    for (int k = 3; k <= 6; k += 3) {
      xf_normal(s.xforms[p.xform].w2o, rec14 + k, q);
      normalize3(q);
      memcpy(rec14 + k, q, 12);
    }
    float wo[3] = {-ray8[4], -ray8[5], -ray8[6]};
    normalize3(wo);
    memcpy(rec14 + 11, wo, 12);
  }
  return true;
}

bool slab(const Node& n, const float* o, const float* inv, float tmin, float tbest) {
  float t0 = tmin, t1 = tbest;
  for (int a = 0; a < 3; a++) {
    float ta = (n.lo[a] - o[a]) * inv[a], tb = (n.hi[a] - o[a]) * inv[a];
    if (ta > tb) std::swap(ta, tb);
    if (ta > t0) t0 = ta;  // NaN-tolerant: a degenerate slab (0 * inf) leaves the interval alone
    if (tb < t1) t1 = tb;
  }
  return t0 <= t1 * 1.0000004f + 1e-30f;
}

// closest hit in (tmin, tmax), ties to the lowest primitive id
int closest(const Scene& s, const float* ray8, unsigned seed, float* hit3, float* rec14) {
  int best = -1;
  float tbest = ray8[7];
  float h3[3], rec[14];
  auto consider = [&](int id) {
    // leaves are visited in tree order, hence the explicit tie rule (equal t: the lowest primitive id)
    if (intersect(s, id, ray8, seed, h3, rec)) {
      if (h3[0] < tbest || (h3[0] == tbest && (best < 0 || id < best))) {
        tbest = h3[0];
        best = id;
        memcpy(hit3, h3, 12);
        memcpy(rec14, rec, 56);
      }
    }
  };
  if (s.nodes.empty()) {
    for (int i = 0; i < (int)s.prims.size(); i++) consider(i);
    return best;
  }
  float inv[3];
  for (int a = 0; a < 3; a++) {
    float d = ray8[4 + a];
    if (fabsf(d) < 1e-30f) d = signbit(d) ? -1e-30f : 1e-30f;
    inv[a] = 1.0f / d;
  }
  int stack[128], sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const Node& n = s.nodes[stack[--sp]];
    if (!slab(n, ray8, inv, ray8[3], tbest)) continue;
    if (n.count > 0) {
      for (int i = 0; i < n.count; i++) consider(s.order[n.first + i]);
    } else {
      stack[sp++] = n.left;
      stack[sp++] = n.left + 1;
    }
  }
  return best;
}

void build_bvh(Scene& s) {
  const int n = (int)s.prims.size();
  s.nodes.clear();
  s.order.resize(n);
  for (int i = 0; i < n; i++) s.order[i] = i;
  if (n <= 8) return;  // brute force
  s.nodes.push_back({});
  struct Job { int node, first, count; };
  std::vector<Job> jobs{{0, 0, n}};
  while (!jobs.empty()) {
    Job j = jobs.back();
    jobs.pop_back();
    Node nd;
    for (int a = 0; a < 3; a++) { nd.lo[a] = INFINITY; nd.hi[a] = -INFINITY; }
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = j.first; i < j.first + j.count; i++) {
      const Prim& p = s.prims[s.order[i]];
      for (int a = 0; a < 3; a++) {
        // pad like the oracle does: the reference's own pads (1e-4 .. 1e-3, none for spheres) are too tight for a
        // float slab test at large coordinates
        const float pad = 1e-3f + 1e-5f * std::max(fabsf(p.lo[a]), fabsf(p.hi[a]));
        nd.lo[a] = std::min(nd.lo[a], p.lo[a] - pad);
        nd.hi[a] = std::max(nd.hi[a], p.hi[a] + pad);
        const float c = 0.5f * (p.lo[a] + p.hi[a]);
        clo[a] = std::min(clo[a], c);
        chi[a] = std::max(chi[a], c);
      }
    }
    nd.first = j.first;
    nd.count = 0;
    nd.left = -1;
    if (j.count <= 4) {
      nd.count = j.count;
      s.nodes[j.node] = nd;
      continue;
    }
    int ax = 0;
    for (int a = 1; a < 3; a++)
      if (chi[a] - clo[a] > chi[ax] - clo[ax]) ax = a;
    const int mid = j.first + j.count / 2;
    std::nth_element(s.order.begin() + j.first, s.order.begin() + mid, s.order.begin() + j.first + j.count, [&](int a, int b) {
      return s.prims[a].lo[ax] + s.prims[a].hi[ax] < s.prims[b].lo[ax] + s.prims[b].hi[ax];
    });
    nd.left = (int)s.nodes.size();
    s.nodes[j.node] = nd;
    s.nodes.push_back({});
    s.nodes.push_back({});
    jobs.push_back({nd.left, j.first, mid - j.first});
    jobs.push_back({nd.left + 1, mid, j.first + j.count - mid});
  }
  // leaves list their primitives in ascending id (ties are resolved explicitly anyway)
  for (Node& nd : s.nodes)
    if (nd.count > 0) std::sort(s.order.begin() + nd.first, s.order.begin() + nd.first + nd.count);
}

// ---- textures: the reference's texture programs; sub-textures of checker / vector are evaluated first --------
void image_lookup(const Scene::Img& im, float u, float v, float* rgb) {  // software bilinear, repeat (the oracle's pin)
  rgb[0] = rgb[1] = rgb[2] = 0.f;
  if (!(std::isfinite(u) && std::isfinite(v)) || im.w <= 0 || im.h <= 0) return;
  float x = u * (float)im.w - 0.5f, y = v * (float)im.h - 0.5f;
  float fx0 = floorf(x), fy0 = floorf(y), fx = x - fx0, fy = y - fy0;
  auto wrap = [](float f, int n) {
    float m = fmodf(f, (float)n);
    int i = (int)m;
    if (i < 0) i += n;
    if (i >= n) i -= n;
    return i;
  };
  int x0 = wrap(fx0, im.w), y0 = wrap(fy0, im.h);
  int x1 = x0 + 1 == im.w ? 0 : x0 + 1, y1 = y0 + 1 == im.h ? 0 : y0 + 1;
  for (int c = 0; c < 3; c++) {
    auto T = [&](int xx, int yy) { return im.rgba[4 * ((size_t)yy * im.w + xx) + c]; };
    float top = T(x0, y0) + fx * (T(x1, y0) - T(x0, y0));
    float bot = T(x0, y1) + fx * (T(x1, y1) - T(x0, y1));
    rgb[c] = top + fy * (bot - top);
  }
}
void texture_eval(int tex, float u, float v, const float* p, int index, float* rgb) {
  const Scene& s = *g;
  rgb[0] = rgb[1] = rgb[2] = 0.f;
  if (tex < 0 || tex >= (int)s.textures.size()) return;
  const pt_texture& t = s.textures[tex];
  switch (t.type) {
    case PT_TEX_CONSTANT:
      reftex_constant(t.color, 0, u, v, p, index, rgb);
      break;
    case PT_TEX_CHECKER: {
      float q[6];
      texture_eval(t.a, u, v, p, 0, q);
      texture_eval(t.b, u, v, p, 0, q + 3);
      reftex_checker(q, 0, u, v, p, index, rgb);
      break;
    }
    case PT_TEX_IMAGE:
      if (t.a >= 0 && t.a < (int)s.images.size()) image_lookup(s.images[t.a], u, v, rgb);
      break;
    case PT_TEX_VECTOR: {
      std::vector<float> q((size_t)3 * std::max(t.b, 1));
      for (int i = 0; i < t.b; i++) texture_eval(s.texvec[t.a + i], u, v, p, 0, q.data() + 3 * i);
      reftex_vector(q.data(), t.b, u, v, p, index, rgb);
      break;
    }
    case PT_TEX_NOISE: {
      std::vector<float> q(2 + 1536);
      q[0] = t.color[0];
      q[1] = (float)t.b;
      memcpy(q.data() + 2, s.texdata.data() + t.a, 1536 * 4);
      reftex_noise(q.data(), 0, u, v, p, index, rgb);
      break;
    }
    case PT_TEX_GRADIENT:
      reftex_gradient(s.texdata.data() + t.a, 0, u, v, p, index, rgb);
      break;
  }
}
void texture_hook(int tex, float u, float v, const float* p, int index, float* rgb) { texture_eval(tex, u, v, p, index, rgb); }

// ---- shadow rays: the any-hit rule (hit.cu) of the closest hit; a miss stays "in shadow" ------------------------
int shadow_hook(const float* ray8, const float* normal) {
  float h3[3], rec[14];
  const int id = closest(*g, ray8, g_shadow_seed, h3, rec);
  if (id < 0) return 1;
  const int light = g->materials[g->prims[id].material].type == PT_MAT_DIFFUSE_LIGHT ? 1 : 0;
  return refcu_any_hit(light, normal, ray8 + 4);
}

// ---- one rtTrace of a radiance ray -------------------------------------------------------------------------------
typedef void (*mat_fn)(const float*, const float*, float, const float*, const float*, const float*, int, float*, float*);
mat_fn material_program(int type) {
  switch (type) {
    case PT_MAT_LAMBERTIAN: return refmat_lambertian;
    case PT_MAT_METAL: return refmat_metal;
    case PT_MAT_DIELECTRIC: return refmat_dielectric;
    case PT_MAT_DIFFUSE_LIGHT: return refmat_diffuse_light;
    case PT_MAT_OREN_NAYAR: return refmat_oren_nayar;
    case PT_MAT_TORRANCE_SPARROW: return refmat_torrance_sparrow;
    case PT_MAT_ASHIKHMIN_SHIRLEY: return refmat_ashikhmin_shirley;
    case PT_MAT_ISOTROPIC: return refmat_isotropic;
    default: return refmat_normal_shader;
  }
}
void radiance_hook(const float* ray8, float* prd16) {
  const Scene& s = *g;
  uint32_t seed;
  memcpy(&seed, prd16, 4);
  g_time = prd16[1];
  float h3[3], rec[14];
  const int id = closest(s, ray8, seed, h3, rec);
  if (id >= 0) {
    const pt_material& m = s.materials[s.prims[id].material];
    const float tex6[6] = {(float)m.tex0, 0.f, 0.f, (float)m.tex1, 0.f, 0.f};  // texture ids: colours come from the hook
    const float mp3[3] = {m.p0, m.p1, (float)(m.flags & 1)};
    float sh[8];
    // Direct_Light draws the light index and two sample coordinates (3 LCG steps, light_sample.cuh:48-56) before its
    // shadow ray: that state seeds the volumes a shadow ray meets (the payload of a shadow ray carries no seed)
    g_shadow_seed = seed;
    for (int k = 0; k < 3; k++) g_shadow_seed = 1664525u * g_shadow_seed + 1013904223u;
    refshim_hit_index = s.prims[id].tex_index;
    material_program(m.type)(rec, ray8, h3[0], tex6, mp3, s.lights.data(), 1, prd16, sh);
    refshim_hit_index = 0;
    return;
  }
  // miss program (miss.cu): image backgrounds sample at coordinates the program computes
  const pt_miss& ms = s.miss;
  float thr[3], uv[2], texel[3] = {0.f, 0.f, 0.f};
  int ev = 0;
  if (ms.type == PT_MISS_IMG || ms.type == PT_MISS_HDR) {
    refmiss(ms.type, ms.spherical, ms.color1, ms.color2, texel, ray8 + 4, prd16 + 2, thr, uv, &ev);
    if (ms.image >= 0 && ms.image < (int)s.images.size()) image_lookup(s.images[ms.image], uv[0], uv[1], texel);
  }
  refmiss(ms.type, ms.spherical, ms.color1, ms.color2, texel, ray8 + 4, prd16 + 2, thr, uv, &ev);
  memcpy(prd16 + 2, thr, 12);
  prd16[14] = (float)ev;
}

}  // namespace

extern "C" {

void* refrender_create(const pt_scene_desc* d) {
  Scene* s = new Scene();
  s->xforms.assign(d->xforms, d->xforms + d->n_xforms);
  s->materials.assign(d->materials, d->materials + d->n_materials);
  s->textures.assign(d->textures, d->textures + d->n_textures);
  s->texvec.assign(d->texvec, d->texvec + d->n_texvec);
  s->texdata.assign(d->texdata, d->texdata + d->n_texdata);
  for (int i = 0; i < d->n_images; i++)
    s->images.push_back({d->images[i].width, d->images[i].height,
                         std::vector<float>(d->images[i].rgba, d->images[i].rgba + (size_t)4 * d->images[i].width * d->images[i].height)});
  s->lights.push_back((float)d->n_lights);
  for (int i = 0; i < d->n_lights; i++) {
    const pt_light& L = d->lights[i];
    const float q[10] = {(float)L.type, (float)L.axis, L.a0, L.a1, L.b0, L.b1, L.k, L.emission[0], L.emission[1], L.emission[2]};
    s->lights.insert(s->lights.end(), q, q + 10);
  }
  s->lights.resize(s->lights.size() + 10, 0.f);
  s->cam = d->camera;
  s->miss = d->miss;
  for (int i = 0; i < d->n_prims; i++) {
    const pt_prim_ref& r = d->prims[i];
    Prim p;
    p.kind = r.type;
    p.material = r.material;
    p.xform = r.xform;
    p.tex_index = (r.type == PT_PRIM_TRIANGLE && d->tri_tex) ? d->tri_tex[r.index] : 0;
    switch (r.type) {
      case PT_PRIM_TRIANGLE: {
        p.prm.assign(26, 0.f);
        memcpy(p.prm.data(), d->tri_v + 9 * (size_t)r.index, 36);
        if (d->tri_n) { p.prm[9] = 1.f; memcpy(p.prm.data() + 10, d->tri_n + 9 * (size_t)r.index, 36); }
        if (d->tri_uv) { p.prm[19] = 1.f; memcpy(p.prm.data() + 20, d->tri_uv + 6 * (size_t)r.index, 24); }
        break;
      }
      case PT_PRIM_SPHERE: { const pt_sphere& q = d->spheres[r.index]; p.prm = {q.center[0], q.center[1], q.center[2], q.radius}; break; }
      case PT_PRIM_RECT: { const pt_rect& q = d->rects[r.index]; p.prm = {q.a0, q.a1, q.b0, q.b1, q.k, (float)q.axis, (float)q.flip}; break; }
      case PT_PRIM_BOX: { const pt_box& q = d->boxes[r.index]; p.prm.assign(q.bmin, q.bmin + 6); break; }
      case PT_PRIM_MOVING_SPHERE: { const pt_moving_sphere& q = d->moving_spheres[r.index]; p.prm.assign(q.center0, q.center0 + 9); break; }
      case PT_PRIM_VOLUME_SPHERE: { const pt_volume_sphere& q = d->volume_spheres[r.index]; p.prm.assign(q.center, q.center + 5); break; }
      default: { const pt_volume_box& q = d->volume_boxes[r.index]; p.prm.assign(q.bmin, q.bmin + 7); break; }
    }
    // world bounds: the reference's bounding-box program (object space), its 8 corners through the Transform
    float ray8[8] = {0, 0, 0, 1e-3f, 0, 0, 1, 1e27f}, h3[3], rec[14], b[6];
    prim_program(p.kind)(p.prm.data(), ray8, d->camera.time0, 0u, h3, rec, b);
    if (p.kind == PT_PRIM_MOVING_SPHERE) {  // + the camera's shutter interval (oracle.cpp header)
      for (float t : {d->camera.time0, d->camera.time1}) {
        const float* q = p.prm.data();
        const float sft = (t - q[7]) / (q[8] - q[7]);
        for (int a = 0; a < 3; a++) {
          const float c = q[a] + sft * (q[3 + a] - q[a]);
          if (c == c) { b[a] = std::min(b[a], c - q[6]); b[3 + a] = std::max(b[3 + a], c + q[6]); }
        }
      }
    }
    for (int a = 0; a < 3; a++) { p.lo[a] = INFINITY; p.hi[a] = -INFINITY; }
    for (int c = 0; c < 8; c++) {
      float q[3] = {(c & 1) ? b[3] : b[0], (c & 2) ? b[4] : b[1], (c & 4) ? b[5] : b[2]}, w[3];
      if (p.xform >= 0) xf_point(s->xforms[p.xform].o2w, q, w);
      else memcpy(w, q, 12);
      for (int a = 0; a < 3; a++) { p.lo[a] = std::min(p.lo[a], w[a]); p.hi[a] = std::max(p.hi[a], w[a]); }
    }
    s->prims.push_back(std::move(p));
  }
  build_bvh(*s);
  return s;
}
void refrender_destroy(void* s) { delete (Scene*)s; }

// frames [first, first + n) of rows [y0, y1) (pixelID.y, 0 = bottom row) ADDED into acc (W*H float4, row 0 = top);
// disp: the display buffer of those rows after the last frame
void refrender_render(void* scene, int W, int H, int y0, int y1, int first, int n, int max_depth, int russian,
                      float* acc, unsigned char* disp) {
  g = (Scene*)scene;
  refshim_shadow_hook = shadow_hook;
  refshim_texture_hook = texture_hook;
  const pt_camera& c = g->cam;
  const float cam[21] = {c.origin[0], c.origin[1], c.origin[2], c.lower_left_corner[0], c.lower_left_corner[1], c.lower_left_corner[2],
                         c.horizontal[0], c.horizontal[1], c.horizontal[2], c.vertical[0], c.vertical[1], c.vertical[2],
                         c.u[0], c.u[1], c.u[2], c.v[0], c.v[1], c.v[2], c.lens_radius, c.time0, c.time1};
  refray_render_rows(cam, W, H, y0, y1, first, n, max_depth, russian, radiance_hook, acc, disp);
  refshim_shadow_hook = nullptr;
  refshim_texture_hook = nullptr;
  g = nullptr;
}
}
