// api.cu — C ABI (include/ptb200.h): scene construction, upload, commit.
// The renderer / traversal entry points live in integrator.cu.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "scene_host.h"

static thread_local char g_err[512] = "";

void pt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

// world->object from a row-major 3x4 object->world (double precision inverse)
bool invert_3x4(const float* m, float* out) {
  double a[3][3] = {{m[0], m[1], m[2]}, {m[4], m[5], m[6]}, {m[8], m[9], m[10]}};
  double t[3] = {m[3], m[7], m[11]};
  double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  double det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
  if (det == 0.0 || !std::isfinite(det)) return false;
  double id = 1.0 / det;
  double inv[3][3];
  inv[0][0] = c00 * id;
  inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  inv[1][0] = c01 * id;
  inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  inv[2][0] = c02 * id;
  inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) out[4 * r + c] = (float)inv[r][c];
    out[4 * r + 3] = (float)(-(inv[r][0] * t[0] + inv[r][1] * t[1] + inv[r][2] * t[2]));
  }
  return true;
}

bool is_identity(const float* m) {
  static const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  return memcmp(m, I, sizeof(I)) == 0;
}

int add_xform(pt_scene* s, const float* o2w, int* out_index) {
  if (!o2w || is_identity(o2w)) {
    *out_index = -1;
    return PT_OK;
  }
  pt_xform x;
  memcpy(x.o2w, o2w, sizeof(x.o2w));
  if (!invert_3x4(o2w, x.w2o)) {
    pt_set_error("singular object->world transform");
    return PT_ERR_INVALID;
  }
  *out_index = (int)s->xforms.size();
  s->xforms.push_back(x);
  return PT_OK;
}

template <typename T>
int add_analytic(pt_scene* s, std::vector<T>& store, int kind, const T* prims, const float* xf, const int32_t* material,
                 int32_t n, int32_t* first) {
  if (!s || n < 0 || (n > 0 && (!prims || !material))) {
    pt_set_error("add primitives: bad arguments");
    return PT_ERR_INVALID;
  }
  if (first) *first = (int32_t)s->prims.size();
  for (int i = 0; i < n; i++) {
    int xi = -1;
    int rc = add_xform(s, xf ? xf + 12 * (size_t)i : nullptr, &xi);
    if (rc) return rc;
    pt_prim_ref r;
    r.type = kind;
    r.index = (int32_t)store.size();
    r.material = material[i];
    r.xform = xi;
    store.push_back(prims[i]);
    s->prims.push_back(r);
  }
  s->committed = false;
  return PT_OK;
}

template <typename T>
int upload_vec(DevBuf& b, const std::vector<T>& v) {
  PT_CUDA(b.alloc(v.size() * sizeof(T)));
  if (!v.empty()) PT_CUDA(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return PT_OK;
}

}  // namespace

extern "C" {

const char* pt_last_error(void) { return g_err; }
int pt_abi_version(void) { return PTB200_ABI_VERSION; }

int pt_scene_create(pt_scene** out) {
  if (!out) return PT_ERR_INVALID;
  pt_scene* s = new pt_scene();
  s->miss.type = PT_MISS_CONSTANT;
  s->miss.image = -1;
  s->miss.spherical = 1;
  *out = s;
  return PT_OK;
}

int pt_scene_destroy(pt_scene* scene) {
  delete scene;
  return PT_OK;
}

int pt_scene_add_spheres(pt_scene* s, const pt_sphere* p, const float* xf, const int32_t* m, int32_t n, int32_t* first) {
  return add_analytic(s, s->spheres, PT_PRIM_SPHERE, p, xf, m, n, first);
}
int pt_scene_add_rects(pt_scene* s, const pt_rect* p, const float* xf, const int32_t* m, int32_t n, int32_t* first) {
  if (s && p)
    for (int i = 0; i < n; i++)
      if (p[i].axis < 0 || p[i].axis > 2) {
        pt_set_error("AARect: invalid axis %d", p[i].axis);  // aarect.cu:62 "Error: invalid axis"
        return PT_ERR_INVALID;
      }
  return add_analytic(s, s->rects, PT_PRIM_RECT, p, xf, m, n, first);
}
int pt_scene_add_boxes(pt_scene* s, const pt_box* p, const float* xf, const int32_t* m, int32_t n, int32_t* first) {
  return add_analytic(s, s->boxes, PT_PRIM_BOX, p, xf, m, n, first);
}

int pt_scene_add_moving_spheres(pt_scene* s, const pt_moving_sphere* p, const float* xf, const int32_t* m, int32_t n,
                                int32_t* first) {
  return add_analytic(s, s->moving_spheres, PT_PRIM_MOVING_SPHERE, p, xf, m, n, first);
}
int pt_scene_add_volume_spheres(pt_scene* s, const pt_volume_sphere* p, const float* xf, const int32_t* m, int32_t n,
                                int32_t* first) {
  if (s && p)
    for (int i = 0; i < n; i++)
      if (!(p[i].density > 0.f)) {
        pt_set_error("Volumetric_Sphere: density must be > 0");
        return PT_ERR_INVALID;
      }
  return add_analytic(s, s->volume_spheres, PT_PRIM_VOLUME_SPHERE, p, xf, m, n, first);
}
int pt_scene_add_volume_boxes(pt_scene* s, const pt_volume_box* p, const float* xf, const int32_t* m, int32_t n,
                              int32_t* first) {
  if (s && p)
    for (int i = 0; i < n; i++)
      if (!(p[i].density > 0.f)) {
        pt_set_error("Volumetric_Box: density must be > 0");
        return PT_ERR_INVALID;
      }
  return add_analytic(s, s->volume_boxes, PT_PRIM_VOLUME_BOX, p, xf, m, n, first);
}

int pt_scene_add_mesh(pt_scene* s, const float* v, const float* nrm, const float* uv, const int32_t* tex_index,
                      int32_t ntris, const float* xf, int32_t material, int32_t* first) {
  if (!s || ntris < 0 || (ntris > 0 && !v)) {
    pt_set_error("pt_scene_add_mesh: bad arguments");
    return PT_ERR_INVALID;
  }
  if (first) *first = (int32_t)s->prims.size();
  const size_t t0 = s->tri_v.size() / 9;
  pt_xform x;
  bool has_x = xf && !is_identity(xf);
  if (has_x) {
    memcpy(x.o2w, xf, sizeof(x.o2w));
    if (!invert_3x4(xf, x.w2o)) {
      pt_set_error("singular mesh transform");
      return PT_ERR_INVALID;
    }
  }
  // lazily materialise optional per-triangle arrays so all meshes share one layout
  auto ensure = [&](std::vector<float>& a, bool& flag, int per) {
    if (!flag) {
      a.assign(t0 * per, 0.f);
      flag = true;
    }
  };
  const bool fresh_n = nrm && !s->any_n;
  if (nrm) ensure(s->tri_n, s->any_n, 9);
  if (uv) ensure(s->tri_uv, s->any_uv, 6);
  if (tex_index && !s->any_tex) {
    s->tri_tex.assign(t0, 0);
    s->any_tex = true;
  }
  if (fresh_n && t0 > 0) {
    // earlier meshes had no normals: give them their geometric normal
    for (size_t t = 0; t < t0; t++) {
      float* n9 = s->tri_n.data() + 9 * t;
      if (n9[0] == 0.f && n9[1] == 0.f && n9[2] == 0.f && n9[3] == 0.f && n9[4] == 0.f && n9[5] == 0.f) {
        const float* q = s->tri_v.data() + 9 * t;
        float3 e1 = f3(q[3] - q[0], q[4] - q[1], q[5] - q[2]), e2 = f3(q[6] - q[0], q[7] - q[1], q[8] - q[2]);
        float3 g = normalize(cross(e1, e2));
        for (int k = 0; k < 3; k++) {
          n9[3 * k] = g.x;
          n9[3 * k + 1] = g.y;
          n9[3 * k + 2] = g.z;
        }
      }
    }
  }
  s->tri_v.reserve(s->tri_v.size() + (size_t)ntris * 9);
  s->prims.reserve(s->prims.size() + ntris);
  for (int t = 0; t < ntris; t++) {
    float w[9];
    for (int k = 0; k < 3; k++) {
      float3 p = f3(v[9 * (size_t)t + 3 * k], v[9 * (size_t)t + 3 * k + 1], v[9 * (size_t)t + 3 * k + 2]);
      if (has_x) p = xf_point(x.o2w, p);
      w[3 * k] = p.x;
      w[3 * k + 1] = p.y;
      w[3 * k + 2] = p.z;
    }
    s->tri_v.insert(s->tri_v.end(), w, w + 9);
    if (s->any_n) {
      float nn[9];
      if (nrm) {
        for (int k = 0; k < 3; k++) {
          float3 q = f3(nrm[9 * (size_t)t + 3 * k], nrm[9 * (size_t)t + 3 * k + 1], nrm[9 * (size_t)t + 3 * k + 2]);
          if (has_x) q = xf_normal(x.w2o, q);
          nn[3 * k] = q.x;
          nn[3 * k + 1] = q.y;
          nn[3 * k + 2] = q.z;
        }
      } else {
        float3 e1 = f3(w[3] - w[0], w[4] - w[1], w[5] - w[2]), e2 = f3(w[6] - w[0], w[7] - w[1], w[8] - w[2]);
        float3 g = normalize(cross(e1, e2));
        for (int k = 0; k < 3; k++) {
          nn[3 * k] = g.x;
          nn[3 * k + 1] = g.y;
          nn[3 * k + 2] = g.z;
        }
      }
      s->tri_n.insert(s->tri_n.end(), nn, nn + 9);
    }
    if (s->any_uv) {
      float z[6] = {0, 0, 0, 0, 0, 0};
      const float* src = uv ? uv + 6 * (size_t)t : z;
      s->tri_uv.insert(s->tri_uv.end(), src, src + 6);
    }
    if (s->any_tex) s->tri_tex.push_back(tex_index ? tex_index[t] : 0);
    pt_prim_ref r;
    r.type = PT_PRIM_TRIANGLE;
    r.index = (int32_t)(t0 + t);
    r.material = material;
    r.xform = -1;
    s->prims.push_back(r);
  }
  s->committed = false;
  return PT_OK;
}

int pt_scene_set_materials(pt_scene* s, const pt_material* m, int32_t n) {
  if (!s || n < 0 || (n > 0 && !m)) return PT_ERR_INVALID;
  for (int i = 0; i < n; i++)
    if (m[i].type < 0 || m[i].type >= PT_MAT_TYPE_COUNT) {
      pt_set_error("material %d: unknown type %d", i, m[i].type);
      return PT_ERR_INVALID;
    }
  s->materials.assign(m, m + n);
  s->committed = false;
  return PT_OK;
}
int pt_scene_set_textures(pt_scene* s, const pt_texture* t, int32_t n) {
  if (!s || n < 0 || (n > 0 && !t)) return PT_ERR_INVALID;
  s->textures.assign(t, t + n);
  s->committed = false;
  return PT_OK;
}
int pt_scene_set_texture_vector(pt_scene* s, const int32_t* ids, int32_t n) {
  if (!s || n < 0 || (n > 0 && !ids)) return PT_ERR_INVALID;
  s->texvec.assign(ids, ids + n);
  s->committed = false;
  return PT_OK;
}
int pt_scene_set_texture_data(pt_scene* s, const float* data, int32_t n) {
  if (!s || n < 0 || (n > 0 && !data)) return PT_ERR_INVALID;
  s->texdata.assign(data, data + n);
  s->committed = false;
  return PT_OK;
}
int pt_scene_add_image(pt_scene* s, const float* rgba, int32_t w, int32_t h, int32_t* image_id) {
  if (!s || !rgba || w <= 0 || h <= 0) {
    pt_set_error("pt_scene_add_image: bad arguments");
    return PT_ERR_INVALID;
  }
  HostImage im;
  im.width = w;
  im.height = h;
  im.rgba.assign(rgba, rgba + (size_t)w * h * 4);
  if (image_id) *image_id = (int32_t)s->images.size();
  s->images.push_back(std::move(im));
  s->committed = false;
  s->env_tables_valid = false;
  return PT_OK;
}
int pt_scene_set_lights(pt_scene* s, const pt_light* l, int32_t n) {
  if (!s || n < 0 || (n > 0 && !l)) return PT_ERR_INVALID;
  for (int i = 0; i < n; i++)
    if (!((l[i].type == PT_LIGHT_RECT && l[i].axis >= 0 && l[i].axis <= 2) || l[i].type == PT_LIGHT_SPHERE)) {
      pt_set_error("light %d: unknown light type %d / axis %d", i, l[i].type, l[i].axis);
      return PT_ERR_INVALID;
    }
  s->lights.assign(l, l + n);
  s->committed = false;
  return PT_OK;
}
int pt_scene_set_camera(pt_scene* s, const pt_camera* cam) {
  if (!s || !cam) return PT_ERR_INVALID;
  s->camera = *cam;
  if (s->committed) {
    // Cheap to change after commit - unless the bounds of moving spheres depend on it: their boxes cover the
    // shutter interval of the camera the BVH was built for (bvh_core.h prim_bounds), and a ray whose time lies
    // outside it would silently miss them.  A changed interval therefore needs a new commit.
    if (!s->moving_spheres.empty() && (cam->time0 != s->built_time0 || cam->time1 != s->built_time1)) s->committed = false;
    else s->dev.camera = *cam;
  }
  return PT_OK;
}
int pt_scene_set_miss(pt_scene* s, const pt_miss* miss) {
  if (!s || !miss) return PT_ERR_INVALID;
  if (miss->type < PT_MISS_GRADIENT || miss->type > PT_MISS_HDR) {
    pt_set_error("Parameters invalid, miss program unknown or not yet implemented");  // programs.hpp:59,89
    return PT_ERR_INVALID;
  }
  const bool tables_change = miss->env_nee != s->miss.env_nee || miss->image != s->miss.image ||
                             miss->type != s->miss.type || miss->spherical != s->miss.spherical;
  const bool had_tables = s->dev.env_n > 0;
  s->miss = *miss;
  if (tables_change) {
    s->env_tables_valid = false;
    // the alias tables are built by commit; a committed scene that has them (or now needs them) must not keep
    // rendering with stale ones (env_nee_active() keys on dev.env_n)
    if (miss->env_nee || had_tables) s->committed = false;
  }
  if (s->committed) s->dev.miss = *miss;
  return PT_OK;
}

int pt_scene_commit(pt_scene* s) {
  if (!s) return PT_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    pt_set_error("no CUDA device: libptb200 has no CPU fallback");
    cudaGetLastError();
    return PT_ERR_CUDA;
  }
  s->committed = false;  // a commit that fails midway leaves device buffers half replaced: renderers must refuse the scene
  PT_CUDA(cudaGetDevice(&s->device));
  const int nm = (int)s->materials.size();
  if (s->prims.empty()) {
    pt_set_error("scene has no primitives");
    return PT_ERR_INVALID;
  }
  s->material_type_mask = 0;
  for (const pt_prim_ref& p : s->prims) {
    if (p.material < 0 || p.material >= nm) {
      pt_set_error("primitive references material %d but only %d materials are set", p.material, nm);
      return PT_ERR_INVALID;
    }
    s->material_type_mask |= 1u << s->materials[p.material].type;
  }
  if (nm >= (1 << 24)) {
    pt_set_error("too many materials");
    return PT_ERR_INVALID;
  }
  for (const pt_texture& t : s->textures) {
    if (t.type == PT_TEX_VECTOR && (t.a < 0 || t.b < 0 || t.a + t.b > (int)s->texvec.size())) {
      pt_set_error("vector texture range outside the texture-vector table");
      return PT_ERR_INVALID;
    }
    if ((t.type == PT_TEX_NOISE && (t.a < 0 || t.a + 1536 > (int)s->texdata.size())) ||
        (t.type == PT_TEX_GRADIENT && (t.a < 0 || t.a + 9 > (int)s->texdata.size()))) {
      pt_set_error("noise/gradient texture data outside the texture-data table");
      return PT_ERR_INVALID;
    }
    if (t.type < PT_TEX_CONSTANT || t.type > PT_TEX_GRADIENT) {
      pt_set_error("unknown texture type %d", t.type);
      return PT_ERR_INVALID;
    }
  }
  int rc;
  if ((rc = upload_vec(s->d_prim_ref, s->prims))) return rc;
  if ((rc = upload_vec(s->d_spheres, s->spheres))) return rc;
  if ((rc = upload_vec(s->d_rects, s->rects))) return rc;
  if ((rc = upload_vec(s->d_boxes, s->boxes))) return rc;
  if ((rc = upload_vec(s->d_moving_spheres, s->moving_spheres))) return rc;
  if ((rc = upload_vec(s->d_volume_spheres, s->volume_spheres))) return rc;
  if ((rc = upload_vec(s->d_volume_boxes, s->volume_boxes))) return rc;
  if ((rc = upload_vec(s->d_tri_v, s->tri_v))) return rc;
  if ((rc = upload_vec(s->d_tri_n, s->tri_n))) return rc;
  if ((rc = upload_vec(s->d_tri_uv, s->tri_uv))) return rc;
  if ((rc = upload_vec(s->d_tri_tex, s->tri_tex))) return rc;
  if ((rc = upload_vec(s->d_xforms, s->xforms))) return rc;
  if ((rc = upload_vec(s->d_materials, s->materials))) return rc;
  if ((rc = upload_vec(s->d_textures, s->textures))) return rc;
  if ((rc = upload_vec(s->d_texvec, s->texvec))) return rc;
  if ((rc = upload_vec(s->d_texdata, s->texdata))) return rc;
  if ((rc = upload_vec(s->d_lights, s->lights))) return rc;
  for (auto* b : s->d_image_texels) delete b;
  s->d_image_texels.clear();
  std::vector<ImageDev> imgs;
  for (const HostImage& im : s->images) {
    DevBuf* b = new DevBuf();
    s->d_image_texels.push_back(b);
    PT_CUDA(b->alloc(im.rgba.size() * 4));
    PT_CUDA(cudaMemcpy(b->p, im.rgba.data(), im.rgba.size() * 4, cudaMemcpyHostToDevice));
    ImageDev d;
    d.texels = b->as<float4>();
    d.width = im.width;
    d.height = im.height;
    imgs.push_back(d);
  }
  if ((rc = upload_vec(s->d_images, imgs))) return rc;
  if ((rc = pt_env_tables_build(s))) return rc;  // GPU construction (env_alias.cu); sets dev.env_n / env_w / env_h

  if ((rc = pt_bvh_build(s))) return rc;
  s->built_time0 = s->camera.time0;
  s->built_time1 = s->camera.time1;

  SceneDev& d = s->dev;
  d.nodes8 = s->d_bvh.as<uint4>();
  d.leaves = (const LeafRec*)((const char*)s->d_bvh.p + s->bvh_leaves_off);
  d.prim_ref = s->d_prim_ref.as<pt_prim_ref>();
  d.prim_leaf = s->d_prim_leaf.as<int>();
  d.xforms = s->d_xforms.as<pt_xform>();
  d.tri_n = s->any_n ? s->d_tri_n.as<float>() : nullptr;
  d.tri_uv = s->any_uv ? s->d_tri_uv.as<float>() : nullptr;
  d.tri_tex = s->any_tex ? s->d_tri_tex.as<int>() : nullptr;
  d.materials = s->d_materials.as<pt_material>();
  d.textures = s->d_textures.as<pt_texture>();
  d.texvec = s->d_texvec.as<int>();
  d.texdata = s->d_texdata.as<float>();
  d.images = s->d_images.as<ImageDev>();
  d.lights = s->d_lights.as<pt_light>();
  d.env_alias_prob = s->d_env_prob.as<float>();
  d.env_alias_idx = s->d_env_idx.as<int>();
  d.env_texel_pdf = s->d_env_pdf.as<float>();
  d.tex_need_uv = 0;
  for (const pt_texture& t : s->textures) d.tex_need_uv |= t.type == PT_TEX_IMAGE;
  d.qbias = 0x47000000u;
  d.special_mask = (s->moving_spheres.empty() ? 0 : 1) | (s->volume_spheres.empty() && s->volume_boxes.empty() ? 0 : 2);
  d.n_prims = (int)s->prims.size();
  d.n_materials = nm;
  d.n_textures = (int)s->textures.size();
  d.n_texvec = (int)s->texvec.size();
  d.n_texdata = (int)s->texdata.size();
  d.n_images = (int)s->images.size();
  d.n_lights = (int)s->lights.size();
  d.camera = s->camera;
  d.miss = s->miss;
  s->commit_gen++;
  s->committed = true;
  return PT_OK;
}

int pt_scene_get_desc(const pt_scene* cs, pt_scene_desc* o) {
  if (!cs || !o) return PT_ERR_INVALID;
  pt_scene* s = const_cast<pt_scene*>(cs);
  memset(o, 0, sizeof(*o));
  o->n_prims = (int)s->prims.size();
  o->prims = s->prims.data();
  o->n_spheres = (int)s->spheres.size();
  o->spheres = s->spheres.data();
  o->n_rects = (int)s->rects.size();
  o->rects = s->rects.data();
  o->n_boxes = (int)s->boxes.size();
  o->boxes = s->boxes.data();
  o->n_tris = (int)(s->tri_v.size() / 9);
  o->tri_v = s->tri_v.data();
  o->tri_n = s->any_n ? s->tri_n.data() : nullptr;
  o->tri_uv = s->any_uv ? s->tri_uv.data() : nullptr;
  o->tri_tex = s->any_tex ? s->tri_tex.data() : nullptr;
  o->n_xforms = (int)s->xforms.size();
  o->xforms = s->xforms.data();
  o->n_materials = (int)s->materials.size();
  o->materials = s->materials.data();
  o->n_textures = (int)s->textures.size();
  o->textures = s->textures.data();
  o->n_texvec = (int)s->texvec.size();
  o->texvec = s->texvec.data();
  o->n_texdata = (int)s->texdata.size();
  o->texdata = s->texdata.data();
  s->image_descs.clear();
  for (const HostImage& im : s->images) s->image_descs.push_back(pt_image{im.width, im.height, im.rgba.data()});
  o->n_images = (int)s->image_descs.size();
  o->images = s->image_descs.data();
  o->n_lights = (int)s->lights.size();
  o->lights = s->lights.data();
  o->camera = s->camera;
  o->miss = s->miss;
  if (!s->env_tables_valid && pt_env_tables_fetch(s)) return PT_ERR_CUDA;
  o->env_n = (int)s->env_alias_prob.size();
  o->env_alias_prob = s->env_alias_prob.data();
  o->env_alias_idx = s->env_alias_idx.data();
  o->env_texel_pdf = s->env_texel_pdf.data();
  o->n_moving_spheres = (int)s->moving_spheres.size();
  o->moving_spheres = s->moving_spheres.data();
  o->n_volume_spheres = (int)s->volume_spheres.size();
  o->volume_spheres = s->volume_spheres.data();
  o->n_volume_boxes = (int)s->volume_boxes.size();
  o->volume_boxes = s->volume_boxes.data();
  return PT_OK;
}

int pt_scene_bvh_info(const pt_scene* s, pt_bvh_info* out) {
  if (!s || !out) return PT_ERR_INVALID;
  if (!s->committed) {
    pt_set_error("pt_scene_bvh_info: scene not committed");
    return PT_ERR_STATE;
  }
  *out = s->info;
  return PT_OK;
}

}  // extern "C"
