// optix_shim.hpp — the sliver of optixpp / optixu that the reference's scene
// code touches (Context, Group, Acceleration, Program, Buffer, float3 math,
// Matrix4x4), re-implemented over the ptb200 C ABI.  No OptiX anywhere: B200
// has no RT cores and the SDK is not present; acceleration structures,
// traversal and program dispatch live in libptb200.so (csrc/).
//
// Reference surface mirrored here: Context::create(),
// ctx->createGroup(), ctx->createAcceleration("Trbvh"), group->setAcceleration,
// ctx["world"]->set(group), ctx["name"]->setInt/setFloat/set3fv/set(Buffer),
// ctx->validate(), ctx->launch(0,W,H), Buffer->map()/unmap()
// (host_includes/scenes.hpp:31-32,77; main.cpp:30-33,57-63,92-97).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ptb200.h"

#ifndef __VECTOR_TYPES_H__
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int3 { int x, y, z; };
struct uint3 { unsigned x, y, z; };
struct uchar4 { unsigned char x, y, z, w; };
#endif
typedef unsigned int uint;

namespace optix {
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float3 make_float3(float s) { return {s, s, s}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint3 make_uint3(unsigned x, unsigned y, unsigned z) { return {x, y, z}; }
inline float3 operator+(const float3& a, const float3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(const float3& a, const float3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(const float3& a) { return {-a.x, -a.y, -a.z}; }
inline float3 operator*(const float3& a, const float3& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator*(const float3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, const float3& a) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }
inline float3 operator/(const float3& a, const float3& b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3& operator+=(float3& a, const float3& b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float s) { a = a * s; return a; }
inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(const float3& a, const float3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(const float3& v) { return sqrtf(dot(v, v)); }
inline float3 normalize(const float3& v) { return v * (1.0f / sqrtf(dot(v, v))); }
inline float clamp(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
inline float3 clamp(const float3& v, float a, float b) { return {clamp(v.x, a, b), clamp(v.y, a, b), clamp(v.z, a, b)}; }

// row-major 4x4 with the optixu Matrix4x4 factory semantics
struct Matrix4x4 {
  float m[16];
  static Matrix4x4 identity() {
    Matrix4x4 r;
    for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.f : 0.f;
    return r;
  }
  static Matrix4x4 translate(const float3& t) {
    Matrix4x4 r = identity();
    r.m[3] = t.x; r.m[7] = t.y; r.m[11] = t.z;
    return r;
  }
  static Matrix4x4 scale(const float3& s) {
    Matrix4x4 r = identity();
    r.m[0] = s.x; r.m[5] = s.y; r.m[10] = s.z;
    return r;
  }
  static Matrix4x4 rotate(float radians, const float3& axis) {
    Matrix4x4 r = identity();
    float3 u = normalize(axis);
    float s = sinf(radians), c = cosf(radians), ic = 1.f - c;
    r.m[0] = u.x * u.x + c * (1 - u.x * u.x); r.m[1] = u.x * u.y * ic - u.z * s; r.m[2] = u.z * u.x * ic + u.y * s;
    r.m[4] = u.x * u.y * ic + u.z * s; r.m[5] = u.y * u.y + c * (1 - u.y * u.y); r.m[6] = u.y * u.z * ic - u.x * s;
    r.m[8] = u.z * u.x * ic - u.y * s; r.m[9] = u.y * u.z * ic + u.x * s; r.m[10] = u.z * u.z + c * (1 - u.z * u.z);
    return r;
  }
  Matrix4x4 operator*(const Matrix4x4& b) const {
    Matrix4x4 r;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        double s = 0;
        for (int k = 0; k < 4; k++) s += (double)m[4 * i + k] * b.m[4 * k + j];
        r.m[4 * i + j] = (float)s;
      }
    return r;
  }
  bool isIdentity() const {
    Matrix4x4 i = identity();
    return memcmp(m, i.m, sizeof(m)) == 0;
  }
  const float* getData() const { return m; }
};

// ---------------------------------------------------------------------------------------------
// recorded scene + lazily created device objects
// ---------------------------------------------------------------------------------------------
class ContextObj;
typedef std::shared_ptr<ContextObj> ContextPtr;

// Program: what Texture::assignTo / PDF::create* hand back in the reference
// (an rtProgram).  Here: an id into the context's texture table or a light record.
struct ProgramObj {
  int texture_id = -1;
  pt_light light{};
  bool is_light = false;
};
struct Program {
  std::shared_ptr<ProgramObj> p;
  ProgramObj* operator->() const { return p.get(); }
  explicit operator bool() const { return (bool)p; }
};
struct Material {  // reference: optix::Material; here: id into the material table
  int id = -1;
};
struct AccelerationObj {
  std::string builder;
  void markDirty() {}
};
struct Acceleration {
  std::shared_ptr<AccelerationObj> p;
  AccelerationObj* operator->() const { return p.get(); }
};
struct GroupObj {
  Acceleration accel;
  void setAcceleration(Acceleration a) { accel = a; }
  Acceleration getAcceleration() {
    if (!accel.p) accel.p = std::make_shared<AccelerationObj>();
    return accel;
  }
};
struct Group {
  std::shared_ptr<GroupObj> p;
  GroupObj* operator->() const { return p.get(); }
};

struct BufferObj {  // accumulation (float4) or display (uchar4) buffer bound to a context
  ContextObj* ctx = nullptr;
  int W = 0, H = 0;
  bool is_accum = true;
  std::vector<float> host;
  std::vector<unsigned char> host8;
  void* map();
  void unmap() {}
  // B200-first extension (checkpoint / resume): overwrite the device accumulation buffer with
  // W*H float4 saved from an earlier map(); needs a launch (e.g. launch(0,0,0) + renderer) first
  void restore(const float* rgba, int Wr, int Hr);
};
struct Buffer {
  std::shared_ptr<BufferObj> p;
  BufferObj* operator->() const { return p.get(); }
};

class VariableObj {
 public:
  VariableObj(ContextObj* c, std::string n) : ctx(c), name(std::move(n)) {}
  void setInt(int v);
  void setFloat(float v);
  void set3fv(const float* v);
  void set(Group g);
  void set(Buffer b);
 private:
  ContextObj* ctx;
  std::string name;
};
struct Variable {
  std::shared_ptr<VariableObj> p;
  VariableObj* operator->() const { return p.get(); }
};

class ContextObj {
 public:
  // recorded scene
  std::vector<pt_texture> textures;
  std::vector<int32_t> texvec;
  std::vector<float> texdata;  // noise tables / gradient colours (pt_scene_set_texture_data)
  std::vector<pt_material> materials;
  std::vector<pt_light> lights;
  struct Image { int w, h; std::vector<float> rgba; };
  std::vector<Image> images;
  struct Batch {  // primitives in insertion order
    int kind;
    pt_sphere s; pt_rect r; pt_box b;
    pt_moving_sphere ms; pt_volume_sphere vs; pt_volume_box vb;
    std::vector<float> v, n, uv; std::vector<int32_t> tex;
    float xf[12]; int material;
  };
  std::vector<Batch> batches;
  pt_camera camera{};
  pt_miss miss{};
  std::map<const void*, int> texture_cache, material_cache;
  std::map<std::string, int> ints;
  std::map<std::string, float> floats;
  bool has_world = false;

  // device
  pt_scene* scene = nullptr;
  pt_renderer* renderer = nullptr;
  int rW = 0, rH = 0;
  bool dirty = true;

  ContextObj() {
    miss.type = PT_MISS_CONSTANT;
    miss.image = -1;
    miss.spherical = 1;
  }
  ~ContextObj() {
    if (renderer) pt_renderer_destroy(renderer);
    if (scene) pt_scene_destroy(scene);
  }
  Group createGroup() { dirty = true; return Group{std::make_shared<GroupObj>()}; }
  Acceleration createAcceleration(const char* builder) {
    auto a = std::make_shared<AccelerationObj>();
    a->builder = builder;  // "Trbvh" in the reference; here always the LBVH+SAH -> BVH8 builder
    return Acceleration{a};
  }
  void setRayTypeCount(int) {}
  void setMaxTraceDepth(int) {}
  void setEntryPointCount(int) {}
  void validate() {
    if (!has_world) throw "context has no world";
  }
  static void check(int rc) {
    if (rc != 0) throw std::string("ptb200: ") + pt_last_error();
  }
  // (re)build the device scene: the reference's first launch ("OptiX Building Time")
  void commit() {
    if (scene) { if (renderer) { pt_renderer_destroy(renderer); renderer = nullptr; } pt_scene_destroy(scene); scene = nullptr; }
    check(pt_scene_create(&scene));
    check(pt_scene_set_textures(scene, textures.data(), (int)textures.size()));
    check(pt_scene_set_texture_vector(scene, texvec.data(), (int)texvec.size()));
    check(pt_scene_set_texture_data(scene, texdata.data(), (int)texdata.size()));
    check(pt_scene_set_materials(scene, materials.data(), (int)materials.size()));
    for (auto& im : images) check(pt_scene_add_image(scene, im.rgba.data(), im.w, im.h, nullptr));
    check(pt_scene_set_lights(scene, lights.data(), (int)lights.size()));
    for (auto& b : batches) {
      if (b.kind == PT_PRIM_SPHERE) check(pt_scene_add_spheres(scene, &b.s, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_RECT) check(pt_scene_add_rects(scene, &b.r, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_BOX) check(pt_scene_add_boxes(scene, &b.b, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_MOVING_SPHERE) check(pt_scene_add_moving_spheres(scene, &b.ms, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_VOLUME_SPHERE) check(pt_scene_add_volume_spheres(scene, &b.vs, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_VOLUME_BOX) check(pt_scene_add_volume_boxes(scene, &b.vb, b.xf, &b.material, 1, nullptr));
      else
        check(pt_scene_add_mesh(scene, b.v.data(), b.n.empty() ? nullptr : b.n.data(), b.uv.empty() ? nullptr : b.uv.data(),
                                b.tex.empty() ? nullptr : b.tex.data(), (int)(b.v.size() / 9), b.xf, b.material, nullptr));
    }
    check(pt_scene_set_camera(scene, &camera));
    check(pt_scene_set_miss(scene, &miss));
    check(pt_scene_commit(scene));
    dirty = false;
  }
  // flat scene without touching the GPU (tests)
  pt_scene* buildHostScene() {
    pt_scene* s = nullptr;
    check(pt_scene_create(&s));
    check(pt_scene_set_textures(s, textures.data(), (int)textures.size()));
    check(pt_scene_set_texture_vector(s, texvec.data(), (int)texvec.size()));
    check(pt_scene_set_texture_data(s, texdata.data(), (int)texdata.size()));
    check(pt_scene_set_materials(s, materials.data(), (int)materials.size()));
    for (auto& im : images) check(pt_scene_add_image(s, im.rgba.data(), im.w, im.h, nullptr));
    check(pt_scene_set_lights(s, lights.data(), (int)lights.size()));
    for (auto& b : batches) {
      if (b.kind == PT_PRIM_SPHERE) check(pt_scene_add_spheres(s, &b.s, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_RECT) check(pt_scene_add_rects(s, &b.r, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_BOX) check(pt_scene_add_boxes(s, &b.b, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_MOVING_SPHERE) check(pt_scene_add_moving_spheres(s, &b.ms, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_VOLUME_SPHERE) check(pt_scene_add_volume_spheres(s, &b.vs, b.xf, &b.material, 1, nullptr));
      else if (b.kind == PT_PRIM_VOLUME_BOX) check(pt_scene_add_volume_boxes(s, &b.vb, b.xf, &b.material, 1, nullptr));
      else
        check(pt_scene_add_mesh(s, b.v.data(), b.n.empty() ? nullptr : b.n.data(), b.uv.empty() ? nullptr : b.uv.data(),
                                b.tex.empty() ? nullptr : b.tex.data(), (int)(b.v.size() / 9), b.xf, b.material, nullptr));
    }
    check(pt_scene_set_camera(s, &camera));
    check(pt_scene_set_miss(s, &miss));
    return s;
  }
  // optixpp's Context::setDevices(begin, end) (OptiX 6 spreads a launch over the listed devices by itself);
  // here the frames of a launch are split over them and the accumulation buffers merged by one NCCL reduce
  // (pt_renderer_create_multi).  An empty list or one entry = the current device.
  std::vector<int> devices;
  template <class It>
  void setDevices(It begin, It end) {
    devices.assign(begin, end);
    if (renderer) { pt_renderer_destroy(renderer); renderer = nullptr; }
  }
  void ensureRenderer(int W, int H) {
    if (dirty || !scene) commit();
    if (!renderer || rW != W || rH != H) {
      if (renderer) pt_renderer_destroy(renderer);
      renderer = nullptr;
      if (devices.size() > 1) check(pt_renderer_create_multi(scene, W, H, devices.data(), (int)devices.size(), &renderer));
      else check(pt_renderer_create(scene, W, H, nullptr, &renderer));
      rW = W;
      rH = H;
    }
  }
  // reference: one launch == one sample per pixel of frame `frame` (main.cpp:264-265).
  // launch(0,0,0) is the reference's build-only launch (main.cpp:99).
  void launch(int /*entry*/, int W, int H) {
    if (W <= 0 || H <= 0) {
      if (dirty || !scene) commit();
      return;
    }
    launchFrames(ints.count("frame") ? ints["frame"] : 0, 1, W, H);
  }
  // B200-first extension: many frames per call keep the path pool full
  void launchFrames(int first_frame, int nframes, int W, int H) {
    ensureRenderer(W, H);
    int depth = ints.count("maxDepth") ? ints["maxDepth"] : 50;
    int russian = ints.count("russian") ? ints["russian"] : 1;
    check(pt_render(renderer, first_frame, nframes, depth, russian, 0));
  }
  Variable operator[](const std::string& name) { return Variable{std::make_shared<VariableObj>(this, name)}; }
};

inline void VariableObj::setInt(int v) { ctx->ints[name] = v; }
inline void VariableObj::setFloat(float v) { ctx->floats[name] = v; }
inline void VariableObj::set3fv(const float* v) {
  float* dst = nullptr;
  pt_camera& c = ctx->camera;
  if (name == "camera_lower_left_corner") dst = c.lower_left_corner;
  else if (name == "camera_horizontal") dst = c.horizontal;
  else if (name == "camera_vertical") dst = c.vertical;
  else if (name == "camera_origin") dst = c.origin;
  else if (name == "camera_u") dst = c.u;
  else if (name == "camera_v") dst = c.v;
  else if (name == "camera_w") dst = c.w;
  if (dst) memcpy(dst, v, 12);
}
inline void VariableObj::set(Group) {
  if (name == "world") ctx->has_world = true;
}
inline void VariableObj::set(Buffer b) { b->ctx = ctx; }
inline void BufferObj::restore(const float* rgba, int Wr, int Hr) {
  if (!ctx || !is_accum) throw "Buffer::restore: not an accumulation buffer bound to a context";
  ctx->ensureRenderer(Wr, Hr);
  ContextObj::check(pt_write_accum(ctx->renderer, rgba));
}
inline void* BufferObj::map() {
  if (!ctx || !ctx->renderer) throw "Buffer::map before the first launch";
  if (is_accum) {
    host.resize((size_t)W * H * 4);
    ContextObj::check(pt_read_accum(ctx->renderer, host.data()));
    return host.data();
  }
  // display buffer: make_Color of raygen.cu:136-146, converted on the device
  host8.resize((size_t)W * H * 4);
  ContextObj::check(pt_read_display(ctx->renderer, host8.data()));
  return host8.data();
}

// Context handle: `Context ctx = Context::create(); ctx->...; ctx["name"]->...`
struct Context {
  ContextPtr p;
  static Context create() { return Context{std::make_shared<ContextObj>()}; }
  ContextObj* operator->() const { return p.get(); }
  Variable operator[](const std::string& name) const { return (*p)[name]; }
  Variable operator[](const char* name) const { return (*p)[std::string(name)]; }
};

}  // namespace optix

using namespace optix;
