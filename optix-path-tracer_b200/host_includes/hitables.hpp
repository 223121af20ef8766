// hitables.hpp — host geometry classes (reference surface:
// host_includes/hitables.hpp:21-528).  Each hitable records one primitive (with
// its composed object->world matrix and material id) into the context.
#pragma once
#include "host_common.hpp"
#include "materials.hpp"
#include "programs.hpp"
#include "transforms.hpp"

class Hitable {
 public:
  std::vector<TransformParameter> transforms;
  Hitable(BRDF* material) : material(material) {}
  virtual ~Hitable() {}

  // records the primitive under transform `o2w` (replaces getGeometryInstance + addAndTransform)
  virtual void emit(Context& g_context, const Matrix4x4& o2w) = 0;

  virtual void rotate(float angle, AXIS axis) {
    transforms.push_back(TransformParameter(Rotate_Transform, angle, axis, make_float3(0.f), make_float3(0.f)));
  }
  virtual void scale(float3 scale) {
    transforms.push_back(TransformParameter(Scale_Transform, 0.f, X_AXIS, scale, make_float3(0.f)));
  }
  virtual void translate(float3 pos) {
    transforms.push_back(TransformParameter(Translate_Transform, 0.f, X_AXIS, make_float3(0.f), pos));
  }
  // hitables.hpp:61-70: reverses the vector first, so the FIRST call is innermost
  virtual void addTo(Group& /*d_world*/, Context& g_context) {
    std::reverse(transforms.begin(), transforms.end());
    emit(g_context, composeTransforms(transforms));
  }

 protected:
  BRDF* material;
  ContextObj::Batch& newBatch(Context& ctx, int kind, const Matrix4x4& o2w) {
    ctx->batches.emplace_back();
    ContextObj::Batch& b = ctx->batches.back();
    b.kind = kind;
    b.material = material->assignTo(ctx).id;
    toRowMajor3x4(o2w, b.xf);
    ctx->dirty = true;
    return b;
  }
};

class Sphere : public Hitable {
 public:
  Sphere(const float3& c, const float r, BRDF* material) : Hitable(material), center(c), radius(r) {}
  void emit(Context& ctx, const Matrix4x4& o2w) override {
    ContextObj::Batch& b = newBatch(ctx, PT_PRIM_SPHERE, o2w);
    b.s = pt_sphere{{center.x, center.y, center.z}, radius};
  }

 protected:
  const float3 center;
  const float radius;
};

/* Axis aligned rectangle.  k is the coordinate on `ax`; (a0,a1,b0,b1) span
   X_AXIS: (y,z)   Y_AXIS: (x,z)   Z_AXIS: (x,y)   (hitables.hpp:208-215) */
class AARect : public Hitable {
 public:
  AARect(const float a0, const float a1, const float b0, const float b1, const float k, const bool flip, const AXIS ax,
         BRDF* material)
      : Hitable(material), a0(a0), a1(a1), b0(b0), b1(b1), k(k), axis(ax), flip(flip) {}
  void emit(Context& ctx, const Matrix4x4& o2w) override {
    ContextObj::Batch& b = newBatch(ctx, PT_PRIM_RECT, o2w);
    b.r = pt_rect{a0, a1, b0, b1, k, (int)axis, flip ? 1 : 0};
  }

 protected:
  const float a0, a1, b0, b1, k;
  const AXIS axis;
  const bool flip;
};

class Box : public Hitable {
 public:
  Box(const float3& p0, const float3& p1, BRDF* material) : Hitable(material), p0(p0), p1(p1) {}
  void emit(Context& ctx, const Matrix4x4& o2w) override {
    ContextObj::Batch& b = newBatch(ctx, PT_PRIM_BOX, o2w);
    b.b = pt_box{{p0.x, p0.y, p0.z}, {p1.x, p1.y, p1.z}};
  }

 protected:
  const float3 p0, p1;
};

// Sphere whose centre moves linearly from c0 (time t0) to c1 (time t1); a sample sees it at its
// shutter time prd.time (reference: host_includes/hitables.hpp:131-173, constructor argument
// order c0, c1, r, t0, t1 as there).
class Moving_Sphere : public Hitable {
 public:
  Moving_Sphere(const float3& c0, const float3& c1, const float r, const float t0, const float t1, BRDF* material)
      : Hitable(material), center0(c0), center1(c1), radius(r), time0(t0), time1(t1) {}
  void emit(Context& ctx, const Matrix4x4& o2w) override {
    ContextObj::Batch& b = newBatch(ctx, PT_PRIM_MOVING_SPHERE, o2w);
    b.ms = pt_moving_sphere{{center0.x, center0.y, center0.z}, {center1.x, center1.y, center1.z}, radius, time0, time1};
  }

 protected:
  const float3 center0, center1;
  const float radius, time0, time1;
};
// constant-density media (reference: host_includes/hitables.hpp:175-217, 316-350)
class Volumetric_Sphere : public Hitable {
 public:
  Volumetric_Sphere(const float3& c, const float r, const float d, BRDF* material)
      : Hitable(material), center(c), radius(r), density(d) {}
  void emit(Context& ctx, const Matrix4x4& o2w) override {
    ContextObj::Batch& b = newBatch(ctx, PT_PRIM_VOLUME_SPHERE, o2w);
    b.vs = pt_volume_sphere{{center.x, center.y, center.z}, radius, density};
  }

 protected:
  const float3 center;
  const float radius, density;
};
class Volumetric_Box : public Hitable {
 public:
  Volumetric_Box(const float3& p0, const float3& p1, const float d, BRDF* material)
      : Hitable(material), p0(p0), p1(p1), density(d) {}
  void emit(Context& ctx, const Matrix4x4& o2w) override {
    ContextObj::Batch& b = newBatch(ctx, PT_PRIM_VOLUME_BOX, o2w);
    b.vb = pt_volume_box{{p0.x, p0.y, p0.z}, {p1.x, p1.y, p1.z}, density};
  }

 protected:
  const float3 p0, p1;
  const float density;
};
// Out of scope (broken in the reference); declared so scene code that
// mentions them still compiles, rejected when added to a scene.
class Triangle : public Hitable {
 public:
  Triangle(const float3&, const float2&, const float3&, const float2&, const float3&, const float2&, BRDF* m) : Hitable(m) {}
  void emit(Context&, const Matrix4x4&) override { throw "single Triangle hitable is broken in the reference; use Mesh"; }
};
class Cylinder : public Hitable {
 public:
  Cylinder(const float3&, const float&, const float, BRDF* m) : Hitable(m) {}
  void emit(Context&, const Matrix4x4&) override { throw "Cylinder is out of scope"; }
};

class Hitable_List {
 public:
  Hitable_List() {}
  int push(Hitable* hit) {
    hitList.push_back(hit);
    return (int)hitList.size() - 1;
  }
  void rotate(float angle, AXIS axis) {
    transforms.push_back(TransformParameter(Rotate_Transform, angle, axis, make_float3(0.f), make_float3(0.f)));
  }
  void scale(float3 scale) {
    transforms.push_back(TransformParameter(Scale_Transform, 0.f, X_AXIS, scale, make_float3(0.f)));
  }
  void translate(float3 pos) {
    transforms.push_back(TransformParameter(Translate_Transform, 0.f, X_AXIS, make_float3(0.f), pos));
  }
  // hitables.hpp:510-515: the list's transforms wrap a GeometryGroup of untransformed elements
  void addListTo(Group& /*d_world*/, Context& g_context) {
    Matrix4x4 m = composeTransforms(transforms);
    for (Hitable* h : hitList) h->emit(g_context, m);
  }
  // hitables.hpp:518-523: element transforms are NOT reversed: the LAST call is innermost
  void addElementsTo(Group& /*d_world*/, Context& g_context) {
    for (Hitable* h : hitList) h->emit(g_context, composeTransforms(h->transforms));
  }

 protected:
  std::vector<Hitable*> hitList;
  std::vector<TransformParameter> transforms;
};
