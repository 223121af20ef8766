// transforms.hpp — transform parameters (reference surface:
// host_includes/transforms.hpp:28-47) and their composition.  The reference
// nests one OptiX Transform node per parameter, popping from the BACK of the
// vector (transforms.hpp:302-303), so the last element is the innermost node:
//     object->world = P[0] * P[1] * ... * P[n-1].
// Here the chain is composed into one matrix per primitive.
#pragma once
#include "host_common.hpp"
#include "materials.hpp"

typedef enum { Rotate_Transform, Translate_Transform, Scale_Transform } Transform_Type;

struct TransformParameter {
  TransformParameter(Transform_Type t, float a, AXIS ax, float3 s, float3 p) : type(t), angle(a), axis(ax), scale(s), pos(p) {}
  Transform_Type type;
  float angle;
  AXIS axis;
  float3 scale;
  float3 pos;
};

inline Matrix4x4 transformMatrix(const TransformParameter& p) {
  switch (p.type) {
    case Rotate_Transform: {
      float3 ax = p.axis == X_AXIS ? make_float3(1.f, 0.f, 0.f) : (p.axis == Y_AXIS ? make_float3(0.f, 1.f, 0.f) : make_float3(0.f, 0.f, 1.f));
      return Matrix4x4::rotate(p.angle * PI_F / 180.f, ax);
    }
    case Translate_Transform:
      return Matrix4x4::translate(p.pos);
    case Scale_Transform:
      return Matrix4x4::scale(p.scale);
  }
  throw "Invalid Transform operation";
}

inline Matrix4x4 composeTransforms(const std::vector<TransformParameter>& params) {
  Matrix4x4 m = Matrix4x4::identity();
  for (const TransformParameter& p : params) m = m * transformMatrix(p);
  return m;
}

inline void toRowMajor3x4(const Matrix4x4& m, float out[12]) { memcpy(out, m.m, 12 * sizeof(float)); }
