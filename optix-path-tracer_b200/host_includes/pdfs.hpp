// pdfs.hpp — light sampling descriptors (reference surface:
// host_includes/pdfs.hpp:12-109).  createSample/createPDF return Program
// handles that carry the light shape; setRayGenerationProgram pairs them with
// Light_Sampler::emissions.
#pragma once
#include "host_common.hpp"

struct PDF {
  virtual ~PDF() {}
  virtual Program createSample(Context& g_context) const = 0;
  virtual Program createPDF(Context& g_context) const = 0;
};

struct Rectangle_PDF : public PDF {
  Rectangle_PDF(const float aa0, const float aa1, const float bb0, const float bb1, const float kk, const AXIS aax)
      : a0(aa0), a1(aa1), b0(bb0), b1(bb1), k(kk), ax(aax) {}
  Program make() const {
    Program p{std::make_shared<ProgramObj>()};
    p->is_light = true;
    p->light.type = PT_LIGHT_RECT;
    p->light.axis = (int)ax;
    p->light.a0 = a0; p->light.a1 = a1; p->light.b0 = b0; p->light.b1 = b1; p->light.k = k;
    return p;
  }
  Program createSample(Context&) const override { return make(); }
  Program createPDF(Context&) const override { return make(); }
  float a0, a1, b0, b1, k;
  AXIS ax;
};

// sphere lights (reference: host_includes/pdfs.hpp:82-109; device side programs/pdfs/sphere_pdf.cu)
struct Sphere_PDF : public PDF {
  Sphere_PDF(const float3 c, const float r) : radius(r), center(c) {}
  Program make() const {
    Program p{std::make_shared<ProgramObj>()};
    p->is_light = true;
    p->light.type = PT_LIGHT_SPHERE;
    p->light.axis = 0;
    p->light.a0 = center.x; p->light.a1 = center.y; p->light.b0 = center.z; p->light.b1 = 0.f; p->light.k = radius;
    return p;
  }
  Program createSample(Context&) const override { return make(); }
  Program createPDF(Context&) const override { return make(); }
  float radius;
  float3 center;
};
