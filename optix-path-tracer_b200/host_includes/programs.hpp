// programs.hpp — ray-generation / miss / exception program setup (reference
// surface: host_includes/programs.hpp:17-97).  There are no programs to JIT;
// these record lights and the miss mode into the context.
#pragma once
#include "buffers.hpp"
#include "host_common.hpp"
#include "pdfs.hpp"
#include "textures.hpp"

inline void setRayGenerationProgram(Context& g_context, Light_Sampler& lights) {
  if (lights.sample.size() != lights.emissions.size() || lights.pdf.size() != lights.emissions.size())
    throw "Light_Sampler: sample/pdf/emissions sizes differ";
  g_context->lights.clear();
  for (size_t i = 0; i < lights.emissions.size(); i++) {
    if (!lights.pdf[i] || !lights.pdf[i]->is_light) throw "Light_Sampler entry is not a light PDF";
    pt_light l = lights.pdf[i]->light;
    l.emission[0] = lights.emissions[i].x;
    l.emission[1] = lights.emissions[i].y;
    l.emission[2] = lights.emissions[i].z;
    g_context->lights.push_back(l);
  }
  g_context->dirty = true;
}

typedef enum { GRADIENT, CONSTANT, IMG, HDR } Miss_Programs;

// image miss programs (programs/miss.cu:51-95)
inline void setMissProgram(Context& g_context, Miss_Programs id, std::string fileName, bool isSpherical = true) {
  pt_miss m{};
  m.image = -1;
  m.spherical = isSpherical ? 1 : 0;
  ContextObj::Image im;
  if (id == IMG) {
    m.type = PT_MISS_IMG;
    if (!ptio::load_ldr_image(fileName, im.w, im.h, im.rgba, true)) throw "cannot load " + fileName;
  } else if (id == HDR) {
    m.type = PT_MISS_HDR;
    if (!ptio::load_radiance_hdr(fileName, im.w, im.h, im.rgba)) throw "cannot load " + fileName;
  } else {
    throw "Parameters invalid, miss program unknown or not yet implemented";
  }
  m.image = (int)g_context->images.size();
  g_context->images.push_back(std::move(im));
  g_context->miss = m;
  g_context->dirty = true;
}

// colour miss programs (programs/miss.cu:30-46)
inline void setMissProgram(Context& g_context, Miss_Programs id, float3 colorValue1 = make_float3(0.f),
                           float3 colorValue2 = make_float3(0.f)) {
  pt_miss m{};
  m.image = -1;
  m.spherical = 1;
  if (id == GRADIENT) m.type = PT_MISS_GRADIENT;
  else if (id == CONSTANT) m.type = PT_MISS_CONSTANT;
  else throw "Parameters invalid, miss program unknown or not yet implemented";
  m.color1[0] = colorValue1.x; m.color1[1] = colorValue1.y; m.color1[2] = colorValue1.z;
  m.color2[0] = colorValue2.x; m.color2[1] = colorValue2.y; m.color2[2] = colorValue2.z;
  g_context->miss = m;
  g_context->dirty = true;
}

// OptiX exception program (programs/exception.cu): out of scope; CUDA errors surface
// through the C ABI return codes instead
inline void setExceptionProgram(Context&) {}
