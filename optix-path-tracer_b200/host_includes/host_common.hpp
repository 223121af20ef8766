// host_common.hpp — App_State knobs, rnd(), Light_Sampler.
// Mirrors the reference surface host_includes/host_common.hpp:21-81 (field
// names and defaults verbatim so GUI/CLI code written against it keeps working).
#pragma once
#define _USE_MATH_DEFINES 1
#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <iostream>
#include <random>
#include <string>

#include "optix_shim.hpp"

#ifndef PI_F
#define PI_F 3.141592654f  // programs/common.hpp:18
#endif
typedef enum { X_AXIS, Y_AXIS, Z_AXIS } AXIS;  // programs/common.hpp:26

// programs/vec.hpp helpers the host code uses
inline float3 unit_vector(const float3& v) { return v / length(v); }
inline float ffmin(const float& a, const float& b) { return a < b ? a : b; }
inline float ffmax(const float& a, const float& b) { return a > b ? a : b; }
inline float saturate(float x) { return ffmax(0.f, ffmin(1.f, x)); }

struct App_State {
  App_State() {
    context = Context::create();
    W = H = 500;
    samples = 500;
    scene = 2;  // Cornell
    model = 0;
    frequency = 1;
    currentSample = 0;
    showProgress = true;
    RTX = true;  // kept for source compatibility; there are no RT cores to toggle
    russian = true;
    depth = 50;
    start = done = false;
    fileType = 0;  // PNG = 0, HDR = 1
    fileName = "out";
  }
  Context context;
  int W, H, samples, scene, currentSample, model, frequency, fileType, depth;
  bool done, start, showProgress, RTX, russian;
  Buffer accBuffer, displayBuffer;
  std::string fileName;
};

// scene-construction RNG: same generator and distribution as the reference
// (host_includes/host_common.hpp:59-63) so scene functions produce the same scene
inline float rnd() {
  static std::mt19937 gen(0);
  static std::uniform_real_distribution<float> dis(0.f, 1.f);
  return dis(gen);
}

struct Light_Sampler {
  std::vector<Program> sample, pdf;
  std::vector<float3> emissions;
};
