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

// The knobs of the reference application (surface: host_includes/host_common.hpp:21-45).  Scene functions and
// application code address these members by name, so names and default values are the contract; the defaults
// are member initialisers here and the context is created on construction.
struct App_State {
  // what to render
  int scene = 2;             // 0 InOneWeekend, 1 MovingSpheres, 2 Cornell, 3 Final_Next_Week, 4 Test_Scene, 5.. ours
  int model = 0;             // mesh selector of the mesh scenes
  int W = 500;
  int H = 500;
  int samples = 500;         // samples per pixel of the whole job
  int depth = 50;            // maximum path length
  bool russian = true;       // Russian roulette past depth 10
  bool RTX = true;           // source compatibility only: a B200 has no RT cores to toggle
  // progress of the running job
  int currentSample = 0;
  int frequency = 1;         // preview period of the reference GUI
  bool showProgress = true;
  bool start = false;
  bool done = false;
  // output
  int fileType = 0;          // 0 = PNG, 1 = Radiance HDR
  std::string fileName = "out";
  Buffer accBuffer;
  Buffer displayBuffer;
  Context context = Context::create();
};

// Scene-construction random numbers.  The reference's scene functions draw from std::mt19937 seeded with 0
// through uniform_real_distribution<float>(0, 1) (host_includes/host_common.hpp:59-63); the same generator and
// distribution objects give the same stream, hence the same scenes (pinned: tests/test_oracle_host_golden.py).
namespace ptb200_detail {
struct SceneRng {
  std::mt19937 engine{0u};
  std::uniform_real_distribution<float> unit{0.f, 1.f};
  float next() { return unit(engine); }
};
inline SceneRng& scene_rng() {
  static SceneRng r;
  return r;
}
}  // namespace ptb200_detail
inline float rnd() { return ptb200_detail::scene_rng().next(); }

struct Light_Sampler {
  std::vector<Program> sample, pdf;
  std::vector<float3> emissions;
};
