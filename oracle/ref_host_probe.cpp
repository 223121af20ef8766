// ref_host_probe.cpp — the REFERENCE's host-side camera set-up (host_includes/camera.hpp) and scene RNG
// (host_includes/host_common.hpp rnd(): std::mt19937(0) + uniform_real_distribution<float>) compiled UNMODIFIED
// against a stub of the few optixpp types those two headers name.  TEST INFRASTRUCTURE ONLY
// (tests/test_oracle_host_golden.py): pins the camera vectors the facade / test helpers compute and the random
// stream the scene functions consume.
#define saturate shim_saturate  // host_common.hpp defines its own saturate()
#include <optix.h>
#undef saturate
#include <cstdio>
#define sprintf_s(buf, ...) snprintf(buf, sizeof(buf), __VA_ARGS__)  // STBI_MSC_SECURE_CRT path of stb_image_write.h

#include <map>
#include <memory>
#include <string>
#include <vector>

// ---- stub of the optixpp surface host_common.hpp / camera.hpp touch --------------------------------------
#define RT_SUCCESS 0
namespace optix {
struct ProgramObj {
  int get() { return 0; }
};
typedef std::shared_ptr<ProgramObj> Program;
struct BufferObj {};
typedef std::shared_ptr<BufferObj> Buffer;
struct VariableObj {
  std::vector<float> v;
  void set3fv(const float* p) { v.assign(p, p + 3); }
  void setFloat(float f) { v.assign(1, f); }
};
struct ContextObj {
  std::map<std::string, std::shared_ptr<VariableObj>> vars;
  Program createProgramFromPTXString(const char*, const std::string&) { return std::make_shared<ProgramObj>(); }
};
struct Context {
  std::shared_ptr<ContextObj> p;
  static Context create() { return Context{std::make_shared<ContextObj>()}; }
  ContextObj* operator->() { return p.get(); }
  std::shared_ptr<VariableObj> operator[](const std::string& name) {
    auto& v = p->vars[name];
    if (!v) v = std::make_shared<VariableObj>();
    return v;
  }
};
}  // namespace optix
static inline int rtProgramValidate(int) { return RT_SUCCESS; }
using namespace optix;

#include "camera.hpp"  // pulls in host_common.hpp (and through it vec.hpp, stb_image, HDRloader.h)

extern "C" {
// out21: origin, lower_left_corner, horizontal, vertical, u, v, w (3 floats each; the order pt_camera uses), then
// lens_radius, time0, time1 — exactly what Camera::set hands to the ray-generation program
void refhost_camera(const float* from, const float* at, const float* up, float vfov, float aspect, float aperture,
                    float focus, float t0, float t1, float* out24) {
  Camera cam(make_float3(from[0], from[1], from[2]), make_float3(at[0], at[1], at[2]), make_float3(up[0], up[1], up[2]), vfov,
             aspect, aperture, focus, t0, t1);
  Context ctx = Context::create();
  cam.set(ctx);
  const char* names[7] = {"camera_origin", "camera_lower_left_corner", "camera_horizontal", "camera_vertical",
                          "camera_u", "camera_v", "camera_w"};
  for (int k = 0; k < 7; k++)
    for (int c = 0; c < 3; c++) out24[3 * k + c] = ctx[names[k]]->v[c];
  out24[21] = ctx["camera_lens_radius"]->v[0];
  out24[22] = ctx["time0"]->v[0];
  out24[23] = ctx["time1"]->v[0];
}
// the next n values of the scene functions' rnd() stream (a fresh process starts it at mt19937(0))
void refhost_rnd(int n, float* out) {
  for (int i = 0; i < n; i++) out[i] = rnd();
}
}
