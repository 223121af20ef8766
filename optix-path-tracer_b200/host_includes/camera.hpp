// camera.hpp — thin-lens camera with a shutter interval.  Same constructor arguments and `set(Context&)` as the
// reference's class (surface: host_includes/camera.hpp:13-51), but the camera IS the ABI record here: the
// constructor fills a `pt_camera` (include/ptb200.h) and `set` hands that record to the scene — there are no
// named context variables to assign one by one on this side of the boundary.
//
// The frame is derived in the reference's operation order (pinned bit for bit against its own class by
// tests/test_oracle_host_golden.py): w = unit(from - at), u = unit(vup x w), v = w x u; the image plane sits at
// focus_dist and spans 2 * tan(vfov / 2) * focus_dist vertically.
#pragma once
#include "host_common.hpp"

class Camera {
 public:
  Camera() : rec_{} {}

  // vfov: full vertical field of view in degrees; aperture: lens DIAMETER; [t0, t1]: shutter interval
  Camera(const float3& lookfrom, const float3& lookat, const float3& vup, float vfov, float aspect, float aperture,
         float focus_dist, float t0, float t1)
      : rec_{} {
    const float half_h = tanf((vfov * PI_F / 180.f) / 2.f), half_w = aspect * half_h;
    const float3 back = unit_vector(lookfrom - lookat);
    const float3 right = unit_vector(cross(vup, back));
    const float3 up = cross(back, right);
    store(rec_.origin, lookfrom);
    store(rec_.u, right);
    store(rec_.v, up);
    store(rec_.w, back);
    store(rec_.lower_left_corner, lookfrom - half_w * focus_dist * right - half_h * focus_dist * up - focus_dist * back);
    store(rec_.horizontal, 2.f * half_w * focus_dist * right);
    store(rec_.vertical, 2.f * half_h * focus_dist * up);
    rec_.lens_radius = aperture / 2.f;
    rec_.time0 = t0;
    rec_.time1 = t1;
  }

  // Makes this the context's camera; a committed scene picks it up without a rebuild (pt_scene_set_camera).
  void set(Context& g_context) {
    g_context->camera = rec_;
    if (g_context->scene) pt_scene_set_camera(g_context->scene, &g_context->camera);
  }

  const pt_camera& record() const { return rec_; }

 private:
  static void store(float* dst, const float3& s) {
    dst[0] = s.x;
    dst[1] = s.y;
    dst[2] = s.z;
  }
  pt_camera rec_;
};
