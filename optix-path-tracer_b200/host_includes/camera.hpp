// camera.hpp — thin-lens camera (reference surface: host_includes/camera.hpp:13-51).
#pragma once
#include "host_common.hpp"

class Camera {
 public:
  Camera() {}
  Camera(const float3& lookfrom, const float3& lookat, const float3& vup, float vfov, float aspect, float aperture,
         float focus_dist, float t0, float t1) {
    lens_radius = aperture / 2.f;
    time0 = t0;
    time1 = t1;
    float theta = vfov * PI_F / 180.f;
    float half_height = tanf(theta / 2.f);
    float half_width = aspect * half_height;
    origin = lookfrom;
    w = unit_vector(lookfrom - lookat);
    u = unit_vector(cross(vup, w));
    v = cross(w, u);
    lower_left_corner = origin - half_width * focus_dist * u - half_height * focus_dist * v - focus_dist * w;
    horizontal = 2.f * half_width * focus_dist * u;
    vertical = 2.f * half_height * focus_dist * v;
  }
  void set(Context& g_context) {
    pt_camera& c = g_context->camera;
    auto put = [](float* d, const float3& s) { d[0] = s.x; d[1] = s.y; d[2] = s.z; };
    put(c.origin, origin);
    put(c.lower_left_corner, lower_left_corner);
    put(c.horizontal, horizontal);
    put(c.vertical, vertical);
    put(c.u, u);
    put(c.v, v);
    put(c.w, w);
    c.lens_radius = lens_radius;
    c.time0 = time0;
    c.time1 = time1;
    if (g_context->scene) pt_scene_set_camera(g_context->scene, &c);
  }

 private:
  float3 origin, lower_left_corner, horizontal, vertical, u, v, w;
  float time0, time1, lens_radius;
};
