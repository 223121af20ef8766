// scenes.hpp — scene functions with the reference's names and signatures
// (host_includes/scenes.hpp:17,95,188,263,377) plus the two BASELINE configs the
// reference has no function for (C3 Cornell_Materials, C4/C5 Mesh_HDRI).
//
// The reference's own scenes.hpp also compiles UNCHANGED against these headers
// (tests/test_facade.py builds it from /root/reference and compares the flat scenes).
// rnd() draws inside one argument list are taken right-to-left, which is what g++
// (and MSVC x64, the reference's toolchain) do for the reference's expressions.
#pragma once
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>

#include "camera.hpp"
#include "hitables.hpp"
#include "mesh.hpp"
#include "pdfs.hpp"

namespace scene_detail {
// procedural assets are written to a directory of this PROCESS: with one process per GPU all ranks build the
// same scene at the same time, and a shared path had one rank reading a file another was still writing
inline std::string scratch_dir() {
  static const std::string dir = [] {
    std::string p = "/tmp/ptb200_" + std::to_string((long)getpid());
    mkdir(p.c_str(), 0700);
    return p;
  }();
  return dir;
}
inline Group beginScene(App_State& app, Light_Sampler& lights) {
  setRayGenerationProgram(app.context, lights);
  setExceptionProgram(app.context);
  Group group = app.context->createGroup();
  group->setAcceleration(app.context->createAcceleration("Trbvh"));
  return group;
}
inline void addRectLight(Light_Sampler& lights, App_State& app, float a0, float a1, float b0, float b1, float k, AXIS ax,
                         float emission) {
  Rectangle_PDF pdf(a0, a1, b0, b1, k, ax);
  lights.pdf.push_back(pdf.createPDF(app.context));
  lights.sample.push_back(pdf.createSample(app.context));
  lights.emissions.push_back(make_float3(emission));
}
inline void finishScene(App_State& app, Group& group, const float3& from, const float3& at, float fovy, float aperture,
                        float dist, std::chrono::system_clock::time_point t0) {
  app.context["world"]->set(group);
  Camera camera(from, at, make_float3(0.f, 1.f, 0.f), fovy, float(app.W) / float(app.H), aperture, dist, 0.0, 1.0);
  camera.set(app.context);
  float secs = std::chrono::duration<float>(std::chrono::system_clock::now() - t0).count();
  printf("Done assigning scene data, which took %.2f seconds.\n", secs);
}
// three rnd() draws, last component first
inline float3 rnd3_rtl(float sx, float ox) {
  float z = sx * (ox + rnd());
  float y = sx * (ox + rnd());
  float x = sx * (ox + rnd());
  return make_float3(x, y, z);
}
}  // namespace scene_detail

// Peter Shirley's "In One Weekend" final scene: 1 + 22x22 + 3 = 488 spheres, gradient sky
inline void InOneWeekend(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;  // no lights: the sky illuminates
  setMissProgram(app.context, GRADIENT, make_float3(1.f), make_float3(0.5f, 0.7f, 1.f));
  Group group = beginScene(app, lights);

  Hitable_List list;
  list.push(new Sphere(make_float3(0.f, -1000.f, -1.f), 1000.f, new Lambertian(new Constant_Texture(0.5f))));
  for (int a = -11; a < 11; a++)
    for (int b = -11; b < 11; b++) {
      float choose_mat = rnd();
      float cz = b + rnd();
      float cx = a + rnd();
      float3 center = make_float3(cx, 0.2f, cz);
      BRDF* mt;
      if (choose_mat < 0.8f) {
        float3 c = rnd3_rtl(1.f, 0.f);
        mt = new Lambertian(new Constant_Texture(c.x, c.y, c.z));
      } else if (choose_mat < 0.95f) {
        float3 c = rnd3_rtl(0.5f, 1.f);
        mt = new Metal(new Constant_Texture(c.x, c.y, c.z), 0.5f * rnd());
      } else {
        Texture* white = new Constant_Texture(1.f);
        float3 c = rnd3_rtl(1.f, 0.f);
        mt = new Dielectric(white, new Constant_Texture(c.x, c.y, c.z), 1.5, 0.f);
      }
      list.push(new Sphere(center, 0.2f, mt));
    }
  Texture* one = new Constant_Texture(1.f);
  list.push(new Sphere(make_float3(4.f, 1.f, 0.f), 1.f, new Dielectric(one, one, 1.5, 0.f)));
  list.push(new Sphere(make_float3(0.f, 1.f, 0.5f), 1.f, new Lambertian(new Constant_Texture(0.4f, 0.2f, 0.1f))));
  list.push(new Sphere(make_float3(-4.f, 1.f, 1.f), 1.f, new Metal(new Constant_Texture(0.7f, 0.6f, 0.5f), 0.f)));
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(13.f, 2.f, 3.f), make_float3(0.f), 20.f, 0.1f, 10.f, t0);
}

// "The Next Week" moving-spheres scene.  Motion blur is broken in the reference
// (hitables.hpp:130 FIXME) and it uses static Spheres too.  The earth image (JPEG decoding) is a
// "next" row: a constant stand-in keeps the scene usable.
inline void MovingSpheres(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  addRectLight(lights, app, 3.f, 5.f, 1.f, 3.f, -0.5f, Z_AXIS, 4.f);
  setMissProgram(app.context, CONSTANT);
  Group group = beginScene(app, lights);

  Hitable_List list;
  Texture* checker = new Checker_Texture(new Constant_Texture(0.2f, 0.3f, 0.1f), new Constant_Texture(0.9f, 0.9f, 0.9f));
  list.push(new Sphere(make_float3(0.f, -1000.f, -1.f), 1000.f, new Lambertian(checker)));
  for (int a = -11; a < 11; a++)
    for (int b = -11; b < 11; b++) {
      float choose_mat = rnd();
      float cz = b + rnd();
      float cx = a + rnd();
      float3 center = make_float3(cx, 0.2f, cz);
      (void)rnd();  // center2 (unused by the static spheres)
      float3 c = rnd3_rtl(0.5f, 1.f);
      Texture* tx = new Constant_Texture(c.x, c.y, c.z);
      BRDF* mt;
      if (choose_mat < (1.f / 3)) mt = new Lambertian(tx);
      else if (choose_mat < (2.f / 3)) mt = new Metal(tx, 0.5f * rnd());
      else mt = new Dielectric(tx, tx, 1.5, 0.f);
      list.push(new Sphere(center, 0.2f, mt));
    }
  list.push(new Sphere(make_float3(-4.f, 1.f, 2.f), 1.f, new Lambertian(new Constant_Texture(0.2f, 0.4f, 0.7f))));  // earth
  float3 g = rnd3_rtl(1.f, 0.f);
  list.push(new Sphere(make_float3(4.f, 1.f, 1.f), 1.f,
                       new Dielectric(new Constant_Texture(1.f), new Constant_Texture(g.x, g.y, g.z), 1.5, 0.f)));
  list.push(new Sphere(make_float3(0.f, 1.f, 1.5f), 1.f, new Metal(new Noise_Texture(4.f), 0.f)));  // 'rusty' metal
  list.push(new AARect(3.f, 5.f, 1.f, 3.f, -0.5f, false, Z_AXIS, new Diffuse_Light(new Constant_Texture(4.f))));
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(13.f, 2.f, 3.f), make_float3(0.f), 20.f, 0.1f, 10.f, t0);
}

// Cornell box: five walls, a 130x105 rect light (emission 7) and the tall box
inline void Cornell(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  addRectLight(lights, app, 213.f, 343.f, 227.f, 332.f, 554.f, Y_AXIS, 7.f);
  setMissProgram(app.context, CONSTANT);
  Group group = beginScene(app, lights);

  BRDF* red = new Lambertian(new Constant_Texture(0.65f, 0.05f, 0.05f));
  BRDF* white = new Lambertian(new Constant_Texture(0.73f));
  BRDF* green = new Lambertian(new Constant_Texture(0.12f, 0.45f, 0.15f));
  BRDF* light = new Diffuse_Light(new Constant_Texture(7.f));
  struct Wall { float a0, a1, b0, b1, k; bool flip; AXIS ax; BRDF* m; };
  const Wall walls[] = {{0.f, 555.f, 0.f, 555.f, 555.f, true, X_AXIS, red},
                        {0.f, 555.f, 0.f, 555.f, 0.f, false, X_AXIS, green},
                        {213.f, 343.f, 227.f, 332.f, 554.f, true, Y_AXIS, light},
                        {0.f, 555.f, 0.f, 555.f, 555.f, true, Y_AXIS, white},
                        {0.f, 555.f, 0.f, 555.f, 0.f, false, Y_AXIS, white},
                        {0.f, 555.f, 0.f, 555.f, 555.f, true, Z_AXIS, white}};
  Hitable_List list;
  for (const Wall& w : walls) list.push(new AARect(w.a0, w.a1, w.b0, w.b1, w.k, w.flip, w.ax, w.m));
  Box tall(make_float3(0.f), make_float3(165.f, 330.f, 165.f), green);
  tall.translate(make_float3(265.f, 0.f, 295.f));
  tall.rotate(15.f, Y_AXIS);
  list.push(&tall);
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(278.f, 278.f, -800.f), make_float3(278.f, 278.f, 0.f), 40.f, 0.f, 10.f, t0);
}

// "The Next Week" final scene (reference: host_includes/scenes.hpp:263-375) including its two
// Volumetric_Spheres; a constant texture stands in for the earth image (the asset is not in the tree)
inline void Final_Next_Week(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  addRectLight(lights, app, 113.f, 443.f, 127.f, 432.f, 554.f, Y_AXIS, 7.f);
  setMissProgram(app.context, CONSTANT);
  Group group = beginScene(app, lights);

  Hitable_List list;
  BRDF* ground = new Lambertian(new Constant_Texture(0.48f, 0.83f, 0.53f));
  for (int i = 0; i < 20; i++)
    for (int j = 0; j < 20; j++) {
      float w = 100.f, x0 = -1000 + i * w, z0 = -1000 + j * w;
      float y1 = 100 * (rnd() + 0.01f);
      list.push(new Box(make_float3(x0, 0.f, z0), make_float3(x0 + w, y1, z0 + w), ground));
    }
  list.push(new AARect(113.f, 443.f, 127.f, 432.f, 554.f, true, Y_AXIS, new Diffuse_Light(new Constant_Texture(7.f))));
  list.push(new Sphere(make_float3(400.f, 400.f, 200.f), 50.f, new Lambertian(new Constant_Texture(0.7f, 0.3f, 0.1f))));
  Texture* one = new Constant_Texture(1.f);
  BRDF* glass = new Dielectric(one, one, 1.5f);
  list.push(new Sphere(make_float3(260.f, 150.f, 45.f), 50.f, glass));
  list.push(new Sphere(make_float3(0.f, 150.f, 145.f), 50.f, new Metal(new Constant_Texture(0.8f, 0.8f, 0.9f), 10.f)));
  list.push(new Sphere(make_float3(360.f, 150.f, 45.f), 70.f, glass));
  list.push(new Volumetric_Sphere(make_float3(360.f, 150.f, 45.f), 70.f, 0.2f,
                                  new Isotropic(new Constant_Texture(0.2f, 0.4f, 0.9f))));  // blue fog inside the glass
  list.push(new Volumetric_Sphere(make_float3(0.f), 5000.f, 0.0001f, new Isotropic(one)));  // white fog everywhere
  list.push(new Sphere(make_float3(400.f, 200.f, 400.f), 100.f, new Lambertian(new Constant_Texture(0.2f, 0.4f, 0.7f))));
  list.push(new Sphere(make_float3(220.f, 280.f, 300.f), 80.f, new Lambertian(new Noise_Texture(0.1f))));  // Perlin sphere
  Hitable_List spheres;
  BRDF* white = new Lambertian(new Constant_Texture(0.73f));
  for (int j = 0; j < 1000; j++) {
    float z = 165 * rnd(), y = 165 * rnd(), x = 165 * rnd();
    spheres.push(new Sphere(make_float3(x, y, z), 10.f, white));
  }
  spheres.translate(make_float3(-100.f, 270.f, 395.f));
  spheres.rotate(15.f, Y_AXIS);
  spheres.addListTo(group, app.context);
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(478.f, 278.f, -600.f), make_float3(278.f, 278.f, 0.f), 40.f, 0.f, 10.f, t0);
}

// Motion blur: the bouncing spheres of "The Next Week" through Moving_Sphere
// (reference class: host_includes/hitables.hpp:131-173; its own MovingSpheres scene computes
// `center2` but then adds static Spheres, host_includes/scenes.hpp:124-142)
inline void Motion_Blur(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  setMissProgram(app.context, GRADIENT, make_float3(1.f), make_float3(0.5f, 0.7f, 1.f));
  Group group = beginScene(app, lights);
  Hitable_List list;
  Texture* checker = new Checker_Texture(new Constant_Texture(0.2f, 0.3f, 0.1f), new Constant_Texture(0.9f, 0.9f, 0.9f));
  list.push(new Sphere(make_float3(0.f, -1000.f, 0.f), 1000.f, new Lambertian(checker)));
  for (int a = -6; a < 6; a++)
    for (int b = -6; b < 6; b++) {
      float choose_mat = rnd();
      float cz = b + 0.9f * rnd();
      float cx = a + 0.9f * rnd();
      float3 center = make_float3(cx, 0.2f, cz);
      float3 c = rnd3_rtl(0.5f, 1.f);
      Texture* tx = new Constant_Texture(c.x, c.y, c.z);
      if (choose_mat < 0.7f) {
        float3 center2 = center + make_float3(0.f, 0.5f * rnd(), 0.f);
        list.push(new Moving_Sphere(center, center2, 0.2f, 0.f, 1.f, new Lambertian(tx)));
      } else if (choose_mat < 0.9f) {
        list.push(new Sphere(center, 0.2f, new Metal(tx, 0.5f * rnd())));
      } else {
        list.push(new Sphere(center, 0.2f, new Dielectric(tx, tx, 1.5, 0.f)));
      }
    }
  list.push(new Moving_Sphere(make_float3(-3.f, 1.f, 0.f), make_float3(-3.f, 1.f, 1.5f), 1.f, 0.f, 1.f,
                              new Lambertian(new Constant_Texture(0.4f, 0.2f, 0.1f))));
  list.push(new Sphere(make_float3(3.f, 1.f, 0.f), 1.f, new Metal(new Constant_Texture(0.7f, 0.6f, 0.5f), 0.f)));
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(13.f, 2.f, 3.f), make_float3(0.f), 20.f, 0.f, 10.f, t0);
}

// Volumes: Cornell box with two blocks of smoke (Volumetric_Box + Isotropic) and a fog ball
// (Volumetric_Sphere), the "cornell_smoke" arrangement of "The Next Week"
// (reference classes: host_includes/hitables.hpp:175-217,316-350, materials.hpp:125-141)
inline void Cornell_Smoke(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  addRectLight(lights, app, 113.f, 443.f, 127.f, 432.f, 554.f, Y_AXIS, 7.f);
  setMissProgram(app.context, CONSTANT);
  Group group = beginScene(app, lights);
  Hitable_List list;
  BRDF* red = new Lambertian(new Constant_Texture(0.65f, 0.05f, 0.05f));
  BRDF* white = new Lambertian(new Constant_Texture(0.73f));
  BRDF* green = new Lambertian(new Constant_Texture(0.12f, 0.45f, 0.15f));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 555.f, true, X_AXIS, red));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 0.f, false, X_AXIS, green));
  list.push(new AARect(113.f, 443.f, 127.f, 432.f, 554.f, true, Y_AXIS, new Diffuse_Light(new Constant_Texture(7.f))));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 555.f, true, Y_AXIS, white));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 0.f, false, Y_AXIS, white));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 555.f, true, Z_AXIS, white));
  Volumetric_Box smoke0(make_float3(0.f), make_float3(165.f, 330.f, 165.f), 0.01f, new Isotropic(new Constant_Texture(0.f)));
  smoke0.translate(make_float3(265.f, 0.f, 295.f));
  smoke0.rotate(15.f, Y_AXIS);
  list.push(&smoke0);
  Volumetric_Box smoke1(make_float3(0.f), make_float3(165.f, 165.f, 165.f), 0.01f, new Isotropic(new Constant_Texture(1.f)));
  smoke1.translate(make_float3(130.f, 0.f, 65.f));
  smoke1.rotate(-18.f, Y_AXIS);
  list.push(&smoke1);
  list.push(new Volumetric_Sphere(make_float3(400.f, 420.f, 150.f), 70.f, 0.02f,
                                  new Isotropic(new Constant_Texture(0.2f, 0.4f, 0.9f))));
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(278.f, 278.f, -800.f), make_float3(278.f, 278.f, 0.f), 40.f, 0.f, 10.f, t0);
}

// ---- procedural assets for the mesh scenes (written locally, then loaded through Mesh / HDR paths) ----
inline std::string write_heightfield_obj(const std::string& path, int n, bool with_normals = false) {
  FILE* f = fopen(path.c_str(), "w");
  if (!f) throw "cannot write " + path;
  const double size = 10.0, amp = 0.6;
  uint32_t h = 1234u;
  auto height = [&](double x, double z) { return amp * (sin(1.3 * x) * cos(0.9 * z) + 0.5 * sin(2.7 * x + 1.0) * sin(3.1 * z)); };
  for (int i = 0; i <= n; i++)
    for (int j = 0; j <= n; j++) {
      double x = -size / 2 + size * i / n, z = -size / 2 + size * j / n;
      h = h * 1664525u + 1013904223u;
      double jitter = 0.05 * (size / n) * (((h >> 8) & 0xffff) / 65535.0 - 0.5);
      fprintf(f, "v %.7g %.7g %.7g\n", x, height(x, z) + jitter, z);
    }
  if (with_normals)
    for (int i = 0; i <= n; i++)
      for (int j = 0; j <= n; j++) {
        double x = -size / 2 + size * i / n, z = -size / 2 + size * j / n, e = 1e-3;
        double dx = (height(x + e, z) - height(x - e, z)) / (2 * e), dz = (height(x, z + e) - height(x, z - e)) / (2 * e);
        double l = sqrt(dx * dx + 1 + dz * dz);
        fprintf(f, "vn %.6g %.6g %.6g\n", -dx / l, 1 / l, -dz / l);
      }
  auto vid = [&](int i, int j) { return i * (n + 1) + j + 1; };
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      int a = vid(i, j), b = vid(i, j + 1), c = vid(i + 1, j + 1), d = vid(i + 1, j);
      if (with_normals) fprintf(f, "f %d//%d %d//%d %d//%d\nf %d//%d %d//%d %d//%d\n", a, a, b, b, c, c, a, a, c, c, d, d);
      else fprintf(f, "f %d %d %d\nf %d %d %d\n", a, b, c, a, c, d);
    }
  fclose(f);
  return path;
}
inline std::string write_synthetic_hdri(const std::string& path, int w = 2048, int h = 1024, float peak = 0.5f) {
  std::vector<float> img((size_t)w * h * 3);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      double u = (x + 0.5) / w, v = (y + 0.5) / h;
      double sky = 0.15 + 0.25 * (1.0 - v);
      double sun = exp(-(pow((u - 0.3) / 0.03, 2) + pow((v - 0.25) / 0.04, 2)));
      float* p = &img[3 * ((size_t)y * w + x)];
      p[0] = fminf(peak, (float)(sky * 0.9 + peak * sun));
      p[1] = fminf(peak, (float)(sky * 1.0 + peak * sun * 0.95));
      p[2] = fminf(peak, (float)(sky * 1.2 + peak * sun * 0.8));
    }
  if (!ptio::write_radiance_hdr(path, w, h, img.data())) throw "cannot write " + path;
  return path;
}

// 3D-model test scene (scenes.hpp:377-526).  The reference's asset files are not
// distributed; app.model selects a procedural mesh instead: 0 -> ~1M triangles (708^2*2),
// 1 -> ~10M (2237^2*2), >= 2 -> model*model*2 triangles.  Gradient sky as in the reference.
inline void Test_Scene(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  setMissProgram(app.context, GRADIENT, make_float3(1.f), make_float3(0.5f, 0.7f, 1.f));
  Group group = beginScene(app, lights);
  int n = app.model == 0 ? 708 : (app.model == 1 ? 2237 : app.model);
  std::string obj = write_heightfield_obj(scene_detail::scratch_dir() + "/heightfield_" + std::to_string(n) + ".obj", n);
  BRDF* whiteMt = new Lambertian(new Constant_Texture(0.73f));
  Mesh model(obj, whiteMt, app.RTX);
  model.scale(make_float3(100.f));
  model.translate(make_float3(0.f, -300.f, 0.f));
  model.addTo(group, app.context);
  Hitable_List list;
  list.push(new AARect(-1000.f, 1000.f, -500.f, 500.f, -600.f, false, Y_AXIS, whiteMt));
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(0.f, 10.f, -800.f), make_float3(0.f), 100.f, 0.f, 0.8f, t0);
}

// BASELINE config C3: Cornell walls + glass sphere + Oren-Nayar / Torrance-Sparrow /
// Ashikhmin-Shirley spheres (on curved normals: those BSDFs are NaN where N || Y)
inline void Cornell_Materials(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  addRectLight(lights, app, 213.f, 343.f, 227.f, 332.f, 554.f, Y_AXIS, 7.f);
  setMissProgram(app.context, CONSTANT);
  Group group = beginScene(app, lights);
  BRDF* red = new Lambertian(new Constant_Texture(0.65f, 0.05f, 0.05f));
  BRDF* white = new Lambertian(new Constant_Texture(0.73f));
  BRDF* green = new Lambertian(new Constant_Texture(0.12f, 0.45f, 0.15f));
  Hitable_List list;
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 555.f, true, X_AXIS, red));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 0.f, false, X_AXIS, green));
  list.push(new AARect(213.f, 343.f, 227.f, 332.f, 554.f, true, Y_AXIS, new Diffuse_Light(new Constant_Texture(7.f))));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 555.f, true, Y_AXIS, white));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 0.f, false, Y_AXIS, white));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 555.f, true, Z_AXIS, white));
  Box tall(make_float3(0.f), make_float3(165.f, 330.f, 165.f), green);
  tall.translate(make_float3(265.f, 0.f, 295.f));
  tall.rotate(15.f, Y_AXIS);
  list.push(&tall);
  Texture* one = new Constant_Texture(1.f);
  Texture* alum = new Constant_Texture(0.8f, 0.85f, 0.88f);
  list.push(new Sphere(make_float3(130.f, 90.f, 120.f), 90.f, new Dielectric(one, one, 1.5f)));
  list.push(new Sphere(make_float3(420.f, 70.f, 150.f), 70.f, new Oren_Nayar(new Constant_Texture(0.8f, 0.6f, 0.2f), 0.5f)));
  list.push(new Sphere(make_float3(120.f, 320.f, 380.f), 80.f, new Torrance_Sparrow(alum, 0.2f, 0.2f)));
  list.push(new Sphere(make_float3(300.f, 420.f, 300.f), 60.f,
                       new Ashikhmin_Shirley(new Constant_Texture(0.3f, 0.4f, 0.8f), new Constant_Texture(0.9f), 100.f, 10.f)));
  list.push(new Sphere(make_float3(450.f, 250.f, 420.f), 55.f, new Metal(alum, 0.1f)));
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(278.f, 278.f, -800.f), make_float3(278.f, 278.f, 0.f), 40.f, 0.f, 10.f, t0);
}

// Cornell box lit by an emissive sphere sampled through Sphere_PDF plus the usual ceiling
// rectangle: two lights of different PDF kinds
inline void Cornell_Sphere_Light(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  Sphere_PDF ball(make_float3(278.f, 400.f, 278.f), 60.f);
  lights.pdf.push_back(ball.createPDF(app.context));
  lights.sample.push_back(ball.createSample(app.context));
  lights.emissions.push_back(make_float3(5.f));
  addRectLight(lights, app, 213.f, 343.f, 227.f, 332.f, 554.f, Y_AXIS, 7.f);
  setMissProgram(app.context, CONSTANT);
  Group group = beginScene(app, lights);
  BRDF* red = new Lambertian(new Constant_Texture(0.65f, 0.05f, 0.05f));
  BRDF* white = new Lambertian(new Constant_Texture(0.73f));
  BRDF* green = new Lambertian(new Constant_Texture(0.12f, 0.45f, 0.15f));
  BRDF* light = new Diffuse_Light(new Constant_Texture(7.f));
  Hitable_List list;
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 555.f, true, X_AXIS, red));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 0.f, false, X_AXIS, green));
  list.push(new AARect(213.f, 343.f, 227.f, 332.f, 554.f, true, Y_AXIS, light));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 555.f, true, Y_AXIS, white));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 0.f, false, Y_AXIS, white));
  list.push(new AARect(0.f, 555.f, 0.f, 555.f, 555.f, true, Z_AXIS, white));
  Box tall(make_float3(0.f), make_float3(165.f, 330.f, 165.f), green);
  tall.translate(make_float3(265.f, 0.f, 295.f));
  tall.rotate(15.f, Y_AXIS);
  list.push(&tall);
  list.push(new Sphere(make_float3(278.f, 400.f, 278.f), 60.f, light));
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(278.f, 278.f, -800.f), make_float3(278.f, 278.f, 0.f), 40.f, 0.f, 10.f, t0);
}

// MTL-driven mesh (reference path: mesh.hpp:74-102 + vector.cu + triangle.cu:119-146):
// a small terrain written as OBJ + MTL + PNG into /tmp, loaded WITHOUT a given material, so its
// faces pick one of three MTL materials — a map_Kd image, and two plain Ka colours — through a
// Vector_Texture, with vt texture coordinates and vn shading normals interpolated per hit.
inline std::string write_textured_obj(const std::string& dir, int n) {
  const std::string base = dir + "/ptb200_textured";
  {  // 32x32 RGB PNG: coloured checker with a gradient (decoded again by image_decode.hpp)
    const int S = 32;
    std::vector<unsigned char> px((size_t)S * S * 3);
    for (int y = 0; y < S; y++)
      for (int x = 0; x < S; x++) {
        const bool c = ((x / 4) + (y / 4)) & 1;
        unsigned char* o = &px[3 * ((size_t)y * S + x)];
        o[0] = (unsigned char)(c ? 230 : 40 + 4 * x);
        o[1] = (unsigned char)(c ? 200 - 3 * y : 60);
        o[2] = (unsigned char)(c ? 50 : 220 - 2 * x);
      }
    if (!ptio::write_png((base + ".png").c_str(), S, S, px.data())) throw "cannot write " + base + ".png";
  }
  {
    FILE* f = fopen((base + ".mtl").c_str(), "w");
    if (!f) throw "cannot write " + base + ".mtl";
    fprintf(f, "newmtl painted\nKa 1 1 1\nKd 1 1 1\nmap_Kd ptb200_textured.png\n\n");
    fprintf(f, "newmtl clay\nKa 0.75 0.42 0.3\nKd 0.75 0.42 0.3\n\nnewmtl moss\nKa 0.2 0.55 0.25\nKd 0.2 0.55 0.25\n");
    fclose(f);
  }
  FILE* f = fopen((base + ".obj").c_str(), "w");
  if (!f) throw "cannot write " + base + ".obj";
  fprintf(f, "mtllib ptb200_textured.mtl\n");
  const double size = 8.0;
  auto height = [&](double x, double z) { return 0.5 * sin(1.1 * x) * cos(0.8 * z); };
  for (int i = 0; i <= n; i++)
    for (int j = 0; j <= n; j++) {
      double x = -size / 2 + size * i / n, z = -size / 2 + size * j / n;
      fprintf(f, "v %.7g %.7g %.7g\n", x, height(x, z), z);
    }
  for (int i = 0; i <= n; i++)
    for (int j = 0; j <= n; j++) fprintf(f, "vt %.6g %.6g\n", 2.0 * i / n, 2.0 * j / n);  // wraps twice (REPEAT)
  for (int i = 0; i <= n; i++)
    for (int j = 0; j <= n; j++) {
      double x = -size / 2 + size * i / n, z = -size / 2 + size * j / n, e = 1e-3;
      double dx = (height(x + e, z) - height(x - e, z)) / (2 * e), dz = (height(x, z + e) - height(x, z - e)) / (2 * e);
      double l = sqrt(dx * dx + 1 + dz * dz);
      fprintf(f, "vn %.6g %.6g %.6g\n", -dx / l, 1 / l, -dz / l);
    }
  auto vid = [&](int i, int j) { return i * (n + 1) + j + 1; };
  const char* names[3] = {"painted", "clay", "moss"};
  int cur = -1;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      int m = (i < n / 2) ? 0 : ((j < n / 2) ? 1 : 2);
      if (m != cur) fprintf(f, "usemtl %s\n", names[cur = m]);
      int a = vid(i, j), b = vid(i, j + 1), c = vid(i + 1, j + 1), d = vid(i + 1, j);
      fprintf(f, "f %d/%d/%d %d/%d/%d %d/%d/%d\nf %d/%d/%d %d/%d/%d %d/%d/%d\n", a, a, a, b, b, b, c, c, c, a, a, a, c, c, c, d, d, d);
    }
  fclose(f);
  return "ptb200_textured.obj";
}
inline void Textured_Mesh(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  addRectLight(lights, app, -2.f, 2.f, -2.f, 2.f, 6.f, Y_AXIS, 6.f);
  setMissProgram(app.context, GRADIENT, make_float3(1.f), make_float3(0.5f, 0.7f, 1.f));
  Group group = beginScene(app, lights);
  const std::string dir = scratch_dir() + "/";
  Mesh model(write_textured_obj(scratch_dir(), app.model > 0 ? app.model : 24), dir, app.RTX);  // no given material: MTL path
  model.addTo(group, app.context);
  Hitable_List list;
  list.push(new AARect(-2.f, 2.f, -2.f, 2.f, 6.f, true, Y_AXIS, new Diffuse_Light(new Constant_Texture(6.f))));
  list.addElementsTo(group, app.context);
  finishScene(app, group, make_float3(0.f, 5.f, -9.f), make_float3(0.f), 40.f, 0.f, 10.f, t0);
}

// BASELINE configs C4/C5: procedural triangle mesh (OBJ written locally, loaded via Mesh) lit by a
// synthetic equirect HDRI (Radiance .hdr written locally, loaded via the HDR miss program)
inline void Mesh_HDRI(App_State& app) {
  using namespace scene_detail;
  auto t0 = std::chrono::system_clock::now();
  Light_Sampler lights;
  setMissProgram(app.context, HDR, write_synthetic_hdri(scene_detail::scratch_dir() + "/sky.hdr"), true);
  Group group = beginScene(app, lights);
  int n = app.model == 0 ? 708 : (app.model == 1 ? 2237 : app.model);
  std::string obj = write_heightfield_obj(scene_detail::scratch_dir() + "/heightfield_" + std::to_string(n) + ".obj", n);
  Mesh model(obj, new Lambertian(new Constant_Texture(0.7f)), app.RTX);
  model.addTo(group, app.context);
  std::remove(obj.c_str());  // loaded: a 1 M-triangle OBJ is 60 MB per process
  finishScene(app, group, make_float3(0.f, 4.5f, -9.f), make_float3(0.f), 40.f, 0.f, 10.f, t0);
}
