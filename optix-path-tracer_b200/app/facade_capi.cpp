// facade_capi.cpp — C entry points over the C++ host facade, so Python (tests, bench.py)
// can build scenes through the reference-compatible scene functions.
// SCENES_HEADER lets tests/test_facade.py compile the REFERENCE's own scenes.hpp
// (copied to a temp dir at test time, never into the repo) against these headers.
#ifndef SCENES_HEADER
#define SCENES_HEADER "../host_includes/scenes.hpp"
#endif
#include SCENES_HEADER

extern "C" {
// builds scene `scene_id` through the facade and returns an un-committed pt_scene (caller owns it)
int pth_build_scene(int scene_id, int W, int H, int model, pt_scene** out) {
  try {
    App_State app;
    app.W = W;
    app.H = H;
    app.scene = scene_id;
    app.model = model;
    switch (scene_id) {
      case 0: InOneWeekend(app); break;
      case 1: MovingSpheres(app); break;    // reference build: needs ../../../assets/other_textures/map.jpg
      case 2: Cornell(app); break;
      case 3: Final_Next_Week(app); break;  // reference build: same asset
#ifndef PTH_REFERENCE_SCENES
      case 4: Test_Scene(app); break;
      case 5: Cornell_Materials(app); break;
      case 6: Mesh_HDRI(app); break;
      case 7: Cornell_Sphere_Light(app); break;
      case 8: Textured_Mesh(app); break;
      case 9: Motion_Blur(app); break;
      case 10: Cornell_Smoke(app); break;
#endif
      default: return PT_ERR_INVALID;
    }
    *out = app.context->buildHostScene();
    return PT_OK;
  } catch (const char* e) {
    fprintf(stderr, "pth_build_scene: %s\n", e);
  } catch (const std::string& e) {
    fprintf(stderr, "pth_build_scene: %s\n", e.c_str());
  }
  return PT_ERR_INVALID;
}

// decodes an image file with the facade's own PNG/JPEG/PPM readers (host_includes/image_decode.hpp).
// Two-call pattern: pass out == NULL to get the size; returns 0 on success, *n_channels as stored.
int pth_decode_image(const char* path, int* w, int* h, int* n_channels, unsigned char* out, long long capacity) {
  ptio::DecodedImage im;
  if (!ptio::decode_image_file(path, im)) {
    fprintf(stderr, "pth_decode_image: %s\n", im.error.c_str());
    return PT_ERR_INVALID;
  }
  *w = im.w;
  *h = im.h;
  *n_channels = im.n;
  if (out) {
    if (capacity < (long long)im.px.size()) return PT_ERR_INVALID;
    memcpy(out, im.px.data(), im.px.size());
  }
  return PT_OK;
}

// loads an OBJ through Mesh(file, material) exactly as a scene function would (io_formats.hpp load_obj + mesh.hpp
// flattening); reports the triangle count and whether per-corner normals / uvs were kept
int pth_mesh_probe(const char* path, int* n_tris, int* has_normals, int* has_uvs) {
  try {
    App_State app;
    Mesh mesh(path, new Lambertian(new Constant_Texture(0.5f)), false);
    Group g = app.context->createGroup();
    mesh.addTo(g, app.context);
    long long tris = 0;
    bool nn = false, uu = false;
    for (auto& b : app.context->batches)
      if (b.kind == PT_PRIM_TRIANGLE) {
        tris += (long long)b.v.size() / 9;
        nn = nn || !b.n.empty();
        uu = uu || !b.uv.empty();
      }
    *n_tris = (int)tris;
    *has_normals = nn;
    *has_uvs = uu;
    return PT_OK;
  } catch (const char* e) {
    fprintf(stderr, "pth_mesh_probe: %s\n", e);
  } catch (const std::string& e) {
    fprintf(stderr, "pth_mesh_probe: %s\n", e.c_str());
  }
  return PT_ERR_INVALID;
}

// Image_Texture's loader (host_includes/textures.hpp: float4 texels, optional vertical flip)
int pth_load_ldr_image(const char* path, int flip_y, int* w, int* h, float* rgba, long long capacity_bytes) {
  std::vector<float> px;
  if (!ptio::load_ldr_image(path, *w, *h, px, flip_y != 0)) return PT_ERR_INVALID;
  if (rgba) {
    if (capacity_bytes < (long long)(px.size() * sizeof(float))) return PT_ERR_INVALID;
    memcpy(rgba, px.data(), px.size() * sizeof(float));
  }
  return PT_OK;
}
}
