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
      case 2: Cornell(app); break;
#ifndef PTH_REFERENCE_SCENES
      case 1: MovingSpheres(app); break;
      case 3: Final_Next_Week(app); break;
      case 4: Test_Scene(app); break;
      case 5: Cornell_Materials(app); break;
      case 6: Mesh_HDRI(app); break;
      case 7: Cornell_Sphere_Light(app); break;
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
}
