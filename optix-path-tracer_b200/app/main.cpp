// main.cpp — headless replacement for the reference's GUI application
// (OptiX-Path-Tracer/main.cpp): same App_State knobs as CLI flags, same
// Optix_Config -> render loop -> Save_PNG/Save_HDR flow, same console messages.
// The ImGui/GLFW window itself is out of scope (no GL on the box).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "../host_includes/image_save.hpp"

static int g_gpus = 1;  // --gpus N: frames split over devices 0..N-1, one NCCL reduce at read-back

static int Scene_Config(App_State& app) {
  if (g_gpus > 1) {
    std::vector<int> devs;
    for (int i = 0; i < g_gpus; i++) devs.push_back(i);
    app.context->setDevices(devs.begin(), devs.end());
  }
  app.context->setRayTypeCount(2);
  app.context->setMaxTraceDepth(5);
  app.context["samples"]->setInt(app.samples);
  app.context["russian"]->setInt(app.russian);
  app.context["maxDepth"]->setInt(app.depth);
  switch (app.scene) {
    case 0: InOneWeekend(app); break;
    case 1: MovingSpheres(app); break;
    case 2: Cornell(app); break;
    case 3: Final_Next_Week(app); break;
    case 4: Test_Scene(app); break;
    case 5: Cornell_Materials(app); break;
    case 6: Mesh_HDRI(app); break;
    case 7: Cornell_Sphere_Light(app); break;
    case 8: Textured_Mesh(app); break;
    case 9: Motion_Blur(app); break;
    case 10: Cornell_Smoke(app); break;
    default: throw "Selected scene is unknown";
  }
  app.accBuffer = createFrameBuffer(app.W, app.H, app.context);
  app.context["acc_buffer"]->set(app.accBuffer);
  app.displayBuffer = createDisplayBuffer(app.W, app.H, app.context);
  app.context["display_buffer"]->set(app.displayBuffer);
  auto t0 = std::chrono::system_clock::now();
  app.context->validate();
  app.context->launch(0, 0, 0);  // scene upload + GPU BVH build
  printf("BVH Building Time: %.2f\n", std::chrono::duration<float>(std::chrono::system_clock::now() - t0).count());
  return 0;
}

int main(int ac, char** av) {
  App_State app;
  int batch = 0;
  std::string checkpoint, resume, preview;  // progressive-render extras
  for (int i = 1; i < ac; i++) {
    auto arg = [&](const char* n) { return !strcmp(av[i], n) && i + 1 < ac; };
    if (arg("--width")) app.W = atoi(av[++i]);
    else if (arg("--height")) app.H = atoi(av[++i]);
    else if (arg("--samples")) app.samples = atoi(av[++i]);
    else if (arg("--scene")) app.scene = atoi(av[++i]);
    else if (arg("--model")) app.model = atoi(av[++i]);
    else if (arg("--depth")) app.depth = atoi(av[++i]);
    else if (arg("--russian")) app.russian = atoi(av[++i]) != 0;
    else if (arg("--file")) app.fileName = av[++i];
    else if (arg("--type")) app.fileType = atoi(av[++i]);
    else if (arg("--batch")) batch = atoi(av[++i]);
    else if (arg("--gpus")) g_gpus = atoi(av[++i]);
    else if (arg("--checkpoint")) checkpoint = av[++i];
    else if (arg("--resume")) resume = av[++i];
    else if (arg("--preview")) preview = av[++i];
    else {
      printf("usage: %s [--width W] [--height H] [--samples N] [--scene 0..10] [--model M] [--depth D]\n"
             "          [--russian 0|1] [--file name] [--type 0(PNG)|1(HDR)] [--batch frames-per-launch] [--gpus N]\n"
             "          [--checkpoint file: save the accumulation buffer after every batch]\n"
             "          [--resume file: continue a checkpointed render] [--preview file.png: display buffer after every batch]\n",
             av[0]);
      return 1;
    }
  }
  try {
    Scene_Config(app);
    auto t0 = std::chrono::system_clock::now();
    if (batch <= 0) batch = app.samples;  // default: the whole job in one wavefront run
    int first = 0;
    if (!resume.empty()) {  // header: W H samples-done, then W*H float4
      FILE* f = fopen(resume.c_str(), "rb");
      int hdr[3] = {0, 0, 0};
      std::vector<float> acc((size_t)app.W * app.H * 4);
      if (!f || fread(hdr, 4, 3, f) != 3 || hdr[0] != app.W || hdr[1] != app.H ||
          fread(acc.data(), 16, (size_t)app.W * app.H, f) != (size_t)app.W * app.H)
        throw "cannot resume: checkpoint missing or of another resolution";
      fclose(f);
      app.accBuffer->restore(acc.data(), app.W, app.H);
      first = hdr[2];
      printf("Resumed at sample %d\n", first);
    }
    for (app.currentSample = first; app.currentSample < app.samples; app.currentSample += batch) {
      int n = std::min(batch, app.samples - app.currentSample);
      if (batch == 1) {  // the reference's loop: one launch per sample (main.cpp:264-293)
        app.context["frame"]->setInt(app.currentSample);
        app.context->launch(0, app.W, app.H);
      } else {
        app.context->launchFrames(app.currentSample, n, app.W, app.H);
      }
      if (!preview.empty()) {  // the reference's showProgress copy of the display buffer (main.cpp:268-272)
        const unsigned char* px = (const unsigned char*)app.displayBuffer->map();
        std::vector<unsigned char> rgb((size_t)app.W * app.H * 3);
        for (size_t i = 0; i < (size_t)app.W * app.H; i++)
          for (int k = 0; k < 3; k++) rgb[3 * i + k] = px[4 * i + k];
        app.displayBuffer->unmap();
        ptio::write_png(preview.c_str(), app.W, app.H, rgb.data());
      }
      if (!checkpoint.empty()) {
        const float* acc = (const float*)app.accBuffer->map();
        FILE* f = fopen(checkpoint.c_str(), "wb");
        if (!f) throw "cannot write the checkpoint";
        int hdr[3] = {app.W, app.H, app.currentSample + n};
        fwrite(hdr, 4, 3, f);
        fwrite(acc, 16, (size_t)app.W * app.H, f);
        fclose(f);
        app.accBuffer->unmap();
      }
    }
    float renderTime = std::chrono::duration<float>(std::chrono::system_clock::now() - t0).count();
    int ok = app.fileType == 0 ? Save_PNG(app, app.accBuffer) : Save_HDR(app, app.accBuffer);
    printf("Render time: %.3f s (%.1f Msamples/s)\n", renderTime, double(app.W) * app.H * app.samples / renderTime / 1e6);
    printf("%s %s\n", ok ? "Saved" : "FAILED to save", app.fileName.c_str());
    return ok ? 0 : 2;
  } catch (const char* e) {
    fprintf(stderr, "error: %s\n", e);
  } catch (const std::string& e) {
    fprintf(stderr, "error: %s\n", e.c_str());
  }
  return 3;
}
