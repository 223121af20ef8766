// image_save.hpp — image output (reference surface: host_includes/image_save.hpp:7-73).
#pragma once
#include "io_formats.hpp"
#include "scenes.hpp"

// average, sqrt gamma, clamp, x255.99 (image_save.hpp:13-29); PNG via ptio::write_png
inline int Save_PNG(App_State& app, Buffer& buffer) {
  std::vector<unsigned char> arr((size_t)app.W * app.H * 3);
  const float4* cols = (const float4*)buffer->map();
  for (int j = app.H - 1; j >= 0; j--)
    for (int i = 0; i < app.W; i++) {
      int index = app.W * j + i;
      float3 col = make_float3(cols[index].x, cols[index].y, cols[index].z);
      col = col / float(app.samples);
      col = make_float3(sqrtf(col.x), sqrtf(col.y), sqrtf(col.z));
      col = 255.99f * clamp(col, 0.f, 1.f);
      arr[3 * index + 0] = (unsigned char)(int)col.x;
      arr[3 * index + 1] = (unsigned char)(int)col.y;
      arr[3 * index + 2] = (unsigned char)(int)col.z;
    }
  buffer->unmap();
  app.fileName += ".png";
  return ptio::write_png(app.fileName.c_str(), app.W, app.H, arr.data());
}

// Reinhard tone mapping, Radiance .hdr (image_save.hpp:40-73; the reference's row loop
// bound `j = app.W - 1` is a bug for non-square images and is not reproduced)
inline int Save_HDR(App_State& app, Buffer& buffer) {
  std::vector<float> arr((size_t)app.W * app.H * 3);
  const float4* cols = (const float4*)buffer->map();
  for (int j = 0; j < app.H; j++)
    for (int i = 0; i < app.W; i++) {
      int index = app.W * j + i;
      float3 col = make_float3(cols[index].x, cols[index].y, cols[index].z) / float(app.samples);
      col = col / (make_float3(1.f) + col);
      arr[3 * index + 0] = col.x;
      arr[3 * index + 1] = col.y;
      arr[3 * index + 2] = col.z;
    }
  buffer->unmap();
  app.fileName += ".hdr";
  return ptio::write_radiance_hdr(app.fileName, app.W, app.H, arr.data()) ? 1 : 0;
}
