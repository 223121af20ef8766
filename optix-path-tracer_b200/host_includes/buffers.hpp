// buffers.hpp — output buffers (reference surface: host_includes/buffers.hpp:13-26).
#pragma once
#include "host_common.hpp"

// float4 accumulation buffer; lives on the device inside the renderer, map() reads it back
inline Buffer createFrameBuffer(int Nx, int Ny, Context& g_context) {
  Buffer b{std::make_shared<BufferObj>()};
  b->ctx = g_context.p.get();
  b->W = Nx;
  b->H = Ny;
  b->is_accum = true;
  return b;
}
// uchar4 preview buffer (make_Color, programs/raygen.cu:136-146), produced on map()
inline Buffer createDisplayBuffer(int Nx, int Ny, Context& g_context) {
  Buffer b = createFrameBuffer(Nx, Ny, g_context);
  b->is_accum = false;
  return b;
}
