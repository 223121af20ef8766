// sample_order.h — the order in which k_generate emits the camera samples of one pt_render call (host + device:
// tests/hostsim compiles the same function and checks that it is a bijection, tests/test_hostsim.py).
//
// Samples are independent (seed = tea<64>(W*y + x, frame), programs/raygen.cu:150; accumulation is a sum, :166), so
// the order is free - and it decides how coherent the 32 rays are that a traversal warp fetches together (camera
// rays are more than half of all rays of the mesh scenes):
//  * within a frame, pixels are enumerated in 8x4 tiles, and inside a tile in Morton order (2 consecutive positions =
//    2x1 pixels, 4 = 2x2, 8 = 4x2, 16 = 4x4, 32 = the tile) instead of 32-pixel strips of a scanline;
//  * frames are taken in groups of F = 2^kmax: ids run over (group, pixel, frame within the group) with the frame
//    fastest, so 32 consecutive ids are 32 / F neighbouring pixels x F samples of each (F >= 32: samples of ONE
//    pixel) - the rays of one pixel differ by the sub-pixel jitter only.  Frames that do not fill a group follow in
//    groups of the next smaller powers of two (nframes = 77, F = 32: 32 + 32 + 8 + 4 + 1).
// Every (pixel, frame) pair is generated exactly once and the frames of a pixel come in ascending order.
// C4 (1 M triangles): F = 1 / 4 / 32 / 256 -> 1896 / 1970 / 2075 / 2113 Msamples/s (profiles/r02_frame_groups.md).
#pragma once
#include "hd.h"

struct SampleOrder {
  uint32_t W, npix;  // pixels per row, pixels per frame (of the band being rendered)
  uint32_t nfr;      // frames of this render
  uint32_t top;      // frames in full groups: (nfr >> kmax) << kmax
  int kmax;          // log2 of the frames per group; npix << kmax <= 2^30
  int tiled;         // 8x4 tiles with Morton order inside (W % 8 == 0 && H % 4 == 0), else scanline order
};

PT_HD SampleOrder sample_order_make(int W, int H, uint32_t nfr, int kmax, int tiled) {
  SampleOrder so;
  so.W = (uint32_t)W;
  so.npix = (uint32_t)W * (uint32_t)H;
  so.nfr = nfr;
  so.kmax = kmax;
  so.top = (nfr >> kmax) << kmax;
  so.tiled = tiled;
  return so;
}

PT_HD int pt_log2_floor(uint32_t v) {  // v > 0
#ifdef __CUDA_ARCH__
  return 31 - __clz((int)v);
#else
  return 31 - __builtin_clz(v);
#endif
}

// sample id = fb * npix + off (fb < nfr, off < npix)  ->  pixel (x, y) and frame index relative to the first frame
PT_HD void sample_order_decode(const SampleOrder& so, uint32_t fb, uint32_t off, int& x, int& y, uint32_t& frame) {
  // the group of 2^j frames [f0, f0 + 2^j) that block fb belongs to: full groups first, then the binary digits of
  // the remaining frame count in descending order
  int j = so.kmax;
  uint32_t f0 = fb & ~((1u << so.kmax) - 1u);
  if (fb >= so.top) {
    const uint32_t nl = so.nfr - so.top, fl = fb - so.top;  // fl < nl: the highest differing bit is set in nl
    j = pt_log2_floor(nl ^ fl);
    f0 = so.top + ((nl >> (j + 1)) << (j + 1));
  }
  const uint32_t rem = (fb - f0) * so.npix + off;  // position inside the group: < npix << j
  const uint32_t po = rem >> j;                    // position of the pixel in the enumeration of one frame
  frame = f0 + (rem & ((1u << j) - 1u));
  if (so.tiled) {
    const uint32_t tile = po >> 5, in = po & 31u, tiles_x = so.W >> 3;
    const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
    x = (int)(tx * 8u + ((in & 1u) | ((in >> 1) & 2u) | ((in >> 2) & 4u)));
    y = (int)(ty * 4u + (((in >> 1) & 1u) | ((in >> 2) & 2u)));
  } else {
    x = (int)(po % so.W);
    y = (int)(po / so.W);
  }
}
