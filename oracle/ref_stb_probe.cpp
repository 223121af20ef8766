// ref_stb_probe.cpp — TEST INFRASTRUCTURE.  Compiles the reference's vendored image reader
// (/root/reference/OptiX-Path-Tracer/lib/stb_image.h, included where it lies; nothing is copied)
// into oracle/_ref/libref_stb.so so that tests can compare the facade's own PNG/JPEG decoders
// (host_includes/image_decode.hpp) with what the reference's `stbi_load(file,&nx,&ny,&nn,0)`
// (host_includes/textures.hpp:120-121) returns for the same file.
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"

#include <cstring>

extern "C" int ref_stbi_load(const char* path, int* w, int* h, int* n, unsigned char* out, long long capacity) {
  unsigned char* d = stbi_load(path, w, h, n, 0);
  if (!d) return 1;
  long long bytes = (long long)*w * *h * *n;
  int rc = 0;
  if (out) {
    if (capacity < bytes) rc = 2;
    else memcpy(out, d, (size_t)bytes);
  }
  stbi_image_free(d);
  return rc;
}
