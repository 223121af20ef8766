// io_formats.hpp — small self-contained file-format codecs the host facade
// needs: PNG writer (stored deflate), Radiance .hdr reader/writer, PNG/JPEG/PPM readers
// (image_decode.hpp),
// Wavefront OBJ/MTL reader.  They stand in for the third-party libraries the
// reference vendors (stb_image, stb_image_write, HDRloader, tinyobjloader);
// written from the format specifications, not derived from those libraries.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "image_decode.hpp"

namespace ptio {

// ---- PNG (RGB8, zlib "stored" blocks) ---------------------------------------------------------
inline uint32_t crc32_update(uint32_t c, const unsigned char* p, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t r = i;
      for (int k = 0; k < 8; k++) r = (r & 1) ? 0xEDB88320u ^ (r >> 1) : r >> 1;
      table[i] = r;
    }
    init = true;
  }
  c = ~c;
  for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xff] ^ (c >> 8);
  return ~c;
}
inline void put_be32(std::vector<unsigned char>& o, uint32_t v) {
  o.push_back(v >> 24); o.push_back(v >> 16); o.push_back(v >> 8); o.push_back(v);
}
inline void png_chunk(std::vector<unsigned char>& out, const char* tag, const std::vector<unsigned char>& data) {
  put_be32(out, (uint32_t)data.size());
  std::vector<unsigned char> body(tag, tag + 4);
  body.insert(body.end(), data.begin(), data.end());
  out.insert(out.end(), body.begin(), body.end());
  put_be32(out, crc32_update(0, body.data(), body.size()));
}
// rgb: W*H*3 bytes, row 0 = top.  Returns nonzero on success (like stbi_write_png).
inline int write_png(const char* path, int W, int H, const unsigned char* rgb) {
  std::vector<unsigned char> raw;
  raw.reserve((size_t)H * (W * 3 + 1));
  for (int y = 0; y < H; y++) {
    raw.push_back(0);  // filter: none
    raw.insert(raw.end(), rgb + (size_t)y * W * 3, rgb + (size_t)(y + 1) * W * 3);
  }
  std::vector<unsigned char> z = {0x78, 0x01};
  uint32_t a = 1, b = 0;
  for (unsigned char c : raw) {
    a = (a + c) % 65521;
    b = (b + a) % 65521;
  }
  size_t pos = 0;
  do {
    size_t n = std::min<size_t>(65535, raw.size() - pos);
    z.push_back(pos + n == raw.size() ? 1 : 0);
    z.push_back(n & 0xff); z.push_back(n >> 8);
    z.push_back(~n & 0xff); z.push_back((~n >> 8) & 0xff);
    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
    pos += n;
  } while (pos < raw.size());
  put_be32(z, (b << 16) | a);
  std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  std::vector<unsigned char> ihdr;
  put_be32(ihdr, W);
  put_be32(ihdr, H);
  ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});
  png_chunk(out, "IHDR", ihdr);
  png_chunk(out, "IDAT", z);
  png_chunk(out, "IEND", {});
  FILE* f = fopen(path, "wb");
  if (!f) return 0;
  size_t w = fwrite(out.data(), 1, out.size(), f);
  fclose(f);
  return w == out.size();
}

// ---- Radiance RGBE (.hdr) ----------------------------------------------------------------------------
// Decoding follows the reference's loader semantics (lib/HDRloader.cpp:99-116):
// component = mantissa/256 * 2^(E-128); rows are stored top to bottom and kept in file order.
inline bool load_radiance_hdr(const std::string& path, int& W, int& H, std::vector<float>& rgba) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char line[512];
  if (!fgets(line, sizeof(line), f) || strncmp(line, "#?", 2) != 0) {
    fclose(f);
    return false;
  }
  while (fgets(line, sizeof(line), f) && line[0] != '\n') {}
  if (!fgets(line, sizeof(line), f)) {
    fclose(f);
    return false;
  }
  long h = 0, w = 0;
  if (sscanf(line, "-Y %ld +X %ld", &h, &w) != 2 || w <= 0 || h <= 0) {
    fclose(f);
    return false;
  }
  W = (int)w;
  H = (int)h;
  rgba.assign((size_t)W * H * 4, 0.f);
  std::vector<unsigned char> scan((size_t)W * 4);
  for (int y = 0; y < H; y++) {
    int c0 = fgetc(f), c1 = fgetc(f), c2 = fgetc(f), c3 = fgetc(f);
    if (c3 == EOF) break;
    if (c0 == 2 && c1 == 2 && !(c2 & 128) && ((c2 << 8) | c3) == W && W >= 8 && W < 32768) {
      for (int comp = 0; comp < 4; comp++) {  // adaptive RLE, one channel at a time
        int x = 0;
        while (x < W) {
          int code = fgetc(f);
          if (code == EOF) { fclose(f); return false; }
          if (code > 128) {
            int val = fgetc(f);
            for (code &= 127; code-- && x < W;) scan[4 * x++ + comp] = (unsigned char)val;
          } else {
            while (code-- && x < W) scan[4 * x++ + comp] = (unsigned char)fgetc(f);
          }
        }
      }
    } else {  // flat RGBE pixels
      scan[0] = c0; scan[1] = c1; scan[2] = c2; scan[3] = c3;
      if (fread(scan.data() + 4, 1, (size_t)(W - 1) * 4, f) != (size_t)(W - 1) * 4) { fclose(f); return false; }
    }
    for (int x = 0; x < W; x++) {
      int expo = scan[4 * x + 3] - 128;
      float d = powf(2.f, (float)expo);
      float* o = &rgba[4 * ((size_t)y * W + x)];
      o[0] = scan[4 * x + 0] / 256.0f * d;
      o[1] = scan[4 * x + 1] / 256.0f * d;
      o[2] = scan[4 * x + 2] / 256.0f * d;
      o[3] = 0.f;  // host_includes/textures.hpp:209
    }
  }
  fclose(f);
  return true;
}
// rgb: W*H*3 floats, row 0 = top; flat (non-RLE) RGBE
inline bool write_radiance_hdr(const std::string& path, int W, int H, const float* rgb) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n", H, W);
  std::vector<unsigned char> row((size_t)W * 4);
  for (int y = 0; y < H; y++) {
    for (int x = 0; x < W; x++) {
      const float* p = rgb + 3 * ((size_t)y * W + x);
      float m = fmaxf(p[0], fmaxf(p[1], p[2]));
      unsigned char* o = &row[4 * x];
      if (m < 1e-32f) {
        o[0] = o[1] = o[2] = o[3] = 0;
      } else {
        int e;
        float s = frexpf(m, &e) * 256.0f / m;
        o[0] = (unsigned char)(p[0] * s);
        o[1] = (unsigned char)(p[1] * s);
        o[2] = (unsigned char)(p[2] * s);
        o[3] = (unsigned char)(e + 128);
      }
    }
    fwrite(row.data(), 1, row.size(), f);
  }
  fclose(f);
  return true;
}

// ---- LDR images: PNG, baseline JPEG, binary PPM (image_decode.hpp) --------------------------------------
// Produces the float4 texels the reference's normalised-float sampler would return:
// rgb/255, alpha = 1/255 for 3-channel images (textures.hpp:152-156 `(unsigned char)1.f`).
inline bool load_ldr_image(const std::string& path, int& W, int& H, std::vector<float>& rgba, bool flip_y) {
  DecodedImage im;
  if (!decode_image_file(path, im)) {
    fprintf(stderr, "load_ldr_image(%s): %s\n", path.c_str(), im.error.c_str());
    return false;
  }
  W = im.w;
  H = im.h;
  rgba.resize((size_t)W * H * 4);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const unsigned char* s = &im.px[(size_t)im.n * ((size_t)(flip_y ? H - 1 - y : y) * W + x)];
      float* o = &rgba[4 * ((size_t)y * W + x)];
      // textures.hpp:143-156: RGB copied, alpha kept for 4-channel files, else (unsigned char)1.f.
      // Grey files are replicated (the reference indexes past the pixel there).
      const bool grey = im.n < 3;
      o[0] = s[0] / 255.f;
      o[1] = s[grey ? 0 : 1] / 255.f;
      o[2] = s[grey ? 0 : 2] / 255.f;
      o[3] = im.n == 4 ? s[3] / 255.f : 1.f / 255.f;
    }
  return true;
}

// ---- Wavefront OBJ / MTL -----------------------------------------------------------------------------------
struct ObjMaterial {
  std::string name, diffuse_texname;
  float ambient[3] = {0, 0, 0}, diffuse[3] = {0, 0, 0};
};
struct ObjMesh {  // triangulated, per-corner indices (-1 = absent), like tinyobj's index_t stream
  std::vector<float> vertices, normals, texcoords;
  std::vector<int> vi, ni, ti;   // 3 per triangle
  std::vector<int> material_ids; // per triangle, -1 = none
  std::vector<ObjMaterial> materials;
};
inline bool load_mtl(const std::string& path, std::vector<ObjMaterial>& out) {
  std::ifstream in(path);
  if (!in) return false;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string k;
    if (!(ss >> k)) continue;
    if (k == "newmtl") { out.emplace_back(); ss >> out.back().name; }
    else if (out.empty()) continue;
    else if (k == "Ka") ss >> out.back().ambient[0] >> out.back().ambient[1] >> out.back().ambient[2];
    else if (k == "Kd") ss >> out.back().diffuse[0] >> out.back().diffuse[1] >> out.back().diffuse[2];
    else if (k == "map_Kd") ss >> out.back().diffuse_texname;
  }
  return true;
}
inline bool load_obj(const std::string& path, const std::string& mtl_dir, ObjMesh& m, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  std::map<std::string, int> matmap;
  int cur = -1;
  std::vector<char> buf(1 << 16);
  std::vector<int> fv, fn, ft;
  while (fgets(buf.data(), (int)buf.size(), f)) {
    char* s = buf.data();
    while (*s == ' ' || *s == '\t') s++;
    if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
      float x = 0, y = 0, z = 0;
      sscanf(s + 2, "%f %f %f", &x, &y, &z);
      m.vertices.insert(m.vertices.end(), {x, y, z});
    } else if (s[0] == 'v' && s[1] == 'n') {
      float x = 0, y = 0, z = 0;
      sscanf(s + 3, "%f %f %f", &x, &y, &z);
      m.normals.insert(m.normals.end(), {x, y, z});
    } else if (s[0] == 'v' && s[1] == 't') {
      float x = 0, y = 0;
      sscanf(s + 3, "%f %f", &x, &y);
      m.texcoords.insert(m.texcoords.end(), {x, y});
    } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
      fv.clear(); fn.clear(); ft.clear();
      char* p = s + 2;
      while (*p) {
        while (*p == ' ' || *p == '\t') p++;
        if (*p == '\n' || *p == '\r' || !*p) break;
        int v = 0, t = 0, n = 0;
        char* tok = p;
        v = (int)strtol(p, &p, 10);
        if (p == tok) break;  // not a number (e.g. a trailing "# comment"): the face ends here
        if (*p == '/') {
          p++;
          if (*p != '/') t = (int)strtol(p, &p, 10);
          if (*p == '/') { p++; n = (int)strtol(p, &p, 10); }
        }
        int nv = (int)m.vertices.size() / 3, nt = (int)m.texcoords.size() / 2, nn = (int)m.normals.size() / 3;
        fv.push_back(v > 0 ? v - 1 : (v < 0 ? nv + v : -1));
        ft.push_back(t > 0 ? t - 1 : (t < 0 ? nt + t : -1));
        fn.push_back(n > 0 ? n - 1 : (n < 0 ? nn + n : -1));
      }
      for (size_t k = 1; k + 1 < fv.size(); k++) {  // fan triangulation
        m.vi.insert(m.vi.end(), {fv[0], fv[k], fv[k + 1]});
        m.ni.insert(m.ni.end(), {fn[0], fn[k], fn[k + 1]});
        m.ti.insert(m.ti.end(), {ft[0], ft[k], ft[k + 1]});
        m.material_ids.push_back(cur);
      }
    } else if (!strncmp(s, "usemtl", 6)) {
      char name[256] = "";
      sscanf(s + 6, "%255s", name);
      auto it = matmap.find(name);
      cur = it == matmap.end() ? -1 : it->second;
    } else if (!strncmp(s, "mtllib", 6)) {
      char name[256] = "";
      sscanf(s + 6, "%255s", name);
      size_t before = m.materials.size();
      load_mtl(mtl_dir + name, m.materials);
      for (size_t i = before; i < m.materials.size(); i++) matmap[m.materials[i].name] = (int)i;
    }
  }
  fclose(f);
  for (int v : m.vi)
    if (v < 0 || v >= (int)m.vertices.size() / 3) { err = "face references a missing vertex"; return false; }
  return true;
}

}  // namespace ptio
