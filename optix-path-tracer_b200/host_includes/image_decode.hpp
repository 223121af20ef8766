// image_decode.hpp — PNG and JPEG decoders for Image_Texture / the IMG miss program / MTL texture
// maps.  The reference reads these through its vendored stb_image (`stbi_load(file,&nx,&ny,&nn,0)`,
// host_includes/textures.hpp:120-121); this file is written from the format specifications
// (RFC 1950/1951, the PNG specification, ITU-T T.81) and returns the same thing: 8-bit samples,
// row 0 = top, `nn` interleaved channels as stored in the file (1 = grey, 2 = grey+alpha, 3 = RGB,
// 4 = RGBA).  PNG is lossless, so any conforming decoder is bit-identical; for JPEG the standard
// leaves IDCT rounding and chroma upsampling to the decoder — this one evaluates the exact IDCT in
// floating point and upsamples with the centred triangle filter both stb_image and libjpeg use, which
// keeps it within a few grey levels of either (tests/test_image_decode.py states the tolerance).
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace ptio {

struct DecodedImage {
  int w = 0, h = 0, n = 0;  // n = channels per pixel
  std::vector<unsigned char> px;
  std::string error;
};

// ---------------------------------------------------------------------------------------------
// inflate (RFC 1951) inside a zlib wrapper (RFC 1950)
// ---------------------------------------------------------------------------------------------
namespace inflate_detail {

struct BitReader {
  const unsigned char* p;
  size_t n, pos = 0;
  uint64_t acc = 0;
  int cnt = 0;
  bool eof = false;  // bits were consumed past the end of the input (they read as zero)
  void fill() {
    while (cnt <= 56 && pos < n) {
      acc |= (uint64_t)p[pos++] << cnt;
      cnt += 8;
    }
  }
  uint32_t peek(int k) {
    if (cnt < k) fill();
    return (uint32_t)(acc & ((1ull << k) - 1));
  }
  void drop(int k) {
    acc >>= k;
    cnt -= k;
    if (cnt < 0) {  // underflow: a truncated stream must not decode zeros for ever
      cnt = 0;
      eof = true;
    }
  }
  uint32_t get(int k) {
    uint32_t v = peek(k);
    drop(k);
    return v;
  }
  void align_byte() { drop(cnt & 7); }
};

// full-width lookup: entry = (length << 16) | symbol for every 15-bit window (LSB-first codes)
struct Huff {
  std::vector<uint32_t> tab;
  bool build(const unsigned char* lens, int nsym) {
    tab.assign(1u << 15, 0);
    int count[16] = {0};
    for (int i = 0; i < nsym; i++) count[lens[i]]++;
    count[0] = 0;
    uint32_t next[16], code = 0;
    for (int l = 1; l < 16; l++) {
      code = (code + count[l - 1]) << 1;
      next[l] = code;
    }
    for (int s = 0; s < nsym; s++) {
      int l = lens[s];
      if (!l) continue;
      uint32_t c = next[l]++;
      if (c >> l) return false;  // over-subscribed
      uint32_t rev = 0;
      for (int b = 0; b < l; b++) rev |= ((c >> b) & 1u) << (l - 1 - b);
      for (uint32_t hi = rev; hi < (1u << 15); hi += 1u << l) tab[hi] = ((uint32_t)l << 16) | (uint32_t)s;
    }
    return true;
  }
  int decode(BitReader& br) const {
    uint32_t e = tab[br.peek(15)];
    if (!e) return -1;
    br.drop((int)(e >> 16));
    return (int)(e & 0xffff);
  }
};

}  // namespace inflate_detail

inline bool inflate_zlib(const unsigned char* in, size_t n, std::vector<unsigned char>& out, std::string& err) {
  using namespace inflate_detail;
  if (n < 6 || (in[0] & 0x0f) != 8 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) {
    err = "bad zlib header";
    return false;
  }
  static const int len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const int len_bits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const int dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const int dist_bits[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  static const int cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  BitReader br{in + 2, n - 2};
  Huff lit, dist;
  for (bool last = false; !last;) {
    last = br.get(1) != 0;
    int type = (int)br.get(2);
    if (br.eof) { err = "truncated deflate stream"; return false; }
    if (type == 0) {
      br.align_byte();
      uint32_t len = br.get(16), nlen = br.get(16);
      if ((len ^ nlen) != 0xffffu) { err = "stored block length mismatch"; return false; }
      for (uint32_t i = 0; i < len; i++) {
        if (br.cnt < 8 && br.pos >= br.n) { err = "truncated stored block"; return false; }
        out.push_back((unsigned char)br.get(8));
      }
      continue;
    }
    if (type == 3) { err = "reserved block type"; return false; }
    unsigned char lens[320];
    int nlit = 288, ndist = 30;
    if (type == 1) {
      for (int i = 0; i < 144; i++) lens[i] = 8;
      for (int i = 144; i < 256; i++) lens[i] = 9;
      for (int i = 256; i < 280; i++) lens[i] = 7;
      for (int i = 280; i < 288; i++) lens[i] = 8;
      for (int i = 0; i < 30; i++) lens[288 + i] = 5;
    } else {
      nlit = (int)br.get(5) + 257;
      ndist = (int)br.get(5) + 1;
      int ncl = (int)br.get(4) + 4;
      unsigned char cl[19] = {0};
      for (int i = 0; i < ncl; i++) cl[cl_order[i]] = (unsigned char)br.get(3);
      Huff clh;
      if (!clh.build(cl, 19)) { err = "bad code-length code"; return false; }
      int i = 0;
      while (i < nlit + ndist) {
        int s = clh.decode(br);
        if (s < 0) { err = "bad code-length symbol"; return false; }
        if (br.eof) { err = "truncated deflate stream"; return false; }
        if (s < 16) { lens[i++] = (unsigned char)s; continue; }
        int rep, val = 0;
        if (s == 16) {
          if (i == 0) { err = "repeat without previous length"; return false; }
          val = lens[i - 1];
          rep = 3 + (int)br.get(2);
        } else if (s == 17) rep = 3 + (int)br.get(3);
        else rep = 11 + (int)br.get(7);
        if (i + rep > nlit + ndist) { err = "code lengths overflow"; return false; }
        while (rep--) lens[i++] = (unsigned char)val;
      }
      memmove(lens + 288, lens + nlit, (size_t)ndist);
    }
    if (!lit.build(lens, nlit) || !dist.build(lens + 288, ndist)) { err = "bad Huffman table"; return false; }
    for (;;) {
      int s = lit.decode(br);
      if (s < 0) { err = "bad literal/length code"; return false; }
      if (br.eof) { err = "truncated deflate stream"; return false; }
      if (s < 256) { out.push_back((unsigned char)s); continue; }
      if (s == 256) break;
      s -= 257;
      if (s >= 29) { err = "bad length symbol"; return false; }
      int len = len_base[s] + (int)br.get(len_bits[s]);
      int d = dist.decode(br);
      if (d < 0 || d >= 30) { err = "bad distance code"; return false; }
      size_t back = (size_t)dist_base[d] + br.get(dist_bits[d]);
      if (back > out.size()) { err = "distance beyond start of output"; return false; }
      size_t from = out.size() - back;
      for (int k = 0; k < len; k++) out.push_back(out[from + k]);
    }
    if (br.cnt == 0 && br.pos >= br.n && !last) { err = "truncated deflate stream"; return false; }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// PNG: all colour types, 1/2/4/8/16-bit, palette + tRNS, Adam7 interlace.  16-bit samples keep
// their high byte (stbi_load's 8-bit interface does the same).  Like stb_image, a tRNS chunk on
// a grey or RGB image adds an alpha channel; palette images come out RGB (or RGBA with tRNS).
// ---------------------------------------------------------------------------------------------
namespace png_detail {
inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
  int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
// un-filters one pass of w x h pixels stored at `src`; writes w*h*chan samples (8-bit, expanded)
// to `dst` with the given pixel stride/offset pattern.  Returns bytes consumed or 0 on error.
inline size_t unfilter_pass(const unsigned char* src, size_t avail, int w, int h, int chan, int depth,
                            std::vector<unsigned char>& rows) {
  const size_t bpp_bits = (size_t)chan * depth;
  const size_t stride = ((size_t)w * bpp_bits + 7) / 8;
  const size_t bpp = bpp_bits < 8 ? 1 : bpp_bits / 8;
  if (avail < (stride + 1) * (size_t)h) return 0;
  rows.assign(stride * (size_t)h, 0);
  std::vector<unsigned char> zero(stride, 0);
  for (int y = 0; y < h; y++) {
    const unsigned char* in = src + (stride + 1) * (size_t)y;
    const int ft = in[0];
    in++;
    unsigned char* cur = rows.data() + stride * (size_t)y;
    const unsigned char* up = y ? cur - stride : zero.data();
    if (ft > 4) return 0;
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
      int v = in[i];
      switch (ft) {
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: break;
      }
      cur[i] = (unsigned char)v;
    }
  }
  return (stride + 1) * (size_t)h;
}
}  // namespace png_detail

inline bool decode_png(const unsigned char* data, size_t n, DecodedImage& im) {
  using namespace png_detail;
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (n < 8 || memcmp(data, sig, 8) != 0) { im.error = "not a PNG"; return false; }
  size_t pos = 8;
  int W = 0, H = 0, depth = 0, ctype = -1, interlace = 0;
  std::vector<unsigned char> idat, plte, trns;
  bool seen_end = false;
  while (pos + 12 <= n && !seen_end) {
    uint32_t len = be32(data + pos);
    const unsigned char* tag = data + pos + 4;
    const unsigned char* body = data + pos + 8;
    if (pos + 12 + (size_t)len > n) { im.error = "truncated PNG chunk"; return false; }
    if (!memcmp(tag, "IHDR", 4)) {
      if (len != 13) { im.error = "bad IHDR"; return false; }
      W = (int)be32(body);
      H = (int)be32(body + 4);
      depth = body[8];
      ctype = body[9];
      interlace = body[12];
      if (body[10] != 0 || body[11] != 0 || interlace > 1 || W <= 0 || H <= 0) { im.error = "unsupported PNG header"; return false; }
    } else if (!memcmp(tag, "PLTE", 4)) plte.assign(body, body + len);
    else if (!memcmp(tag, "tRNS", 4)) trns.assign(body, body + len);
    else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
    else if (!memcmp(tag, "IEND", 4)) seen_end = true;
    pos += 12 + (size_t)len;
  }
  int chan;
  switch (ctype) {
    case 0: chan = 1; break;
    case 2: chan = 3; break;
    case 3: chan = 1; break;
    case 4: chan = 2; break;
    case 6: chan = 4; break;
    default: im.error = "bad PNG colour type"; return false;
  }
  const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                        (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                        ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
  if (!depth_ok) { im.error = "bad PNG bit depth"; return false; }
  if (ctype == 3 && plte.size() < 3) { im.error = "palette PNG without PLTE"; return false; }
  std::vector<unsigned char> raw;
  if (!inflate_zlib(idat.data(), idat.size(), raw, im.error)) return false;

  // output channel layout
  const bool add_alpha = !trns.empty() && (ctype == 0 || ctype == 2 || ctype == 3);
  const int out_n = ctype == 3 ? (add_alpha ? 4 : 3) : chan + (add_alpha ? 1 : 0);
  im.w = W;
  im.h = H;
  im.n = out_n;
  im.px.assign((size_t)W * H * out_n, 0);
  int tr[3] = {-1, -1, -1};  // transparent colour key (at file bit depth)
  if (add_alpha && ctype == 0 && trns.size() >= 2) tr[0] = (trns[0] << 8) | trns[1];
  if (add_alpha && ctype == 2 && trns.size() >= 6)
    for (int k = 0; k < 3; k++) tr[k] = (trns[2 * k] << 8) | trns[2 * k + 1];

  auto put_pixel = [&](int x, int y, const int* s) {  // s: chan samples at file depth
    unsigned char* o = &im.px[((size_t)y * W + x) * out_n];
    if (ctype == 3) {
      size_t idx = (size_t)s[0];
      if (idx * 3 + 2 < plte.size()) { o[0] = plte[idx * 3]; o[1] = plte[idx * 3 + 1]; o[2] = plte[idx * 3 + 2]; }
      if (add_alpha) o[3] = idx < trns.size() ? trns[idx] : 255;
      return;
    }
    for (int k = 0; k < chan; k++) {
      int v = s[k];
      if (depth == 16) v >>= 8;
      else if (depth < 8) v = v * 255 / ((1 << depth) - 1);
      o[k] = (unsigned char)v;
    }
    if (add_alpha) {
      bool match = true;
      for (int k = 0; k < chan; k++) match &= s[k] == tr[k];
      o[chan] = match ? 0 : 255;
    }
  };

  static const int ax0[7] = {0, 4, 0, 2, 0, 1, 0}, ay0[7] = {0, 0, 4, 0, 2, 0, 1};
  static const int adx[7] = {8, 8, 4, 4, 2, 2, 1}, ady[7] = {8, 8, 8, 4, 4, 2, 2};
  size_t off = 0;
  std::vector<unsigned char> rows;
  const int npass = interlace ? 7 : 1;
  for (int pass = 0; pass < npass; pass++) {
    const int x0 = interlace ? ax0[pass] : 0, y0 = interlace ? ay0[pass] : 0;
    const int dx = interlace ? adx[pass] : 1, dy = interlace ? ady[pass] : 1;
    const int pw = (W - x0 + dx - 1) / dx, ph = (H - y0 + dy - 1) / dy;
    if (pw <= 0 || ph <= 0) continue;
    size_t used = unfilter_pass(raw.data() + off, raw.size() - off, pw, ph, chan, depth, rows);
    if (!used) { im.error = "PNG image data too short or bad filter"; return false; }
    off += used;
    const size_t stride = ((size_t)pw * chan * depth + 7) / 8;
    for (int y = 0; y < ph; y++) {
      const unsigned char* r = rows.data() + stride * (size_t)y;
      for (int x = 0; x < pw; x++) {
        int s[4];
        for (int k = 0; k < chan; k++) {
          const size_t si = (size_t)x * chan + k;
          if (depth == 8) s[k] = r[si];
          else if (depth == 16) s[k] = (r[2 * si] << 8) | r[2 * si + 1];
          else {
            const size_t bit = si * depth;
            s[k] = (r[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
          }
        }
        put_pixel(x0 + x * dx, y0 + y * dy, s);
      }
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// JPEG (ITU-T T.81): baseline / extended sequential Huffman, 8-bit, 1 or 3 components (YCbCr),
// any sampling factors up to 2x2 per axis relative layout, restart intervals.  Progressive and
// arithmetic-coded files are rejected with a message.
// ---------------------------------------------------------------------------------------------
namespace jpeg_detail {

struct HuffTable {
  // 16-bit window lookup: (length << 8) | value, 0 = invalid
  std::vector<uint16_t> tab;
  bool present = false;
  bool build(const unsigned char* counts, const unsigned char* vals) {
    tab.assign(1u << 16, 0);
    uint32_t code = 0;
    int k = 0;
    for (int l = 1; l <= 16; l++) {
      for (int i = 0; i < counts[l - 1]; i++, k++) {
        if (code >> l) return false;
        const uint32_t lo = code << (16 - l), hi = lo + (1u << (16 - l));
        for (uint32_t w = lo; w < hi; w++) tab[w] = (uint16_t)((l << 8) | vals[k]);
        code++;
      }
      code <<= 1;
    }
    present = true;
    return true;
  }
};

struct Bits {  // MSB-first entropy-coded segment reader with FF00 un-stuffing
  const unsigned char* p;
  size_t n, pos;
  uint32_t acc = 0;
  int cnt = 0;
  int marker = 0;  // pending marker (RSTn/EOI...) once hit
  void fill() {
    while (cnt <= 24) {
      int b = 0;
      if (!marker && pos < n) {
        b = p[pos];
        if (b == 0xff) {
          int b2 = pos + 1 < n ? p[pos + 1] : 0xd9;
          if (b2 == 0) pos += 2;
          else if (b2 == 0xff) { pos++; continue; }  // fill byte
          else { marker = b2; b = 0; }
        } else pos++;
      }
      acc |= (uint32_t)b << (24 - cnt);
      cnt += 8;
    }
  }
  uint32_t peek16() {
    if (cnt < 16) fill();
    return acc >> 16;
  }
  void drop(int k) {
    acc <<= k;
    cnt -= k;
  }
  int receive_extend(int s) {  // T.81 F.2.2.1; s <= 16 (callers check the category)
    if (s <= 0 || s > 16) return 0;
    if (cnt < s) fill();
    int v = (int)(acc >> (32 - s));
    drop(s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
  }
  void reset() {
    acc = 0;
    cnt = 0;
    marker = 0;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int bw = 0, bh = 0;  // plane size in blocks (MCU padded)
  int pred = 0;
  std::vector<unsigned char> plane;  // bw*8 x bh*8
};

static const unsigned char zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                         41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                         15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
                                         62, 63};

// exact separable 8x8 IDCT, level shift, clamp
inline void idct_block(const int* coef, const uint16_t* q, unsigned char* out, int stride) {
  static float basis[8][8];
  static bool init = false;
  if (!init) {
    for (int x = 0; x < 8; x++)
      for (int u = 0; u < 8; u++)
        basis[x][u] = (float)((u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0));
    init = true;
  }
  float f[64], t[64];
  for (int i = 0; i < 64; i++) f[i] = (float)(coef[i] * (int)q[i]);
  for (int y = 0; y < 8; y++)  // rows: t[y][x] = sum_u basis[x][u] f[y][u]
    for (int x = 0; x < 8; x++) {
      float s = 0.f;
      for (int u = 0; u < 8; u++) s += basis[x][u] * f[y * 8 + u];
      t[y * 8 + x] = s;
    }
  for (int x = 0; x < 8; x++)
    for (int y = 0; y < 8; y++) {
      float s = 0.f;
      for (int v = 0; v < 8; v++) s += basis[y][v] * t[v * 8 + x];
      int r = (int)std::floor(s + 128.5f);
      out[y * stride + x] = (unsigned char)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

// centred triangle-filter chroma upsampling (the "fancy upsampling" of libjpeg, also stb_image's)
inline void upsample_h2(const unsigned char* in, int w, unsigned char* out) {  // w -> 2w
  if (w == 1) { out[0] = out[1] = in[0]; return; }
  out[0] = in[0];
  out[1] = (unsigned char)((in[0] * 3 + in[1] + 2) >> 2);
  for (int i = 1; i < w - 1; i++) {
    int n3 = 3 * in[i] + 2;
    out[2 * i] = (unsigned char)((n3 + in[i - 1]) >> 2);
    out[2 * i + 1] = (unsigned char)((n3 + in[i + 1]) >> 2);
  }
  // right edge as the reference's reader (stb_image) computes it: weights 3:1 on in[w-2]:in[w-1],
  // not the mirrored 1:3 libjpeg uses — one column of 4:2:2 files differs between the two
  out[2 * w - 2] = (unsigned char)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
  out[2 * w - 1] = in[w - 1];
}
inline void upsample_v2_row(const unsigned char* nearr, const unsigned char* farr, int w, unsigned char* out) {
  for (int i = 0; i < w; i++) out[i] = (unsigned char)((3 * nearr[i] + farr[i] + 2) >> 2);
}
inline void upsample_h2v2_row(const unsigned char* nearr, const unsigned char* farr, int w, unsigned char* out) {
  if (w == 1) { out[0] = out[1] = (unsigned char)((3 * nearr[0] + farr[0] + 2) >> 2); return; }
  int t1 = 3 * nearr[0] + farr[0];
  out[0] = (unsigned char)((t1 + 2) >> 2);
  for (int i = 1; i < w; i++) {
    int t0 = t1;
    t1 = 3 * nearr[i] + farr[i];
    out[2 * i - 1] = (unsigned char)((3 * t0 + t1 + 8) >> 4);
    out[2 * i] = (unsigned char)((3 * t1 + t0 + 8) >> 4);
  }
  out[2 * w - 1] = (unsigned char)((t1 + 2) >> 2);
}

}  // namespace jpeg_detail

inline bool decode_jpeg(const unsigned char* data, size_t n, DecodedImage& im) {
  using namespace jpeg_detail;
  if (n < 4 || data[0] != 0xff || data[1] != 0xd8) { im.error = "not a JPEG"; return false; }
  uint16_t qt[4][64] = {};
  HuffTable hdc[4], hac[4];
  std::vector<Component> comp;
  int W = 0, H = 0, restart = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
  bool have_frame = false, done = false, adobe = false;
  int adobe_transform = -1;
  size_t pos = 2;
  auto be16 = [&](size_t p) { return (data[p] << 8) | data[p + 1]; };
  while (!done) {
    while (pos < n && data[pos] != 0xff) pos++;
    while (pos < n && data[pos] == 0xff) pos++;
    if (pos >= n) break;
    const int m = data[pos++];
    if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) continue;
    if (m == 0xd9) break;
    if (pos + 2 > n) { im.error = "truncated JPEG"; return false; }
    const size_t len = (size_t)be16(pos);
    if (len < 2 || pos + len > n) { im.error = "bad JPEG segment length"; return false; }
    const unsigned char* s = data + pos + 2;
    const size_t sl = len - 2;
    switch (m) {
      case 0xdb: {  // DQT
        size_t i = 0;
        while (i < sl) {
          const int pq = s[i] >> 4, tq = s[i] & 15;
          i++;
          if (tq > 3 || i + (pq ? 128 : 64) > sl) { im.error = "bad DQT"; return false; }
          for (int k = 0; k < 64; k++) {
            qt[tq][zigzag[k]] = pq ? (uint16_t)((s[i] << 8) | s[i + 1]) : s[i];
            i += pq ? 2 : 1;
          }
        }
        break;
      }
      case 0xc4: {  // DHT
        size_t i = 0;
        while (i + 17 <= sl) {
          const int tc = s[i] >> 4, th = s[i] & 15;
          int total = 0;
          for (int k = 0; k < 16; k++) total += s[i + 1 + k];
          if (th > 3 || tc > 1 || i + 17 + total > sl) { im.error = "bad DHT"; return false; }
          if (!(tc ? hac[th] : hdc[th]).build(s + i + 1, s + i + 17)) { im.error = "bad Huffman code lengths"; return false; }
          i += 17 + (size_t)total;
        }
        break;
      }
      case 0xc0: case 0xc1: {  // SOF0 / SOF1
        if (sl < 6 || s[0] != 8) { im.error = "only 8-bit JPEG is supported"; return false; }
        H = be16(pos + 3);
        W = be16(pos + 5);
        const int nc = s[5];
        if ((nc != 1 && nc != 3) || sl < 6 + 3 * (size_t)nc || W <= 0 || H <= 0) { im.error = "unsupported JPEG component count"; return false; }
        comp.assign(nc, Component());
        for (int c = 0; c < nc; c++) {
          comp[c].id = s[6 + 3 * c];
          comp[c].h = s[7 + 3 * c] >> 4;
          comp[c].v = s[7 + 3 * c] & 15;
          comp[c].tq = s[8 + 3 * c];
          if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) { im.error = "bad sampling factors"; return false; }
          hmax = std::max(hmax, comp[c].h);
          vmax = std::max(vmax, comp[c].v);
        }
        for (int c = 0; c < nc; c++)
          if (hmax % comp[c].h != 0 || vmax % comp[c].v != 0) { im.error = "non-integer sampling ratio"; return false; }
        mcux = (W + 8 * hmax - 1) / (8 * hmax);
        mcuy = (H + 8 * vmax - 1) / (8 * vmax);
        for (auto& c : comp) {
          c.bw = mcux * c.h;
          c.bh = mcuy * c.v;
          c.plane.assign((size_t)c.bw * 8 * c.bh * 8, 0);
        }
        have_frame = true;
        break;
      }
      case 0xc2: im.error = "progressive JPEG is not supported"; return false;
      case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
        im.error = "unsupported JPEG coding process";
        return false;
      case 0xdd: restart = sl >= 2 ? be16(pos + 2) : 0; break;
      case 0xee:
        if (sl >= 12 && !memcmp(s, "Adobe", 5)) { adobe = true; adobe_transform = s[11]; }
        break;
      case 0xda: {  // SOS + entropy-coded data
        if (!have_frame) { im.error = "SOS before SOF"; return false; }
        if (sl < 1 || sl < 1 + 2 * (size_t)s[0] + 3) { im.error = "short SOS segment"; return false; }
        const int ns = s[0];
        if (ns != (int)comp.size() && !(ns == 1 && comp.size() == 1)) {
          im.error = "non-interleaved multi-scan JPEG is not supported";
          return false;
        }
        for (int k = 0; k < ns; k++) {
          const int cid = s[1 + 2 * k];
          bool found = false;
          for (auto& c : comp)
            if (c.id == cid) {
              c.td = s[2 + 2 * k] >> 4;
              c.ta = s[2 + 2 * k] & 15;
              found = c.td <= 3 && c.ta <= 3 && hdc[c.td].present && hac[c.ta].present;
            }
          if (!found) { im.error = "scan refers to a missing component or table"; return false; }
        }
        Bits br{data, n, pos + len};
        for (auto& c : comp) c.pred = 0;
        int todo = restart;
        // a single-component scan is not interleaved: its MCU is one 8x8 block (T.81 A.2.2)
        const bool single = comp.size() == 1;
        const int nx = single ? (W + 7) / 8 : mcux, ny = single ? (H + 7) / 8 : mcuy;
        for (int my = 0; my < ny; my++)
          for (int mx = 0; mx < nx; mx++) {
            for (auto& c : comp) {
              const int bh_ = single ? 1 : c.h, bv_ = single ? 1 : c.v;
              for (int by = 0; by < bv_; by++)
                for (int bx = 0; bx < bh_; bx++) {
                  int coef[64] = {0};
                  uint16_t e = hdc[c.td].tab[br.peek16()];
                  if (!e) { im.error = "bad DC Huffman code"; return false; }
                  br.drop(e >> 8);
                  if ((e & 0xff) > 11) { im.error = "DC category above 11 (baseline JPEG)"; return false; }
                  c.pred += br.receive_extend(e & 0xff);
                  coef[0] = c.pred;
                  for (int k = 1; k < 64;) {
                    e = hac[c.ta].tab[br.peek16()];
                    if (!e) { im.error = "bad AC Huffman code"; return false; }
                    br.drop(e >> 8);
                    const int r = (e & 0xff) >> 4, sz = e & 15;
                    if (sz == 0) {
                      if (r != 15) break;  // EOB
                      k += 16;
                      continue;
                    }
                    k += r;
                    if (k > 63) { im.error = "AC coefficient index overflow"; return false; }
                    coef[zigzag[k]] = br.receive_extend(sz);
                    k++;
                  }
                  const int px = (mx * bh_ + bx) * 8, py = (my * bv_ + by) * 8;
                  if (px + 8 <= c.bw * 8 && py + 8 <= c.bh * 8)
                    idct_block(coef, qt[c.tq], &c.plane[(size_t)py * c.bw * 8 + px], c.bw * 8);
                }
            }
            if (restart && --todo == 0) {
              todo = restart;
              // byte-align and consume the RSTn marker
              br.cnt = 0;
              br.acc = 0;
              if (!br.marker) {  // marker not yet seen by the reader: scan forward to it
                while (br.pos + 1 < n && !(data[br.pos] == 0xff && data[br.pos + 1] >= 0xd0 && data[br.pos + 1] <= 0xd7)) br.pos++;
              }
              if (br.pos + 1 < n && data[br.pos] == 0xff && data[br.pos + 1] >= 0xd0 && data[br.pos + 1] <= 0xd7) br.pos += 2;
              br.reset();
              for (auto& c : comp) c.pred = 0;
            }
          }
        done = true;
        break;
      }
      default: break;  // APPn, COM, ...
    }
    pos += len;
  }
  if (!done) { im.error = im.error.empty() ? "JPEG without image data" : im.error; return false; }

  const int nc = (int)comp.size();
  im.w = W;
  im.h = H;
  im.n = nc;
  im.px.assign((size_t)W * H * nc, 0);
  // per-component full-resolution rows
  std::vector<std::vector<unsigned char>> row(nc, std::vector<unsigned char>((size_t)mcux * hmax * 8 + 16));
  const bool ycc = nc == 3 && !(adobe && adobe_transform == 0);
  for (int y = 0; y < H; y++) {
    for (int c = 0; c < nc; c++) {
      const Component& C = comp[c];
      const int pw = C.bw * 8, ph = C.bh * 8;
      const int cw = (W * C.h + hmax - 1) / hmax;  // samples of this component that carry image data
      const int chh = (H * C.v + vmax - 1) / vmax;
      const int hs = hmax / C.h, vs = vmax / C.v;
      unsigned char* out = row[c].data();
      if (hs == 1 && vs == 1) {
        memcpy(out, &C.plane[(size_t)y * pw], (size_t)W);
      } else if ((hs == 1 || hs == 2) && (vs == 1 || vs == 2) && hmax % C.h == 0 && vmax % C.v == 0) {
        const unsigned char* nr;
        const unsigned char* fr;
        if (vs == 2) {
          const int sy = y >> 1;
          const int fy = (y & 1) ? std::min(sy + 1, chh - 1) : std::max(sy - 1, 0);
          nr = &C.plane[(size_t)sy * pw];
          fr = &C.plane[(size_t)fy * pw];
        } else {
          nr = fr = &C.plane[(size_t)y * pw];
        }
        if (hs == 2 && vs == 2) upsample_h2v2_row(nr, fr, cw, out);
        else if (hs == 2) upsample_h2(nr, cw, out);
        else upsample_v2_row(nr, fr, cw, out);
      } else {  // other ratios: nearest sample
        for (int x = 0; x < W; x++) {
          const int sx = std::min(x * C.h / hmax, pw - 1), sy = std::min(y * C.v / vmax, ph - 1);
          out[x] = C.plane[(size_t)sy * pw + sx];
        }
      }
    }
    unsigned char* o = &im.px[(size_t)y * W * nc];
    if (nc == 1) memcpy(o, row[0].data(), (size_t)W);
    else
      for (int x = 0; x < W; x++) {
        if (!ycc) {
          o[3 * x] = row[0][x]; o[3 * x + 1] = row[1][x]; o[3 * x + 2] = row[2][x];
          continue;
        }
        const float Y = row[0][x], cb = row[1][x] - 128.f, cr = row[2][x] - 128.f;
        const int r = (int)std::floor(Y + 1.402f * cr + 0.5f);
        const int g = (int)std::floor(Y - 0.344136f * cb - 0.714136f * cr + 0.5f);
        const int b = (int)std::floor(Y + 1.772f * cb + 0.5f);
        o[3 * x] = (unsigned char)(r < 0 ? 0 : (r > 255 ? 255 : r));
        o[3 * x + 1] = (unsigned char)(g < 0 ? 0 : (g > 255 ? 255 : g));
        o[3 * x + 2] = (unsigned char)(b < 0 ? 0 : (b > 255 ? 255 : b));
      }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// binary PPM (P6, maxval 255)
// ---------------------------------------------------------------------------------------------
inline bool decode_ppm(const unsigned char* data, size_t n, DecodedImage& im) {
  if (n < 2 || data[0] != 'P' || data[1] != '6') { im.error = "not a P6 PPM"; return false; }
  size_t pos = 2;
  int vals[3], got = 0;
  while (got < 3 && pos < n) {
    if (data[pos] == '#') { while (pos < n && data[pos] != '\n') pos++; continue; }
    if (isspace(data[pos])) { pos++; continue; }
    if (!isdigit(data[pos])) { im.error = "bad PPM header"; return false; }
    int v = 0;
    while (pos < n && isdigit(data[pos])) v = v * 10 + (data[pos++] - '0');
    vals[got++] = v;
  }
  pos++;  // the single whitespace byte after maxval
  if (got < 3 || vals[2] != 255 || vals[0] <= 0 || vals[1] <= 0 || pos + (size_t)vals[0] * vals[1] * 3 > n) { im.error = "bad or truncated PPM"; return false; }
  im.w = vals[0];
  im.h = vals[1];
  im.n = 3;
  im.px.assign(data + pos, data + pos + (size_t)im.w * im.h * 3);
  return true;
}

// the stbi_load(file, &nx, &ny, &nn, 0) stand-in: format chosen by signature
inline bool decode_image_file(const std::string& path, DecodedImage& im) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { im.error = "cannot open " + path; return false; }
  std::vector<unsigned char> buf;
  unsigned char chunk[65536];
  size_t got;
  while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
  fclose(f);
  if (buf.size() >= 8 && buf[0] == 0x89 && buf[1] == 'P') return decode_png(buf.data(), buf.size(), im);
  if (buf.size() >= 4 && buf[0] == 0xff && buf[1] == 0xd8) return decode_jpeg(buf.data(), buf.size(), im);
  if (buf.size() >= 2 && buf[0] == 'P' && buf[1] == '6') return decode_ppm(buf.data(), buf.size(), im);
  im.error = "unknown image format (PNG, baseline JPEG and binary PPM are supported)";
  return false;
}

}  // namespace ptio
