// textures.hpp — host texture classes (reference surface:
// host_includes/textures.hpp:16-290).  assignTo() registers the texture in the
// context's flat texture table and returns a Program handle carrying its id.
#pragma once
#include "host_common.hpp"
#include "io_formats.hpp"

struct Texture {
  virtual ~Texture() {}
  virtual Program assignTo(Context& g_context) const = 0;

 protected:
  // one table entry per Texture object, however many materials share it
  Program registerOnce(Context& ctx, const pt_texture& t) const {
    auto it = ctx->texture_cache.find(this);
    int id;
    if (it != ctx->texture_cache.end()) {
      id = it->second;
    } else {
      id = (int)ctx->textures.size();
      ctx->textures.push_back(t);
      ctx->texture_cache[this] = id;
    }
    Program p{std::make_shared<ProgramObj>()};
    p->texture_id = id;
    return p;
  }
};

struct Constant_Texture : public Texture {
  Constant_Texture(const float3& c) : color(c) {}
  Constant_Texture(const float& rgb) : color(make_float3(rgb)) {}
  Constant_Texture(const float& r, const float& g, const float& b) : color(make_float3(r, g, b)) {}
  Program assignTo(Context& ctx) const override {
    pt_texture t{};
    t.type = PT_TEX_CONSTANT;
    t.color[0] = color.x; t.color[1] = color.y; t.color[2] = color.z;
    return registerOnce(ctx, t);
  }
  const float3 color;
};

struct Checker_Texture : public Texture {
  Checker_Texture(const Texture* o, const Texture* e) : odd(o), even(e) {}
  Program assignTo(Context& ctx) const override {
    pt_texture t{};
    t.type = PT_TEX_CHECKER;
    t.a = odd->assignTo(ctx)->texture_id;
    t.b = even->assignTo(ctx)->texture_id;
    return registerOnce(ctx, t);
  }
  const Texture* odd;
  const Texture* even;
};

// Perlin turbulence texture (host_includes/textures.hpp:64-108, programs/textures/noise_texture.cu).
// The gradient and permutation tables are drawn from rnd() exactly like the reference does
// (256 gradients, z component drawn first as g++/MSVC evaluate the argument list right to left,
// then perm_x, perm_y, perm_z by Fisher-Yates) and stored in the context's texture-data table.
struct Noise_Texture : public Texture {
  Noise_Texture(const float s, const AXIS ax = X_AXIS) : scale(s), ax(ax) {}
  Program assignTo(Context& ctx) const override {
    auto it = ctx->texture_cache.find(this);
    pt_texture t{};
    t.type = PT_TEX_NOISE;
    t.color[0] = scale;
    t.b = (int)ax;
    if (it == ctx->texture_cache.end()) {
      t.a = (int)ctx->texdata.size();
      for (int i = 0; i < 256; ++i) {
        float z = -1 + 2 * rnd(), y = -1 + 2 * rnd(), x = -1 + 2 * rnd();
        float l = sqrtf(x * x + y * y + z * z);
        ctx->texdata.push_back(x / l);
        ctx->texdata.push_back(y / l);
        ctx->texdata.push_back(z / l);
      }
      for (int table = 0; table < 3; table++) {
        int p[256];
        for (int i = 0; i < 256; i++) p[i] = i;
        for (int i = 256 - 1; i > 0; i--) {
          int target = int(rnd() * (i + 1));
          int tmp = p[i];
          p[i] = p[target];
          p[target] = tmp;
        }
        for (int i = 0; i < 256; i++) ctx->texdata.push_back((float)p[i]);
      }
    }
    return registerOnce(ctx, t);
  }
  const float scale;
  const AXIS ax;
};

struct Image_Texture : public Texture {
  Image_Texture(const std::string f) : fileName(f) {}
  Program assignTo(Context& ctx) const override {
    auto it = ctx->texture_cache.find(this);
    if (it == ctx->texture_cache.end()) {
      ContextObj::Image im;
      // reference flips LDR images vertically on load (textures.hpp:143-156)
      if (!ptio::load_ldr_image(fileName, im.w, im.h, im.rgba, /*flip_y=*/true)) {
        printf("Image is invalid or hasn't been found.\n");
        throw "Image_Texture: cannot load " + fileName;
      }
      ctx->images.push_back(std::move(im));
    }
    pt_texture t{};
    t.type = PT_TEX_IMAGE;
    t.a = (int)ctx->images.size() - 1;
    return registerOnce(ctx, t);
  }
  const std::string fileName;
};

struct HDR_Texture : public Texture {
  HDR_Texture(const std::string f) : fileName(f) {}
  Program assignTo(Context& ctx) const override {
    auto it = ctx->texture_cache.find(this);
    if (it == ctx->texture_cache.end()) {
      ContextObj::Image im;
      if (!ptio::load_radiance_hdr(fileName, im.w, im.h, im.rgba)) {
        printf("HDR Image is invalid or hasn't been found.\n");
        throw "HDR_Texture: cannot load " + fileName;
      }
      ctx->images.push_back(std::move(im));
    }
    pt_texture t{};
    t.type = PT_TEX_IMAGE;
    t.a = (int)ctx->images.size() - 1;
    return registerOnce(ctx, t);
  }
  const std::string fileName;
};

// |unit(p)| weighted blend of three colours (host_includes/textures.hpp:236-252, gradient_texture.cu)
struct Gradient_Texture : public Texture {
  Gradient_Texture(const float3& cA, const float3& cB, const float3& cC) : colorA(cA), colorB(cB), colorC(cC) {}
  Program assignTo(Context& ctx) const override {
    auto it = ctx->texture_cache.find(this);
    pt_texture t{};
    t.type = PT_TEX_GRADIENT;
    if (it == ctx->texture_cache.end()) {
      t.a = (int)ctx->texdata.size();
      for (const float3& c : {colorA, colorB, colorC}) {
        ctx->texdata.push_back(c.x);
        ctx->texdata.push_back(c.y);
        ctx->texdata.push_back(c.z);
      }
    }
    return registerOnce(ctx, t);
  }
  const float3 colorA, colorB, colorC;
};

struct Vector_Texture : public Texture {
  Vector_Texture(const std::vector<Texture*>& tv) : texture_vector(tv) {}
  Program assignTo(Context& ctx) const override {
    auto it = ctx->texture_cache.find(this);
    pt_texture t{};
    t.type = PT_TEX_VECTOR;
    if (it == ctx->texture_cache.end()) {
      std::vector<int> ids;
      for (Texture* s : texture_vector) ids.push_back(s->assignTo(ctx)->texture_id);
      t.a = (int)ctx->texvec.size();
      t.b = (int)ids.size();
      ctx->texvec.insert(ctx->texvec.end(), ids.begin(), ids.end());
    }
    return registerOnce(ctx, t);
  }
  const std::vector<Texture*> texture_vector;
};

struct Texture_List {
  Texture_List() {}
  int push(Texture* t) {
    texList.push_back(t);
    return (int)texList.size() - 1;
  }
  Texture* operator[](const int i) { return texList[i]; }
  std::vector<Texture*> texList;
};
