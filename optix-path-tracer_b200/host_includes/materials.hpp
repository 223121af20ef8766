// materials.hpp — host material (BRDF) classes (reference surface:
// host_includes/materials.hpp:27-242).  assignTo() registers one pt_material;
// parameter conversions (Oren-Nayar sigma -> rA/rB, roughness -> GGX alpha,
// Ashikhmin-Shirley exponent clamp) are done here exactly where the reference
// does them (materials.hpp:173-174,192-198,235-238).
#pragma once
#include "host_common.hpp"
#include "textures.hpp"

struct BRDF {
  virtual ~BRDF() {}
  virtual Material assignTo(Context& g_context) const = 0;

 protected:
  Material registerOnce(Context& ctx, int type, const Texture* t0, const Texture* t1, float p0, float p1,
                        int flags = 0) const {
    auto it = ctx->material_cache.find(this);
    if (it != ctx->material_cache.end()) return Material{it->second};
    pt_material m{};
    m.type = type;
    m.tex0 = t0 ? t0->assignTo(ctx)->texture_id : -1;
    m.tex1 = t1 ? t1->assignTo(ctx)->texture_id : -1;
    m.p0 = p0;
    m.p1 = p1;
    m.flags = flags;
    int id = (int)ctx->materials.size();
    ctx->materials.push_back(m);
    ctx->material_cache[this] = id;
    return Material{id};
  }
};

struct Lambertian : public BRDF {
  Lambertian(const Texture* t) : texture(t) {}
  Material assignTo(Context& c) const override { return registerOnce(c, PT_MAT_LAMBERTIAN, texture, nullptr, 0, 0); }
  const Texture* texture;
};

struct Metal : public BRDF {
  Metal(const Texture* t, const float fuzz) : texture(t), fuzz(fuzz) {}
  Material assignTo(Context& c) const override { return registerOnce(c, PT_MAT_METAL, texture, nullptr, fuzz, 0); }
  const Texture* texture;
  const float fuzz;
};

struct Dielectric : public BRDF {
  Dielectric(const Texture* baseTex, const Texture* extTex, const float ref_idx, const float density = 0.f)
      : baseTex(baseTex), extTex(extTex), ref_idx(ref_idx), density(density) {}
  Material assignTo(Context& c) const override {
    return registerOnce(c, PT_MAT_DIELECTRIC, baseTex, extTex, ref_idx, density);
  }
  const Texture *baseTex, *extTex;
  const float ref_idx, density;
};

struct Diffuse_Light : public BRDF {
  Diffuse_Light(const Texture* t) : texture(t) {}
  Material assignTo(Context& c) const override { return registerOnce(c, PT_MAT_DIFFUSE_LIGHT, texture, nullptr, 0, 0); }
  const Texture* texture;
};

// phase function of the volumes (reference: host_includes/materials.hpp:125-141)
struct Isotropic : public BRDF {
  Isotropic(const Texture* t) : texture(t) {}
  Material assignTo(Context& c) const override { return registerOnce(c, PT_MAT_ISOTROPIC, texture, nullptr, 0, 0); }
  const Texture* texture;
};

struct Normal_Shader : public BRDF {
  Normal_Shader(const bool useShadingNormal = false) : useShadingNormal(useShadingNormal) {}
  Material assignTo(Context& c) const override {
    return registerOnce(c, PT_MAT_NORMAL, nullptr, nullptr, 0, 0, useShadingNormal ? 1 : 0);
  }
  const bool useShadingNormal;
};

struct Ashikhmin_Shirley : public BRDF {
  Ashikhmin_Shirley(const Texture* diffuse_tex, const Texture* specular_tex, const float nu, const float nv)
      : diffuse_tex(diffuse_tex), specular_tex(specular_tex), nu(nu), nv(nv) {}
  Material assignTo(Context& c) const override {
    return registerOnce(c, PT_MAT_ASHIKHMIN_SHIRLEY, diffuse_tex, specular_tex, fmaxf(1.f, nu), fmaxf(1.f, nv));
  }
  const Texture* diffuse_tex;
  const Texture* specular_tex;
  const float nu, nv;
};

struct Oren_Nayar : public BRDF {
  Oren_Nayar(const Texture* t, const float R) : texture(t) {
    float sigma = saturate(R);
    float div = 1.f / (PI_F + ((3.f * PI_F - 4.f) / 6.f) * sigma);
    rA = 1.0f * div;
    rB = sigma * div;
  }
  Material assignTo(Context& c) const override { return registerOnce(c, PT_MAT_OREN_NAYAR, texture, nullptr, rA, rB); }
  const Texture* texture;
  float rA, rB;
};

struct Torrance_Sparrow : public BRDF {
  Torrance_Sparrow(const Texture* texture, const float nu, const float nv) : texture(texture), nu(nu), nv(nv) {}
  Material assignTo(Context& c) const override {
    return registerOnce(c, PT_MAT_TORRANCE_SPARROW, texture, nullptr, roughnessToAlpha(nu), roughnessToAlpha(nv));
  }
  float roughnessToAlpha(float roughness) const {
    float R = fmaxf(roughness, 1e-3f);
    return R * R;
  }
  const Texture* texture;
  const float nu, nv;
};
