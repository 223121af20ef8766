// shading.h — device shading cores of the wavefront integrator: hit-record
// reconstruction, textures, BSDF sample/evaluate, rect-light NEE + MIS, miss
// programs, camera rays.  PT_HD so tests/hostsim can run them on the CPU.
//
// Replaces the reference's closest-hit / miss / callable programs
// (programs/materials/*.cu,*.cuh, programs/miss.cu, programs/textures/*.cu,
// programs/pdfs/rect_pdf.cu, programs/raygen.cu:50-64).  RNG draw order per
// event follows SURVEY.md §8a exactly (single LCG stream per pixel-sample).
//
// Wavefront twist (B200-first, not in the reference): Direct_Light's RNG
// consumption depends on the shadow-ray result (light_sample.cuh:72 returns
// before the BSDF draws).  Instead of stalling the path, shade_hit() computes
// BOTH continuations — V (light visible: MIS terms added, bounce from draws
// 6,7) and S (shadowed: no direct light, bounce from draws 4,5, which is the
// very direction V's BSDF-MIS sample used) — and the shadow kernel picks one.
#pragma once
#include "scene_dev.h"

// ---- RNG: programs/random.cuh:33-65 ------------------------------------------------
PT_HD uint32_t pt_tea64(uint32_t val0, uint32_t val1) {
  uint32_t v0 = val0, v1 = val1, s0 = 0;
  for (int n = 0; n < 64; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
PT_HD float pt_rnd(uint32_t& prev) {
  prev = 1664525u * prev + 1013904223u;
  return (float)(prev & 0x00FFFFFFu) / (float)0x01000000;
}

// ---- sampling: programs/sampling.cuh:22-59 --------------------------------------------
PT_HD float3 pt_random_in_unit_disk(uint32_t& seed) {
  float a = pt_rnd(seed) * 2.f * PT_PI_F;
  float3 xy = f3(sinf(a), cosf(a), 0.f);
  xy *= sqrtf(pt_rnd(seed));
  return xy;
}
PT_HD float3 pt_random_in_unit_sphere(uint32_t& seed) {
  float z = pt_rnd(seed) * 2.f - 1.f;
  float t = pt_rnd(seed) * 2.f * PT_PI_F;
  float r = sqrtf((0.f > (1.f - z * z) ? 0.f : (1.f - z * z)));
  float x = r * cosf(t);
  float y = r * sinf(t);
  float3 res = f3(x, y, z);
  res *= powf(pt_rnd(seed), 1.f / 3.f);
  return res;
}
// optixu_math: cosine_sample_hemisphere + Onb::inverse_transform
PT_HD float3 pt_cosine_dir(float3 N, float u1, float u2) {
  const float r = sqrtf(u1);
  const float phi = 2.0f * PT_OPTIX_PI * u2;
  float3 p;
  p.x = r * cosf(phi);
  p.y = r * sinf(phi);
  p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
  float3 bn;
  if (fabsf(N.x) > fabsf(N.z)) bn = f3(-N.y, N.x, 0.f);
  else bn = f3(0.f, -N.z, N.y);
  bn = normalize(bn);
  float3 tg = cross(bn, N);
  return p.x * tg + p.y * bn + p.z * N;
}

// ---- world-Y shading-frame trig: programs/math/trigonometric.cuh:5-43 --------------------
PT_HD float sfCos2Theta(float3 w) { return w.y * w.y; }
PT_HD float sfAbsCosTheta(float3 w) { return fabsf(w.y); }
PT_HD float sfSin2Theta(float3 w) { return fmaxf(0.f, 1.f - sfCos2Theta(w)); }
PT_HD float sfSinTheta(float3 w) { return sqrtf(sfSin2Theta(w)); }
PT_HD float sfTanTheta(float3 w) { return sfSinTheta(w) / w.y; }
PT_HD float sfCosPhi(float3 w) {
  float st = sfSinTheta(w);
  return (st == 0) ? 1.f : clampf(w.x / st, -1.f, 1.f);
}
PT_HD float sfSinPhi(float3 w) {
  float st = sfSinTheta(w);
  return (st == 0) ? 0.f : clampf(w.z / st, -1.f, 1.f);
}

// ---- programs/materials/material.cuh:30-63 ------------------------------------------------
PT_HD float pt_schlick(float cosine, float ref_idx) {
  float r0 = (1.f - ref_idx) / (1.f + ref_idx);
  r0 = r0 * r0;
  return r0 + (1.f - r0) * powf((1.f - cosine), 5.f);
}
PT_HD float3 pt_schlick3(float3 r0, float cosine) {
  float e = powf(1.f - cosine, 5.f);
  return r0 + (f3(1.f) - r0) * e;
}
PT_HD bool pt_refract(float3 v, float3 n, float ni_over_nt, float3& refracted) {
  float3 uv = normalize(v);
  float dt = dot(uv, n);
  float disc = 1.f - ni_over_nt * ni_over_nt * (1.f - dt * dt);
  if (disc > 0.f) {
    refracted = ni_over_nt * (uv - n * dt) - n * sqrtf(disc);
    return true;
  }
  return false;
}

// ---- programs/materials/microfacets.cuh -------------------------------------------------
PT_HD void pt_sample_quadrant(float nu, float nv, float rx, float ry, float& phi, float& theta) {
  phi = atanf(sqrtf((nu + 1.f) / (nv + 1.f)) * tanf(2.f * PT_PI_F * rx));
  float cp = cosf(phi), sp = sinf(phi);
  theta = powf(ry, 1.f / (nu * cp * cp + nv * sp * sp + 1.f));
}
PT_HD float pt_ggx_d(float3 H, float nu, float nv) {
  const float c2 = sfCos2Theta(H);
  if (c2 <= 0.0f) return 0.f;
  if (nu == nv) {
    float beta = c2 * ((nu * nu) - 1) + 1;
    return (nu * nu) / (PT_PI_F * beta * beta);
  }
  float sp = sfSinPhi(H);
  const float s2nv = (sp * sp) / (nv * nv);
  const float c2nu = (1.f - sp * sp) / (nu * nu);
  float beta = (c2 + (1.f - c2) * (c2nu + s2nv));
  return 1.f / (PT_PI_F * nu * nv * beta * beta);
}
PT_HD float3 pt_ggx_sample(float3 origin, float rx, float ry, float nu, float nv) {
  bool flip = origin.y < 0;
  float3 so = normalize(f3(origin.x * nu, origin.y, origin.z * nv));
  float slopeX, slopeY;
  float cosTheta = so.y;
  if (cosTheta > 0.9999f) {
    float r = sqrtf(rx / (1 - rx));
    float phi = (float)(6.28318530718 * (double)ry);
    slopeX = r * cosf(phi);
    slopeY = r * sinf(phi);
  } else {
    float sinTheta = sqrtf(fmaxf(0.f, 1.f - cosTheta * cosTheta));
    float tanTheta = sinTheta / cosTheta;
    float a = 1.f / tanTheta;
    float G1 = 2.f / (1.f + sqrtf(1.f + 1.f / (a * a)));
    float A = 2.f * rx / G1 - 1.f;
    float tmp = 1.f / (A * A - 1.f);
    if (tmp > 1e10f) tmp = 1e10f;
    float B = tanTheta;
    float D = sqrtf(fmaxf(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
    float s1 = B * tmp - D;
    float s2 = B * tmp + D;
    slopeX = (A < 0 || s2 > 1.f / tanTheta) ? s1 : s2;
    float S;
    if (ry > 0.5f) {
      S = 1.f;
      ry = 2.f * (ry - .5f);
    } else {
      S = -1.f;
      ry = 2.f * (.5f - ry);
    }
    float z = (ry * (ry * (ry * 0.27385f - 0.73369f) + 0.46341f)) /
              (ry * (ry * (ry * 0.093073f + 0.309420f) - 1.f) + 0.597999f);
    slopeY = S * z * sqrtf(1.f + slopeX * slopeX);
  }
  float cp = sfCosPhi(so), sp = sfSinPhi(so);
  float t = cp * slopeX - sp * slopeY;
  slopeY = sp * slopeX + cp * slopeY;
  slopeX = t;
  slopeX = nu * slopeX;
  slopeY = nv * slopeY;
  float3 H = normalize(f3(-slopeX, 1.f, -slopeY));
  if (flip) H *= -1.f;
  return H;
}
PT_HD float pt_ggx_lambda(float3 V, float nu, float nv) {
  float att = fabsf(sfTanTheta(V));
  if (isinf(att)) return 0.f;
  float cp = sfCosPhi(V), sp = sfSinPhi(V);
  float alpha = sqrtf(cp * cp * nu * nu + sp * sp * nv * nv);
  float a2t2 = (alpha * att) * (alpha * att);
  return (-1.f + sqrtf(1.f + a2t2)) / 2.f;
}
PT_HD float pt_ggx_g(float3 Wo, float3 Wi, float nu, float nv) {
  return 1.f / (1.f + pt_ggx_lambda(Wo, nu, nv) + pt_ggx_lambda(Wi, nu, nv));
}

// ---- BSDFs -------------------------------------------------------------------------------
struct Surface {
  float3 color, color2;
  float p0, p1;
};

template <int MAT>
PT_HD float3 bsdf_sample(const Surface& s, float3 Wo, float3 N, uint32_t& seed) {
  if (MAT == PT_MAT_DIFFUSE_LIGHT || MAT == PT_MAT_ISOTROPIC) {
    // diffuse_light.cuh:7-13, isotropic.cuh:7-13: random_on_unit_sphere (3 draws)
    float3 r = pt_random_in_unit_sphere(seed);
    return r / length(r);
  }
  float u1 = pt_rnd(seed), u2 = pt_rnd(seed);  // device (left-to-right) draw order
  if (MAT == PT_MAT_TORRANCE_SPARROW) {  // torrance_sparrow.cuh:10-37
    float3 Nn = normalize(N);
    float3 T = normalize(cross(Nn, f3(0.f, 1.f, 0.f)));
    float3 B = cross(T, Nn);
    float3 H = normalize(pt_ggx_sample(Wo, u1, u2, s.p0, s.p1));
    H = H.x * B + H.y * Nn + H.z * T;
    if (dot(H, Wo) < 0.f) H = -H;
    return normalize(-Wo + 2.f * dot(Wo, H) * H);
  }
  if (MAT == PT_MAT_ASHIKHMIN_SHIRLEY) {  // ashikhmin_shirley.cuh:11-85
    float nu = s.p0, nv = s.p1;
    float3 Nn = normalize(N);
    float3 T = normalize(cross(Nn, f3(0.f, 1.f, 0.f)));
    float3 B = cross(T, Nn);
    float rx = u1, ry = u2;
    if ((double)rx < 0.5) {
      rx = fminf(2 * rx, 1.f - 1e-6f);
      return pt_cosine_dir(Nn, rx, ry);
    }
    rx = fminf(2 * (rx - 0.5f), 1.f - 1e-6f);
    float phi, cos_theta;
    if (nu == nv) {
      phi = 2.f * PT_PI_F * rx;
      cos_theta = powf(ry, 1.f / (nu + 1.f));
    } else if (rx < 0.25f) {
      pt_sample_quadrant(nu, nv, 4.f * rx, ry, phi, cos_theta);
    } else if (rx < 0.5f) {
      pt_sample_quadrant(nu, nv, 4.f * (0.5f - rx), ry, phi, cos_theta);
      phi = PT_PI_F - phi;
    } else if (rx < 0.75f) {
      pt_sample_quadrant(nu, nv, 4.f * (rx - 0.5f), ry, phi, cos_theta);
      phi = PT_PI_F + phi;
    } else {
      pt_sample_quadrant(nu, nv, 4.f * (1.f - rx), ry, phi, cos_theta);
      phi = 2.f * PT_PI_F - phi;
    }
    float sin_theta = sqrtf(fmaxf(0.f, 1.f - cos_theta * cos_theta));
    float3 H = normalize(f3(sin_theta * cosf(phi), cos_theta, sin_theta * sinf(phi)));
    H = H.x * B + H.y * Nn + H.z * T;
    if (dot(H, Wo) < 0.f) H = -H;
    return normalize(-Wo + 2.f * dot(Wo, H) * H);
  }
  return pt_cosine_dir(N, u1, u2);  // lambertian.cuh:7-19, oren_nayar.cuh:8-20
}

// pdf is left untouched on early-outs (callers initialise it to 0: pinned)
template <int MAT>
PT_HD float3 bsdf_eval(const Surface& s, float3 Wo, float3 Wi, float3 N, float& pdf) {
  if (MAT == PT_MAT_ISOTROPIC) {  // isotropic.cuh:15-33: pdf = 0.25 * PI_F, f = 0.25 * PI_F * color
    pdf = 0.25f * PT_PI_F;
    return (0.25f * PT_PI_F) * s.color;
  }
  if (MAT == PT_MAT_LAMBERTIAN) {  // lambertian.cuh:29-43
    float cosine = dot(normalize(Wi), normalize(N));
    if (cosine < 0.f) return f3(0.f);
    pdf = dot(normalize(Wi), normalize(N)) / PT_PI_F;
    return (cosine * s.color) / PT_PI_F;
  }
  if (MAT == PT_MAT_OREN_NAYAR) {  // oren_nayar.cuh:30-65
    float3 WiN = normalize(Wi);
    float sinThetaI = sfSinTheta(WiN), sinThetaO = sfSinTheta(Wo);
    float maxCos = 0;
    if ((double)sinThetaI > 1e-4 && (double)sinThetaO > 1e-4) {
      float dCos = sfCosPhi(WiN) * sfCosPhi(Wo) + sfSinPhi(WiN) * sfSinPhi(Wo);
      maxCos = fmaxf(0.f, dCos);
    }
    float sinAlpha, tanBeta;
    if (sfAbsCosTheta(WiN) > sfAbsCosTheta(Wo)) {
      sinAlpha = sinThetaO;
      tanBeta = sinThetaI / sfAbsCosTheta(WiN);
    } else {
      sinAlpha = sinThetaI;
      tanBeta = sinThetaO / sfAbsCosTheta(Wo);
    }
    pdf = dot(normalize(Wi), normalize(N)) / PT_PI_F;
    return s.color * (1.f / PT_PI_F) * (s.p0 + s.p1 * maxCos * sinAlpha * tanBeta);
  }
  if (MAT == PT_MAT_TORRANCE_SPARROW) {  // torrance_sparrow.cuh:39-65
    float3 Rs = s.color;
    float3 Up = f3(0.f, 1.f, 0.f);
    float NdotI = fmaxf(dot(Up, Wi), 1e-6f), NdotO = fmaxf(dot(Up, Wo), 1e-6f);
    float3 H = Wo + Wi;
    if (H.x == 0.f && H.y == 0.f && H.z == 0.f) return f3(0.f);
    H = normalize(H);
    float HdotI = fabsf(dot(H, Wi));
    float3 F = pt_schlick3(Rs, HdotI);
    float G = pt_ggx_g(Wo, Wi, s.p0, s.p1);
    float D = pt_ggx_d(H, s.p0, s.p1);
    pdf = (pt_ggx_d(H, s.p0, s.p1) * sfAbsCosTheta(H)) / (4.f * dot(Wo, H));
    return Rs * D * G * F / (4.f * NdotI * NdotO);
  }
  if (MAT == PT_MAT_ASHIKHMIN_SHIRLEY) {  // ashikhmin_shirley.cuh:87-151
    float3 Rd = s.color, Rs = s.color2;
    float nu = s.p0, nv = s.p1;
    float3 Up = f3(0.f, 1.f, 0.f);
    float3 Nn = normalize(N);
    float3 T = normalize(cross(Nn, Up));
    float3 B = cross(T, Nn);
    float NdotI = fabsf(dot(Up, Wi)), NdotO = fabsf(dot(Up, Wo));
    float3 dc = Rd * (28.f / (23.f * PT_PI_F));
    dc *= (f3(1.f) - Rs);
    dc *= (1.f - powf(1.f - NdotI * 0.5f, 5.f));
    dc *= (1.f - powf(1.f - NdotO * 0.5f, 5.f));
    float dpdf = dot(Wi, Nn) / PT_PI_F;
    if (dpdf < 0.f) return f3(0.f);
    float3 H = Wo + Wi;
    if (H.x == 0.f && H.y == 0.f && H.z == 0.f) return f3(0.f);
    H = normalize(H);
    float HdotI = fabsf(dot(H, Wi));
    float HdotN = fabsf(dot(H, Nn)), HdotT = fabsf(dot(H, T)), HdotB = fabsf(dot(H, B));
    float norm, lobe;
    if (nu == nv) {
      norm = (nu + 1.f) / (8.f * PT_PI_F);
      lobe = powf(HdotN, nu);
    } else {
      norm = sqrtf((nu + 1.f) * (nv + 1.f)) / (8.f * PT_PI_F);
      if (HdotN < 1.f) lobe = powf(HdotN, (nu * HdotT * HdotT + nv * HdotB * HdotB) / (1.f - HdotN * HdotN));
      else lobe = 1.f;
    }
    float3 sc = pt_schlick3(Rs, HdotI);
    sc *= norm;
    sc *= lobe / (HdotI * fmaxf(NdotI, NdotO));
    float spdf = HdotI / (norm * lobe);
    pdf = (dpdf + spdf) / 2.f;
    return (dc / dpdf) + (sc / spdf);
  }
  if (MAT == PT_MAT_DIFFUSE_LIGHT) {  // diffuse_light.cuh:15-22
    pdf = 1.f;
    return s.color;
  }
  return f3(0.f);
}

// ---- rect light: programs/pdfs/rect_pdf.cu ---------------------------------------------------
PT_HD float3 rect_light_sample(const pt_light& l, float3 P, float ua, float ub) {
  float a = l.a0 + ua * (l.a1 - l.a0);
  float b = l.b0 + ub * (l.b1 - l.b0);
  float3 q = l.axis == PT_X_AXIS ? f3(l.k, a, b) : (l.axis == PT_Y_AXIS ? f3(a, l.k, b) : f3(a, b, l.k));
  return q - P;
}
PT_HD float rect_light_pdf(const pt_light& l, float3 P, float3 Wi) {
  float t, a, b, wn;
  if (l.axis == PT_X_AXIS) {
    t = (l.k - P.x) / Wi.x; a = P.y + t * Wi.y; b = P.z + t * Wi.z; wn = Wi.x;
  } else if (l.axis == PT_Y_AXIS) {
    t = (l.k - P.y) / Wi.y; a = P.x + t * Wi.x; b = P.z + t * Wi.z; wn = Wi.y;
  } else {
    t = (l.k - P.z) / Wi.z; a = P.x + t * Wi.x; b = P.y + t * Wi.y; wn = Wi.z;
  }
  if (a < l.a0 || a > l.a1 || b < l.b0 || b > l.b1) return 0.f;
  if (!(t < 3.402823466e+38f && t > 0.001f)) return 0.f;
  float len = length(Wi);
  float distance_squared = t * t * (len * len);
  // dot(Wi, axis normal) evaluated like the reference: x*nx + y*ny + z*nz with a unit axis
  float d = l.axis == PT_X_AXIS ? (Wi.x * 1.f + Wi.y * 0.f + Wi.z * 0.f)
            : (l.axis == PT_Y_AXIS ? (Wi.x * 0.f + Wi.y * 1.f + Wi.z * 0.f) : (Wi.x * 0.f + Wi.y * 0.f + Wi.z * 1.f));
  (void)wn;
  float cosine = fabsf(d / len);
  float area = (l.a1 - l.a0) * (l.b1 - l.b0);
  return distance_squared / (cosine * area);
}
// ---- sphere light: programs/pdfs/sphere_pdf.cu:8-84 (centre = (a0,a1,b0), radius = k) -------
PT_HD float3 sphere_light_sample(const pt_light& l, float3 P, float r1, float r2) {
  float3 center = f3(l.a0, l.a1, l.b0);
  float radius = l.k;
  float lc = length(center - P);
  float distance_squared = lc * lc;
  float z = 1.f + r2 * (sqrtf(1.f - radius * radius / distance_squared) - 1.f);
  float phi = 2.f * PT_PI_F * r1;
  float x = cosf(phi) * sqrtf(1.f - z * z);
  float y = sinf(phi) * sqrtf(1.f - z * z);
  float3 Wi = f3(x, y, z);
  // :79-80 as written: the frame is built around the sampled direction itself
  float3 n = normalize(Wi);
  float3 bn;
  if (fabsf(n.x) > fabsf(n.z)) bn = f3(-n.y, n.x, 0.f);
  else bn = f3(0.f, -n.z, n.y);
  bn = normalize(bn);
  float3 tg = cross(bn, n);
  return Wi.x * tg + Wi.y * bn + Wi.z * n;
}
PT_HD float sphere_light_pdf(const pt_light& l, float3 P, float3 Wi) {
  float3 center = f3(l.a0, l.a1, l.b0);
  float radius = l.k;
  const float3 oc = P - center;
  const float a = dot(Wi, Wi);
  const float b = dot(oc, Wi);
  const float c = dot(oc, oc) - radius * radius;
  const float discriminant = b * b - a * c;
  if (discriminant < 0.f) return 0.f;
  const float tmin = 0.001f, tmax = 3.402823466e+38f;
  float temp = (-b - sqrtf(discriminant)) / a;
  bool hit = temp < tmax && temp > tmin;
  if (!hit) {
    temp = (-b + sqrtf(discriminant)) / a;
    hit = temp < tmax && temp > tmin;
  }
  if (!hit) return 0.f;
  float lc = length(center - P);
  float distance_squared = lc * lc;
  float cos_theta_max = sqrtf(1.f - radius * radius / distance_squared);
  float solid_angle = 2.f * PT_PI_F * (1.f - cos_theta_max);
  return 1.f / solid_angle;
}
// LK: light kinds the kernel is specialised for (0 = rectangles only — the common case keeps the
// sphere code out of the hot shading kernels; 1 = any mix)
template <int LK>
PT_HD float3 light_sample(const pt_light& l, float3 P, float ua, float ub) {
  if (LK != 0 && l.type == PT_LIGHT_SPHERE) return sphere_light_sample(l, P, ua, ub);
  return rect_light_sample(l, P, ua, ub);
}
template <int LK>
PT_HD float light_pdf(const pt_light& l, float3 P, float3 Wi) {
  if (LK != 0 && l.type == PT_LIGHT_SPHERE) return sphere_light_pdf(l, P, Wi);
  return rect_light_pdf(l, P, Wi);
}
PT_HD float power_heuristic(float fPdf, float gPdf) {  // light_sample.cuh:27-33 (numf=numg=1)
  float f = 1u * fPdf, g = 1u * gPdf;
  return (f * f) / (f * f + g * g);
}

// ---- textures: programs/textures/*.cu ------------------------------------------------------
PT_HD float3 image_lookup(const ImageDev& im, float u, float v) {
  if (!(isfinite(u) && isfinite(v)) || im.width <= 0 || im.height <= 0) return f3(0.f);
  float x = u * (float)im.width - 0.5f;
  float y = v * (float)im.height - 0.5f;
  float fx0 = floorf(x), fy0 = floorf(y);
  float fx = x - fx0, fy = y - fy0;
  int x0 = (int)fmodf(fx0, (float)im.width);
  if (x0 < 0) x0 += im.width;
  if (x0 >= im.width) x0 -= im.width;
  int y0 = (int)fmodf(fy0, (float)im.height);
  if (y0 < 0) y0 += im.height;
  if (y0 >= im.height) y0 -= im.height;
  int x1 = x0 + 1 == im.width ? 0 : x0 + 1;
  int y1 = y0 + 1 == im.height ? 0 : y0 + 1;
  float3 c00 = xyz(im.texels[(size_t)y0 * im.width + x0]), c10 = xyz(im.texels[(size_t)y0 * im.width + x1]);
  float3 c01 = xyz(im.texels[(size_t)y1 * im.width + x0]), c11 = xyz(im.texels[(size_t)y1 * im.width + x1]);
  float3 top = c00 + fx * (c10 - c00);
  float3 bot = c01 + fx * (c11 - c01);
  return top + fy * (bot - top);
}
// Perlin noise: programs/textures/noise_texture.cu:12-60.  data: 768 floats ranvec, then
// perm_x, perm_y, perm_z (256 entries each, stored as floats)
PT_HD float perlin_noise(const float* data, float3 p) {
  float u = p.x - floorf(p.x), v = p.y - floorf(p.y), w = p.z - floorf(p.z);
  int i = (int)floorf(p.x), j = (int)floorf(p.y), k = (int)floorf(p.z);
  float uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w);
  float accum = 0;
  for (int di = 0; di < 2; di++)
    for (int dj = 0; dj < 2; dj++)
      for (int dk = 0; dk < 2; dk++) {
        int idx = (int)data[768 + ((i + di) & 255)] ^ (int)data[1024 + ((j + dj) & 255)] ^ (int)data[1280 + ((k + dk) & 255)];
        float3 c = f3(data[3 * idx], data[3 * idx + 1], data[3 * idx + 2]);
        float3 weight_v = f3(u - di, v - dj, w - dk);
        accum += (di * uu + (1 - di) * (1 - uu)) * (dj * vv + (1 - dj) * (1 - vv)) * (dk * ww + (1 - dk) * (1 - ww)) *
                 dot(c, weight_v);
      }
  return accum;
}
PT_HD float perlin_turb(const float* data, float3 p) {  // noise_texture.cu:46-58
  float accum = 0;
  float3 temp_p = p;
  float weight = 1.0f;
  for (int i = 0; i < 7; i++) {
    accum += weight * perlin_noise(data, temp_p);
    weight *= 0.5f;
    temp_p *= 2.f;
  }
  return fabsf(accum);
}

PT_HD float3 sample_texture(const SceneDev& sc, int tex, float u, float v, float3 p, int i) {
  for (int guard = 0; guard < 8; guard++) {
    if (tex < 0 || tex >= sc.n_textures) return f3(0.f);
    const pt_texture t = sc.textures[tex];
    if (t.type == PT_TEX_CONSTANT) return f3(t.color[0], t.color[1], t.color[2]);
    if (t.type == PT_TEX_CHECKER) {
      float sines = sinf(10 * p.x) * sinf(10 * p.y) * sinf(10 * p.z);
      tex = (sines < 0) ? t.a : t.b;
      i = 0;
    } else if (t.type == PT_TEX_IMAGE) {
      if (t.a < 0 || t.a >= sc.n_images) return f3(0.f);
      return image_lookup(sc.images[t.a], u, v);
    } else if (t.type == PT_TEX_VECTOR) {
      if (i >= t.b || i < 0) return f3(0.f);
      tex = sc.texvec[t.a + i];
      i = 0;
    } else if (t.type == PT_TEX_NOISE) {
      // noise_texture.cu:62-75: the axis switch falls through, so the Z expression always wins
      const float scale = t.color[0];
      float sinValue = sinf(scale * p.z + 5 * perlin_turb(sc.texdata + t.a, scale * p));
      return f3(1.f) * 0.5f * (1 + sinValue);
    } else if (t.type == PT_TEX_GRADIENT) {  // gradient_texture.cu:7-14
      const float* g = sc.texdata + t.a;
      const float3 ud = normalize(p);
      const float x = fabsf(ud.x), y = fabsf(ud.y), z = fabsf(ud.z);
      return x * f3(g[0], g[1], g[2]) + y * f3(g[3], g[4], g[5]) + z * f3(g[6], g[7], g[8]);
    } else {
      return f3(0.f);
    }
  }
  return f3(0.f);
}

// ---- hit record: programs/hitables/*.cu Get_HitRecord ----------------------------------------
struct HitRec {
  float3 P, Ng, Ns, Wo;
  float u, v;
  int index;     // texture index (material_buffer[prim] for meshes)
  int material;
};

PT_HD HitRec make_hit_record(const SceneDev& sc, float3 o, float3 d, float t, int leaf, float bu, float bv,
                             float time = 0.f) {
  const LeafRec L = sc.leaves[leaf];
  HitRec rec;
  rec.Wo = normalize(-d);
  rec.index = 0;
  rec.material = (int)(pt_f2u(L.c.w) & 0xffffffu);
  const uint32_t w = pt_f2u(L.b.w);
  const int kind = PT_LEAF_KIND(w);
  if (kind == PT_PRIM_TRIANGLE) {  // triangle.cu:85-149 (world-space mesh, identity transform)
    float3 a = xyz(L.a), b = xyz(L.b), c = xyz(L.c);
    float3 e1 = b - a, e2 = c - a;
    float b0 = 1.f - bu - bv, b1 = bu, b2 = bv;
    rec.P = a * b0 + b * b1 + c * b2;
    float3 Ng = cross(e1, e2);
    float area = length(Ng) / 2.f;
    Ng /= 2.f * area;
    Ng = normalize(normalize(Ng));
    rec.Ng = Ng;
    const int prim = (int)pt_f2u(L.a.w);
    const int ti = sc.prim_ref[prim].index;
    if (!sc.tri_n) {
      rec.Ns = Ng;
    } else {
      const float* n9 = sc.tri_n + 9 * (size_t)ti;
      float3 Ns = f3(n9[0], n9[1], n9[2]) * b0 + f3(n9[3], n9[4], n9[5]) * b1 + f3(n9[6], n9[7], n9[8]) * b2;
      rec.Ns = normalize(Ns);
    }
    if (!sc.tri_uv) {
      rec.u = rec.v = 0.f;
    } else {
      const float* t6 = sc.tri_uv + 6 * (size_t)ti;
      rec.u = t6[0] * b0 + t6[2] * b1 + t6[4] * b2;
      rec.v = t6[1] * b0 + t6[3] * b1 + t6[5] * b2;
    }
    rec.index = sc.tri_tex ? sc.tri_tex[ti] : 0;
    return rec;
  }
  const int xf = PT_LEAF_XFORM(w);
  const float* o2w = nullptr;
  const float* w2o = nullptr;
  if (xf >= 0) {  // record in object space, then to world (intended rtTransform* semantics)
    o2w = sc.xforms[xf].o2w;
    w2o = sc.xforms[xf].w2o;
    o = xf_point(w2o, o);
    d = xf_vector(w2o, d);
  }
  float3 hp = o + t * d;
  rec.P = o2w ? xf_point(o2w, hp) : hp;
  // texture coordinates of analytic primitives cost an atan2f + asinf (sphere) or two IEEE divisions (rect) per
  // hit and are read by image textures only: a scene without image textures skips them (same colours, exactly)
  const bool need_uv = sc.tex_need_uv != 0;
  float3 n;
  rec.u = rec.v = 0.f;
  if (kind == PT_PRIM_VOLUME_SPHERE || kind == PT_PRIM_VOLUME_BOX) {
    n = f3(1.f, 0.f, 0.f);  // volume_sphere.cu:92-99, volume_box.cu:67-74
  } else if (kind == PT_PRIM_SPHERE || kind == PT_PRIM_MOVING_SPHERE) {  // sphere.cu:65-93, moving_sphere.cu:82-100
    float3 ctr = xyz(L.a);
    if (kind == PT_PRIM_MOVING_SPHERE) {
      const float c0[3] = {L.a.x, L.a.y, L.a.z}, c1[3] = {L.c.x, L.c.y, L.c.z};
      ctr = moving_center(c0, c1, L.b.y, L.b.z, time);
    }
    float3 T = (hp - ctr) / L.b.x;
    n = T;
    if (need_uv) {  // sphere.cu:85-90; only image textures read (u, v): skipped when the scene has none
      float phi = atan2f(T.z, T.x);
      float theta = asinf(T.y);
      rec.u = 1.f - (phi + PT_PI_F) / (2.f * PT_PI_F);
      rec.v = (theta + PT_PI_F / 2.f) / PT_PI_F;
    }
  } else if (kind == PT_PRIM_RECT) {  // aarect.cu:71-115
    const float a0 = L.a.x, a1 = L.a.y, b0 = L.a.z, b1 = L.b.x;
    const uint32_t af = pt_f2u(L.b.z);
    const int axis = af & 3;
    float pa, pb;
    if (axis == PT_X_AXIS) {
      n = f3(1.f, 0.f, 0.f);
      pa = hp.y;
      pb = hp.z;
    } else if (axis == PT_Y_AXIS) {
      n = f3(0.f, 1.f, 0.f);
      pa = hp.x;
      pb = hp.z;
    } else {
      n = f3(0.f, 0.f, 1.f);
      pa = hp.x;
      pb = hp.y;
    }
    if (need_uv) {
      rec.u = (pa - a0) / (a1 - a0);
      rec.v = (pb - b0) / (b1 - b0);
    }
    if (af & 4u) n = -n;
  } else {  // box.cu:24-33,56-81
    float3 t0 = (xyz(L.a) - o) / d;
    float3 t1 = (xyz(L.b) - o) / d;
    float3 neg = f3(t == t0.x ? 1.f : 0.f, t == t0.y ? 1.f : 0.f, t == t0.z ? 1.f : 0.f);
    float3 pos = f3(t == t1.x ? 1.f : 0.f, t == t1.y ? 1.f : 0.f, t == t1.z ? 1.f : 0.f);
    n = pos - neg;
  }
  n = normalize(w2o ? xf_normal(w2o, n) : n);
  rec.Ng = rec.Ns = n;
  return rec;
}

// ---- camera: programs/raygen.cu:50-64,148-162 ----------------------------------------------
// camera ray for a seed that is already computed (the device hashes with pt_tea64_balanced, integrator.cu)
PT_HD void camera_ray_seeded(const pt_camera& c, int W, int H, int x, int y, uint32_t& seed, float3& o, float3& d);
PT_HD void camera_ray(const pt_camera& c, int W, int H, int x, int y, int frame, uint32_t& seed, float3& o,
                      float3& d) {
  seed = pt_tea64((uint32_t)W * (uint32_t)y + (uint32_t)x, (uint32_t)frame);
  camera_ray_seeded(c, W, H, x, y, seed, o, d);
}
PT_HD void camera_ray_seeded(const pt_camera& c, int W, int H, int x, int y, uint32_t& seed, float3& o, float3& d) {
  float u = float(x + pt_rnd(seed)) / W;
  float v = float(y + pt_rnd(seed)) / H;
  float3 rd = c.lens_radius * pt_random_in_unit_disk(seed);
  float3 cu = f3(c.u[0], c.u[1], c.u[2]), cv = f3(c.v[0], c.v[1], c.v[2]);
  float3 lens_offset = cu * rd.x + cv * rd.y;
  o = f3(c.origin[0], c.origin[1], c.origin[2]) + lens_offset;
  d = f3(c.lower_left_corner[0], c.lower_left_corner[1], c.lower_left_corner[2]) +
      u * f3(c.horizontal[0], c.horizontal[1], c.horizontal[2]) +
      v * f3(c.vertical[0], c.vertical[1], c.vertical[2]) - o;
}

// ---- miss: programs/miss.cu:30-95 ------------------------------------------------------------
PT_HD float3 miss_color(const SceneDev& sc, float3 d) {
  const pt_miss& ms = sc.miss;
  if (ms.type == PT_MISS_GRADIENT) {
    float3 ud = normalize(d);
    float t = 0.5f * (ud.y + 1.f);
    float3 c = (1.f - t) * f3(ms.color1[0], ms.color1[1], ms.color1[2]);
    c += t * f3(ms.color2[0], ms.color2[1], ms.color2[2]);
    return c;
  }
  if (ms.type == PT_MISS_CONSTANT) return f3(ms.color1[0], ms.color1[1], ms.color1[2]);
  const bool has = ms.image >= 0 && ms.image < sc.n_images;
  if (ms.type == PT_MISS_IMG) {
    float theta = atan2f(d.x, d.z);
    float phi = PT_OPTIX_PI * 0.5f - acosf(d.y);
    float u = (theta + PT_OPTIX_PI) * (0.5f * PT_OPTIX_1_PI);
    float v = 0.5f * (1.f + sinf(phi));
    return has ? image_lookup(sc.images[ms.image], u, v) : f3(0.f);
  }
  float u, v;
  if (ms.spherical) {
    float r = length(d);
    float lon = atan2f(d.z, d.x);
    float lat = acosf(d.y / r);
    u = lon * (1.f / (PT_PI_F * 2.f));
    v = lat * (1.f / PT_PI_F);
  } else {
    float theta = atan2f(d.x, d.z);
    theta = theta < 0.f ? theta + (2.f * PT_PI_F) : theta;
    float phi = acosf(d.y);
    u = 1.f - (theta / (2.f * PT_PI_F));
    v = phi / PT_PI_F;
  }
  float3 c = has ? image_lookup(sc.images[ms.image], u, v) : f3(0.f);
  return 2.f * c;
}

// ---- environment importance sampling (NEW: pt_miss.env_nee; not in the reference) -----------
// Active when the scene has no rect lights, the miss program is a spherical HDR map and the
// alias table was built (SceneDev::env_n > 0).  Lambertian / Oren-Nayar hits then sample one
// direction from the texel distribution (4 draws: texel, alias coin, 2 jitter), trace an
// any-hit shadow ray and add f * L * w / p with the power heuristic against the BSDF pdf; the
// BSDF-sampled bounce carries its pdf so that a later miss is weighted the other way round
// (and the reference's clamp at miss is dropped: it cannot be applied per estimator).
#ifdef PT_NO_ENV
PT_HD bool env_nee_active(const SceneDev&) { return false; }
#else
PT_HD bool env_nee_active(const SceneDev& sc) { return sc.env_n > 0 && sc.n_lights == 0; }
#endif
PT_HD float3 env_dir_from_uv(float u, float v) {  // inverse of miss.cu:70-79 (u = lon/2pi, v = lat/pi)
  float lon = u * (2.f * PT_PI_F), lat = v * PT_PI_F;
  float st = sinf(lat);
  return f3(st * cosf(lon), cosf(lat), st * sinf(lon));
}
PT_HD float env_pdf_dir(const SceneDev& sc, float3 d) {  // solid-angle pdf of direction d
  float r = length(d);
  float lon = atan2f(d.z, d.x);
  float lat = acosf(d.y / r);
  float u = lon * (1.f / (PT_PI_F * 2.f));
  if (u < 0.f) u += 1.f;
  float v = lat * (1.f / PT_PI_F);
  int x = (int)(u * (float)sc.env_w), y = (int)(v * (float)sc.env_h);
  x = x < 0 ? 0 : (x > sc.env_w - 1 ? sc.env_w - 1 : x);
  y = y < 0 ? 0 : (y > sc.env_h - 1 ? sc.env_h - 1 : y);
  float st = sinf(lat);
  if (!(st > 1e-6f)) return 0.f;
  return sc.env_texel_pdf[(size_t)y * sc.env_w + x] * ((float)sc.env_n / (2.f * PT_PI_F * PT_PI_F * st));
}
PT_HD float3 env_sample_dir(const SceneDev& sc, uint32_t& seed, float& pdf) {
  float u0 = pt_rnd(seed), u1 = pt_rnd(seed), u2 = pt_rnd(seed), u3 = pt_rnd(seed);
  int i = (int)(u0 * (float)sc.env_n);
  i = i > sc.env_n - 1 ? sc.env_n - 1 : i;
  if (u1 >= sc.env_alias_prob[i]) i = sc.env_alias_idx[i];
  int x = i % sc.env_w, y = i / sc.env_w;
  float u = ((float)x + u2) / (float)sc.env_w, v = ((float)y + u3) / (float)sc.env_h;
  float st = sinf(v * PT_PI_F);
  pdf = st > 1e-6f ? sc.env_texel_pdf[i] * ((float)sc.env_n / (2.f * PT_PI_F * PT_PI_F * st)) : 0.f;
  return env_dir_from_uv(u, v);
}

// ---- one shading event -------------------------------------------------------------------------
enum { SH_TERMINATE = 0, SH_BOUNCE = 1, SH_SHADOW = 2 };

struct PathIn {
  float3 o, d;
  float3 thr, rad;
  uint32_t seed;
  int depth;
  bool prev_specular;
  float last_pdf;  // env NEE: BSDF pdf of the direction that produced this ray (0 = unweighted)
  float time;      // prd.time (raygen.cu:69): shutter time of the sample, read by moving spheres
};

struct ShadeOut {
  int kind;
  float3 col;  // SH_TERMINATE: sample colour (before de_nan)
  // SH_BOUNCE, and candidate V of SH_SHADOW
  float3 P, dir, thr, rad;
  uint32_t seed;
  bool specular;
  // SH_SHADOW only: shadow-ray direction and candidate S (light occluded)
  float3 WiL, dirS, thrS, radS;
  uint32_t seedS;
  bool terminal;  // light-surface NEE (numLights >= 2): rad / radS are final colours
  float pdf;      // env NEE: pdf of the sampled bounce direction (carried to the next miss)
  bool env_shadow;  // shadow ray towards the environment: visible <=> nothing is hit
  uint32_t seedShadow;  // SH_SHADOW: prd.seed when the reference traces the shadow ray (volumes read it)
};

// Direct_Light<T> (light_sample.cuh:35-103) for a cosine-type surface, split
// around the shadow ray.  Returns false when no shadow ray is needed.
template <int MAT, int LK>
PT_HD bool nee_begin(const SceneDev& sc, const Surface& surf, float3 P, float3 Wo, float3 N, bool isLight,
                     uint32_t& seed, float3& WiL, float3& directV, uint32_t& seedV, float3& misDir) {
  directV = f3(0.f);
  const int numLights = sc.n_lights;
  if (numLights == 0) return false;
  int index = ((int)(pt_rnd(seed) * numLights)) % numLights;
  if (isLight && numLights == 1) return false;
  const pt_light L = sc.lights[index];
  float3 emission = f3(L.emission[0], L.emission[1], L.emission[2]);
  float ua = pt_rnd(seed);
  float ub = pt_rnd(seed);
  float3 Wi = light_sample<LK>(L, P, ua, ub);
  float lightPDF = light_pdf<LK>(L, P, Wi);
  if (dot(Wi, N) < 0.f) return false;
  // from here the reference traces the shadow ray; if occluded it returns with
  // `seed` as it is now (candidate S).  Candidate V continues:
  WiL = Wi;
  const bool emNull = emission.x == 0.f && emission.y == 0.f && emission.z == 0.f;
  float3 direct = f3(0.f);
  if (lightPDF != 0.f && !emNull) {
    float matPDF = 0.f;
    float3 matValue = bsdf_eval<MAT>(surf, Wo, Wi, N, matPDF);
    if (matPDF != 0.f && !(matValue.x == 0.f && matValue.y == 0.f && matValue.z == 0.f)) {
      float weight = power_heuristic(lightPDF, matPDF);
      direct += matValue * emission * weight / lightPDF;
    }
  }
  seedV = seed;
  float3 Wi2 = bsdf_sample<MAT>(surf, Wo, N, seedV);
  misDir = Wi2;
  float matPDF = 0.f;
  float3 matValue = bsdf_eval<MAT>(surf, Wo, Wi2, N, matPDF);
  if (matPDF != 0.f && !(matValue.x == 0.f && matValue.y == 0.f && matValue.z == 0.f)) {
    float lp = light_pdf<LK>(L, P, Wi2);
    if (!(!lp || emNull)) {
      float weight = power_heuristic(matPDF, lp);
      direct += matValue * emission * weight / matPDF;
    }
  }
  directV = direct;
  return true;
}

template <int MAT, int LK = 1>
PT_HD void shade_hit(const SceneDev& sc, const PathIn& in, float t, int leaf, float bu, float bv, ShadeOut& out) {
  HitRec rec = make_hit_record(sc, in.o, in.d, t, leaf, bu, bv, in.time);
  const pt_material m = sc.materials[rec.material];
  float3 P = rec.P, Wo = rec.Wo, N = rec.Ns;
  Surface surf;
  surf.p0 = m.p0;
  surf.p1 = m.p1;
  surf.color = surf.color2 = f3(0.f);
  out.P = P;
  out.terminal = false;
  out.pdf = 0.f;
  out.env_shadow = false;
  out.seedShadow = 0u;
  out.seed = in.seed;
  out.thr = in.thr;
  out.rad = in.rad;

  if (MAT == PT_MAT_NORMAL) {  // normal_shader.cu:22-35
    float3 n = (m.flags & 1) ? rec.Ns : rec.Ng;
    out.kind = SH_TERMINATE;
    out.col = n * 0.5f + f3(0.5f);
    return;
  }
  surf.color = sample_texture(sc, m.tex0, rec.u, rec.v, P, rec.index);

  if ((MAT == PT_MAT_LAMBERTIAN || MAT == PT_MAT_OREN_NAYAR) && env_nee_active(sc)) {
    uint32_t seed = in.seed;
    float lpdf;
    float3 WiL = env_sample_dir(sc, seed, lpdf);
    out.seedShadow = seed;
    float3 direct = f3(0.f);
    bool need_shadow = false;
    if (dot(WiL, N) > 0.f && lpdf > 0.f) {
      float mpdf = 0.f;
      float3 f = bsdf_eval<MAT>(surf, Wo, WiL, N, mpdf);
      if (mpdf != 0.f && !(f.x == 0.f && f.y == 0.f && f.z == 0.f)) {
        float w = power_heuristic(lpdf, mpdf);
        direct = f * miss_color(sc, WiL) * w / lpdf;
        need_shadow = true;
      }
    }
    float3 Wi = bsdf_sample<MAT>(surf, Wo, N, seed);
    float pdf = 0.f;
    float3 att = bsdf_eval<MAT>(surf, Wo, Wi, N, pdf);
    out.specular = false;
    out.dir = out.dirS = Wi;
    out.thr = out.thrS = in.thr * clamp3(att / pdf, 0.f, 1.f);
    out.seed = out.seedS = seed;
    out.pdf = pdf;
    out.radS = in.rad;
    out.rad = need_shadow ? in.rad + in.thr * direct : in.rad;
    out.WiL = WiL;
    out.env_shadow = need_shadow;
    out.kind = need_shadow ? SH_SHADOW : SH_BOUNCE;
    return;
  }
  if (MAT == PT_MAT_LAMBERTIAN || MAT == PT_MAT_OREN_NAYAR || MAT == PT_MAT_ISOTROPIC) {
    // lambertian.cu:49-73, oren_nayar.cu:46-70, isotropic.cu:34-60 (the latter without the clamp)
    uint32_t seed = in.seed;
    float3 WiL, directV, misDir;
    uint32_t seedV;
    bool need_shadow = nee_begin<MAT, LK>(sc, surf, P, Wo, N, false, seed, WiL, directV, seedV, misDir);
    out.seedShadow = seed;  // state after the light draws
    out.specular = false;
    if (!need_shadow) {
      float3 Wi = bsdf_sample<MAT>(surf, Wo, N, seed);
      float pdf = 0.f;
      float3 att = bsdf_eval<MAT>(surf, Wo, Wi, N, pdf);
      out.kind = SH_BOUNCE;
      out.dir = Wi;
      out.thr = in.thr * (MAT == PT_MAT_ISOTROPIC ? att / pdf : clamp3(att / pdf, 0.f, 1.f));
      out.seed = seed;
      return;
    }
    // candidate S: occluded -> bounce direction from the next two draws == misDir
    {
      float pdf = 0.f;
      float3 att = bsdf_eval<MAT>(surf, Wo, misDir, N, pdf);
      out.dirS = misDir;
      out.thrS = in.thr * (MAT == PT_MAT_ISOTROPIC ? att / pdf : clamp3(att / pdf, 0.f, 1.f));
      out.radS = in.rad;
      out.seedS = seedV;  // two draws after `seed`
    }
    // candidate V: visible
    {
      float3 Wi = bsdf_sample<MAT>(surf, Wo, N, seedV);
      float pdf = 0.f;
      float3 att = bsdf_eval<MAT>(surf, Wo, Wi, N, pdf);
      out.dir = Wi;
      out.thr = in.thr * (MAT == PT_MAT_ISOTROPIC ? att / pdf : clamp3(att / pdf, 0.f, 1.f));
      out.rad = in.rad + in.thr * directV;
      out.seed = seedV;
    }
    out.WiL = WiL;
    // any_hit (hit.cu:16-24) additionally wants dot(N, Wi) > 0
    if (!(dot(N, WiL) > 0.f)) {
      out.kind = SH_BOUNCE;
      out.dir = out.dirS;
      out.thr = out.thrS;
      out.rad = out.radS;
      out.seed = out.seedS;
      return;
    }
    out.kind = SH_SHADOW;
    return;
  }
  if (MAT == PT_MAT_METAL) {  // metal.cu:38-56
    uint32_t seed = in.seed;
    float3 reflected = reflect3(-Wo, N);
    out.dir = reflected + m.p0 * pt_random_in_unit_sphere(seed);
    out.kind = SH_BOUNCE;
    out.thr = in.thr * surf.color;
    out.seed = seed;
    out.specular = true;
    return;
  }
  if (MAT == PT_MAT_DIELECTRIC) {  // dielectric.cu:46-96
    uint32_t seed = in.seed;
    float3 Wd = -rec.Wo;
    float ref_idx = m.p0;
    float ni_over_nt;
    float cosine = dot(Wd, N);
    if (cosine > 0.f) {
      N = -N;
      ni_over_nt = ref_idx;
      cosine = ref_idx * cosine / length(Wd);
    } else {
      ni_over_nt = 1.f / ref_idx;
      cosine = -cosine / length(Wd);
    }
    float3 refracted = f3(0.f);
    float reflect_prob;
    if (pt_refract(Wd, N, ni_over_nt, refracted)) reflect_prob = pt_schlick(cosine, ref_idx);
    else reflect_prob = 1.f;
    if (pt_rnd(seed) < reflect_prob) out.dir = reflect3(Wd, N);
    else out.dir = normalize(refracted);
    out.kind = SH_BOUNCE;
    out.thr = in.thr * (surf.color * f3(1.f));
    out.seed = seed;
    out.specular = true;
    return;
  }
  if (MAT == PT_MAT_DIFFUSE_LIGHT) {  // diffuse_light.cu:30-48 + raygen.cu:85-90
    uint32_t seed = in.seed;
    float3 WiL, directV, misDir;
    uint32_t seedV;
    bool need_shadow = nee_begin<MAT, LK>(sc, surf, P, Wo, N, true, seed, WiL, directV, seedV, misDir);
    out.seedShadow = seed;
    float3 thr = in.thr;
    if (dot(N, Wo) < 0.f) thr *= surf.color;
    const bool add = in.depth == 0 || in.prev_specular;
    float3 radS = in.rad;
    if (add) radS += thr;
    if (!need_shadow || !(dot(N, WiL) > 0.f)) {
      out.kind = SH_TERMINATE;
      out.col = radS;
      return;
    }
    float3 radV = in.rad + in.thr * directV;
    if (add) radV += thr;
    out.kind = SH_SHADOW;
    out.terminal = true;
    out.WiL = WiL;
    out.rad = radV;
    out.radS = radS;
    out.dir = out.dirS = f3(0.f, 1.f, 0.f);
    out.thrS = out.thr;
    out.seedS = out.seed;
    return;
  }
  if (MAT == PT_MAT_TORRANCE_SPARROW || MAT == PT_MAT_ASHIKHMIN_SHIRLEY) {
    uint32_t seed = in.seed;  // torrance_sparrow.cu:39-59, ashikhmin_shirley.cu:48-68
    if (MAT == PT_MAT_ASHIKHMIN_SHIRLEY) surf.color2 = sample_texture(sc, m.tex1, rec.u, rec.v, P, rec.index);
    float3 Wi = bsdf_sample<MAT>(surf, Wo, N, seed);
    float pdf = 0.f;
    float3 att = bsdf_eval<MAT>(surf, Wo, Wi, N, pdf);
    out.kind = SH_BOUNCE;
    out.dir = Wi;
    if (MAT == PT_MAT_TORRANCE_SPARROW) out.thr = in.thr * clamp3(att / pdf, 0.f, 1.f);
    else out.thr = in.thr * clamp3(att, 0.f, 1.f);
    out.seed = seed;
    out.specular = true;
    return;
  }
  out.kind = SH_TERMINATE;
  out.col = in.rad;
}

// raygen.cu:81-82: miss returns radiance + clamp(throughput * background, 0, 1)
PT_HD float3 shade_miss(const SceneDev& sc, const PathIn& in) {
  float3 thr = in.thr * miss_color(sc, in.d);
  if (env_nee_active(sc)) {  // NEW estimator: MIS weight against the env-sampling pdf, no clamp
    float w = 1.f;
    if (in.last_pdf > 0.f) w = power_heuristic(in.last_pdf, env_pdf_dir(sc, in.d));
    return in.rad + thr * w;
  }
  return in.rad + clamp3(thr, 0.f, 1.f);
}

PT_HD float3 pt_de_nan(float3 c) {  // raygen.cu:126-134
  if (!(c.x == c.x)) c.x = 0.f;
  if (!(c.y == c.y)) c.y = 0.f;
  if (!(c.z == c.z)) c.z = 0.f;
  return c;
}

// runtime dispatch on the material type (used where a warp may hold mixed materials)
PT_HD void shade_hit_any(const SceneDev& sc, int mat_type, const PathIn& in, float t, int leaf, float bu, float bv,
                         ShadeOut& out) {
  switch (mat_type) {
    case PT_MAT_LAMBERTIAN: shade_hit<PT_MAT_LAMBERTIAN>(sc, in, t, leaf, bu, bv, out); break;
    case PT_MAT_METAL: shade_hit<PT_MAT_METAL>(sc, in, t, leaf, bu, bv, out); break;
    case PT_MAT_DIELECTRIC: shade_hit<PT_MAT_DIELECTRIC>(sc, in, t, leaf, bu, bv, out); break;
    case PT_MAT_DIFFUSE_LIGHT: shade_hit<PT_MAT_DIFFUSE_LIGHT>(sc, in, t, leaf, bu, bv, out); break;
    case PT_MAT_OREN_NAYAR: shade_hit<PT_MAT_OREN_NAYAR>(sc, in, t, leaf, bu, bv, out); break;
    case PT_MAT_TORRANCE_SPARROW: shade_hit<PT_MAT_TORRANCE_SPARROW>(sc, in, t, leaf, bu, bv, out); break;
    case PT_MAT_ASHIKHMIN_SHIRLEY: shade_hit<PT_MAT_ASHIKHMIN_SHIRLEY>(sc, in, t, leaf, bu, bv, out); break;
    case PT_MAT_ISOTROPIC: shade_hit<PT_MAT_ISOTROPIC>(sc, in, t, leaf, bu, bv, out); break;
    default: shade_hit<PT_MAT_NORMAL>(sc, in, t, leaf, bu, bv, out); break;
  }
}
