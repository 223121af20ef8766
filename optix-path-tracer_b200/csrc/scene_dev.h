// scene_dev.h — device-side scene layout (what lives in HBM after commit).
//
//  nodes8   : compressed 8-wide BVH, 80 B / node = 5 x 128-bit words
//             (Ylitie, Karras, Laine 2017 layout):
//               w0 = { p.x, p.y, p.z, ex | ey<<8 | ez<<16 | imask<<24 }
//               w1 = { child_base, leaf_base, meta[0..3], meta[4..7] }
//               w2 = { qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7] }
//               w3 = { qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7] }
//               w4 = { qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7] }
//             child box = p + q * 2^(e-127) per axis (e stored as a biased
//             float exponent byte); meta: inner child = 0b001sssss with
//             sssss = 24 + slot, leaf child = unary count (1..3) << 5 | offset.
//  leaves   : 48 B records (3 x float4), leaf-contiguous, one per primitive:
//               a = { v0.xyz | centre | (a0,a1,b0) | bmin, prim id }
//               b = { v1.xyz | (r,0,0)  | (b1,k,axis+4*flip) | bmax, kind | (xform+1)<<8 }
//               c = { v2.xyz | unused                           , material id | type<<24 }
//             moving sphere: a = centre0, b = (r, time0, time1), c = centre1
//             volume sphere: a = centre,  b = (r, density, 0);  volume box: a = bmin, b = bmax, c.x = density
//  prim_ref : pt_prim_ref per primitive id (typed index for normals/uvs)
#pragma once
#include "../../include/ptb200.h"
#include "hd.h"

struct LeafRec {
  float4 a, b, c;
};

struct Bvh2Node {  // 32 B; children: >=0 internal node index, <0 leaf ~prim slot
  float lo[3];
  int left;
  float hi[3];
  int right;
};

struct ImageDev {
  const float4* texels;
  int width, height;
};

// everything kernels need, passed by value (fits in constant bank)
struct SceneDev {
  const uint4* nodes8;
  const LeafRec* leaves;
  const pt_prim_ref* prim_ref;
  const int* prim_leaf;  // primitive id -> leaf record index
  const pt_xform* xforms;
  const float* tri_n;    // 9 floats / triangle or null
  const float* tri_uv;   // 6 floats / triangle or null
  const int* tri_tex;    // per-triangle texture index or null
  const pt_material* materials;
  const pt_texture* textures;
  const int* texvec;
  const float* texdata;
  const ImageDev* images;
  const pt_light* lights;
  // environment importance sampling (alias table over the texels of the HDR miss image)
  const float* env_alias_prob;
  const int* env_alias_idx;
  const float* env_texel_pdf;
  int env_n, env_w, env_h;
  int tex_need_uv;   // some texture is an image: hit records must carry (u, v)
  uint32_t qbias;    // 0x47000000, deliberately a run-time value: see pt_byte_to_float (traverse.h)
  int special_mask;  // bit0: moving spheres present (rays carry prd.time), bit1: volumes present (rays carry a seed)
  int n_prims, n_materials, n_textures, n_texvec, n_texdata, n_images, n_lights;
  pt_camera camera;
  pt_miss miss;
};

#define PT_LEAF_KIND(w) ((w) & 0xff)
#define PT_LEAF_XFORM(w) (((int)((w) >> 8)) - 1)
