// mesh.hpp — OBJ mesh hitable (reference surface: host_includes/mesh.hpp:19-345).
// Faces are flattened to unindexed per-corner arrays exactly like the reference
// (mesh.hpp:105-149); without a given material, MTL materials become a
// Vector_Texture of Lambertian albedos indexed per triangle (mesh.hpp:74-102).
#pragma once
#include <map>

#include "hitables.hpp"
#include "io_formats.hpp"

class Mesh {
 public:
  Mesh(std::string fileName, bool RTX) : RTX_MODE(RTX), givenMaterial(nullptr), fileName(fileName), assetsFolder("") {}
  Mesh(std::string fileName, std::string assetsFolder, bool RTX)
      : RTX_MODE(RTX), givenMaterial(nullptr), fileName(fileName), assetsFolder(assetsFolder) {}
  Mesh(std::string fileName, BRDF* givenMaterial, bool RTX)
      : RTX_MODE(RTX), givenMaterial(givenMaterial), fileName(fileName), assetsFolder("") {}
  Mesh(std::string fileName, std::string assetsFolder, BRDF* givenMaterial, bool RTX)
      : RTX_MODE(RTX), givenMaterial(givenMaterial), fileName(fileName), assetsFolder(assetsFolder) {}

  void emit(Context& g_context, const Matrix4x4& o2w) {
    ptio::ObjMesh obj;
    std::string err;
    if (!ptio::load_obj(assetsFolder + fileName, assetsFolder, obj, err)) {
      std::cerr << "ERR: " << err << std::endl;
      throw std::string("Failed to load/parse .obj.");
    }
    BRDF* host_material;
    std::vector<int> remap(obj.materials.size(), 0);
    if (givenMaterial == nullptr) {
      if (obj.materials.empty()) throw std::string("Mesh without MTL needs a given material (mesh.hpp:142 hazard)");
      Texture_List textures;
      for (size_t m = 0; m < obj.materials.size(); m++) {
        const ptio::ObjMaterial& mp = obj.materials[m];
        Texture* tex;
        if (mp.diffuse_texname.length() > 0) tex = new Image_Texture(assetsFolder + mp.diffuse_texname);
        else tex = new Constant_Texture(mp.ambient[0], mp.ambient[1], mp.ambient[2]);
        remap[m] = textures.push(tex);
      }
      host_material = new Lambertian(new Vector_Texture(textures.texList));
    } else {
      host_material = givenMaterial;
    }
    g_context->batches.emplace_back();
    ContextObj::Batch& b = g_context->batches.back();
    b.kind = PT_PRIM_TRIANGLE;
    b.material = host_material->assignTo(g_context).id;
    toRowMajor3x4(o2w, b.xf);
    const size_t nt = obj.vi.size() / 3;
    bool hasN = !obj.ni.empty(), hasT = !obj.ti.empty();
    const int n_normals = (int)obj.normals.size() / 3, n_texcoords = (int)obj.texcoords.size() / 2;
    for (int i : obj.ni) hasN = hasN && i >= 0 && i < n_normals;
    for (int i : obj.ti) hasT = hasT && i >= 0 && i < n_texcoords;
    b.v.reserve(nt * 9);
    for (size_t t = 0; t < nt; t++) {
      for (int k = 0; k < 3; k++) {
        const float* p = &obj.vertices[3 * (size_t)obj.vi[3 * t + k]];
        b.v.insert(b.v.end(), p, p + 3);
        if (hasN) {
          const float* n = &obj.normals[3 * (size_t)obj.ni[3 * t + k]];
          b.n.insert(b.n.end(), n, n + 3);
        }
        if (hasT) {
          const float* u = &obj.texcoords[2 * (size_t)obj.ti[3 * t + k]];
          b.uv.insert(b.uv.end(), u, u + 2);
        }
      }
      if (givenMaterial == nullptr) {
        int mid = obj.material_ids[t];
        if (mid < 0) throw std::string("face without usemtl and no given material");
        b.tex.push_back(remap[mid]);
      }
    }
    g_context->dirty = true;
  }

  void rotate(float angle, AXIS axis) {
    arr.push_back(TransformParameter(Rotate_Transform, angle, axis, make_float3(0.f), make_float3(0.f)));
  }
  void scale(float3 scale) { arr.push_back(TransformParameter(Scale_Transform, 0.f, X_AXIS, scale, make_float3(0.f))); }
  void translate(float3 pos) {
    arr.push_back(TransformParameter(Translate_Transform, 0.f, X_AXIS, make_float3(0.f), pos));
  }
  // mesh.hpp:237-242: reversed, so scale -> rotate -> translate reads naturally
  void addTo(Group& /*d_world*/, Context& g_context) {
    std::reverse(arr.begin(), arr.end());
    emit(g_context, composeTransforms(arr));
  }

 private:
  bool RTX_MODE;  // no RT cores on B200: both modes use the same software traversal
  BRDF* givenMaterial;
  const std::string fileName, assetsFolder;
  std::vector<TransformParameter> arr;
};

class Mesh_List {
 public:
  Mesh_List() {}
  int push(Mesh* mesh) {
    list.push_back(mesh);
    return (int)list.size() - 1;
  }
  void rotate(float angle, AXIS axis) {
    arr.push_back(TransformParameter(Rotate_Transform, angle, axis, make_float3(0.f), make_float3(0.f)));
  }
  void scale(float3 scale) { arr.push_back(TransformParameter(Scale_Transform, 0.f, X_AXIS, scale, make_float3(0.f))); }
  void translate(float3 pos) {
    arr.push_back(TransformParameter(Translate_Transform, 0.f, X_AXIS, make_float3(0.f), pos));
  }
  void addListTo(Group& /*d_world*/, Context& g_context) {  // "FIXME: not working" in the reference
    Matrix4x4 m = composeTransforms(arr);
    for (Mesh* x : list) x->emit(g_context, m);
  }
  void addElementsTo(Group& d_world, Context& g_context) {
    for (Mesh* x : list) x->addTo(d_world, g_context);
  }

 private:
  std::vector<Mesh*> list;
  std::vector<TransformParameter> arr;
};
