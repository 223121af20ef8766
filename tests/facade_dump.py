"""Helper run in a subprocess: build scene `id` through a facade library (mine, or the
reference's scenes.hpp compiled against my headers) and pickle its flat description."""
import ctypes as C
import pickle
import sys

import numpy as np

import helpers as h

pt = h.pt


def desc_arrays(d):
    def arr(ptr, n, dtype, width=1):
        if not ptr or n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * width * np.dtype(dtype).itemsize,)) \
            .view(dtype).copy()
    out = {
        "prims": arr(d.prims, d.n_prims, np.int32, 4), "spheres": arr(d.spheres, d.n_spheres, np.float32, 4),
        "rects": arr(d.rects, d.n_rects, np.uint32, 7), "boxes": arr(d.boxes, d.n_boxes, np.float32, 6),
        "tri_v": arr(d.tri_v, d.n_tris, np.float32, 9), "tri_n": arr(d.tri_n, d.n_tris, np.float32, 9),
        "tri_uv": arr(d.tri_uv, d.n_tris, np.float32, 6), "tri_tex": arr(d.tri_tex, d.n_tris, np.int32),
        "xforms": arr(d.xforms, d.n_xforms, np.float32, 24),
        "materials": arr(d.materials, d.n_materials, np.uint32, 6), "textures": arr(d.textures, d.n_textures, np.uint32, 6),
        "texvec": arr(d.texvec, d.n_texvec, np.int32), "lights": arr(d.lights, d.n_lights, np.uint32, 10),
        "camera": np.frombuffer(bytes(d.camera), dtype=np.float32).copy(),
        "miss": np.frombuffer(bytes(d.miss), dtype=np.uint32).copy(),
        "n_images": d.n_images,
        "moving_spheres": arr(d.moving_spheres, d.n_moving_spheres, np.float32, 9),
        "volume_spheres": arr(d.volume_spheres, d.n_volume_spheres, np.float32, 5),
        "volume_boxes": arr(d.volume_boxes, d.n_volume_boxes, np.float32, 7),
    }
    return out


def build(so, scene_id, W, H, model):
    lib = pt.load()
    host = C.CDLL(so)
    host.pth_build_scene.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    hnd = C.c_void_p()
    rc = host.pth_build_scene(scene_id, W, H, model, C.byref(hnd))
    if rc != 0:
        raise SystemExit(f"pth_build_scene failed: {rc}")
    s = pt.Scene.__new__(pt.Scene)
    s.lib, s.h, s._keep = lib, hnd, []
    return s


if __name__ == "__main__":
    so, scene_id, W, H, model, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    s = build(so, scene_id, W, H, model)
    with open(out, "wb") as fh:
        pickle.dump(desc_arrays(s.desc()), fh)
