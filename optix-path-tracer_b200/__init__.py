"""optix-path-tracer_b200 — Python host binding over the C ABI (include/ptb200.h).

This module is plumbing: ctypes mirrors of the POD types, a loader for the
in-tree CUDA library (csrc/libptb200.so) and thin `Scene` / `Renderer`
wrappers whose methods map 1:1 onto the C entry points.  The compute path is
the CUDA library; there is no CPU fallback — `load()` raises if the library is
missing and every compute call raises `PtError` if CUDA is unavailable.

The directory name contains '-', so import it with
`importlib` (see `__graft_entry__.load_package()`), module name
`optix_path_tracer_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libptb200.so")
HOST_LIB_PATH = os.path.join(_HERE, "csrc", "libptb200_host.so")

# ---- enums (include/ptb200.h) -------------------------------------------------
X_AXIS, Y_AXIS, Z_AXIS = 0, 1, 2
PRIM_TRIANGLE, PRIM_SPHERE, PRIM_RECT, PRIM_BOX = 0, 1, 2, 3
PRIM_MOVING_SPHERE, PRIM_VOLUME_SPHERE, PRIM_VOLUME_BOX = 4, 5, 6
(MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT, MAT_OREN_NAYAR,
 MAT_TORRANCE_SPARROW, MAT_ASHIKHMIN_SHIRLEY, MAT_NORMAL, MAT_ISOTROPIC) = range(9)
TEX_CONSTANT, TEX_CHECKER, TEX_IMAGE, TEX_VECTOR, TEX_NOISE, TEX_GRADIENT = 0, 1, 2, 3, 4, 5
MISS_GRADIENT, MISS_CONSTANT, MISS_IMG, MISS_HDR = 0, 1, 2, 3
LIGHT_RECT, LIGHT_SPHERE = 0, 1
FLAG_COUNT_FETCHES, FLAG_TIME_KERNELS = 1, 2

F3 = C.c_float * 3


class Sphere(C.Structure):
    _fields_ = [("center", F3), ("radius", C.c_float)]


class Rect(C.Structure):
    _fields_ = [("a0", C.c_float), ("a1", C.c_float), ("b0", C.c_float), ("b1", C.c_float),
                ("k", C.c_float), ("axis", C.c_int32), ("flip", C.c_int32)]


class Box(C.Structure):
    _fields_ = [("bmin", F3), ("bmax", F3)]


class MovingSphere(C.Structure):
    _fields_ = [("center0", F3), ("center1", F3), ("radius", C.c_float),
                ("time0", C.c_float), ("time1", C.c_float)]


class VolumeSphere(C.Structure):
    _fields_ = [("center", F3), ("radius", C.c_float), ("density", C.c_float)]


class VolumeBox(C.Structure):
    _fields_ = [("bmin", F3), ("bmax", F3), ("density", C.c_float)]


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("tex0", C.c_int32), ("tex1", C.c_int32),
                ("p0", C.c_float), ("p1", C.c_float), ("flags", C.c_int32)]


class Texture(C.Structure):
    _fields_ = [("type", C.c_int32), ("color", F3), ("a", C.c_int32), ("b", C.c_int32)]


class Light(C.Structure):
    _fields_ = [("type", C.c_int32), ("axis", C.c_int32), ("a0", C.c_float), ("a1", C.c_float),
                ("b0", C.c_float), ("b1", C.c_float), ("k", C.c_float), ("emission", F3)]


class Camera(C.Structure):
    _fields_ = [("origin", F3), ("lower_left_corner", F3), ("horizontal", F3), ("vertical", F3),
                ("u", F3), ("v", F3), ("w", F3), ("lens_radius", C.c_float),
                ("time0", C.c_float), ("time1", C.c_float)]


class Miss(C.Structure):
    _fields_ = [("type", C.c_int32), ("color1", F3), ("color2", F3), ("image", C.c_int32),
                ("spherical", C.c_int32), ("env_nee", C.c_int32)]


class Ray(C.Structure):
    _fields_ = [("origin", F3), ("tmin", C.c_float), ("dir", F3), ("tmax", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("t", C.c_float), ("prim", C.c_int32), ("u", C.c_float), ("v", C.c_float)]


class PrimRef(C.Structure):
    _fields_ = [("type", C.c_int32), ("index", C.c_int32), ("material", C.c_int32),
                ("xform", C.c_int32)]


class Xform(C.Structure):
    _fields_ = [("o2w", C.c_float * 12), ("w2o", C.c_float * 12)]


class Image(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgba", C.POINTER(C.c_float))]


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_prims", C.c_int32), ("prims", C.POINTER(PrimRef)),
        ("n_spheres", C.c_int32), ("spheres", C.POINTER(Sphere)),
        ("n_rects", C.c_int32), ("rects", C.POINTER(Rect)),
        ("n_boxes", C.c_int32), ("boxes", C.POINTER(Box)),
        ("n_tris", C.c_int32), ("tri_v", C.POINTER(C.c_float)), ("tri_n", C.POINTER(C.c_float)),
        ("tri_uv", C.POINTER(C.c_float)), ("tri_tex", C.POINTER(C.c_int32)),
        ("n_xforms", C.c_int32), ("xforms", C.POINTER(Xform)),
        ("n_materials", C.c_int32), ("materials", C.POINTER(Material)),
        ("n_textures", C.c_int32), ("textures", C.POINTER(Texture)),
        ("n_texvec", C.c_int32), ("texvec", C.POINTER(C.c_int32)),
        ("n_texdata", C.c_int32), ("texdata", C.POINTER(C.c_float)),
        ("n_images", C.c_int32), ("images", C.POINTER(Image)),
        ("n_lights", C.c_int32), ("lights", C.POINTER(Light)),
        ("camera", Camera), ("miss", Miss),
        ("env_n", C.c_int32), ("env_alias_prob", C.POINTER(C.c_float)),
        ("env_alias_idx", C.POINTER(C.c_int32)), ("env_texel_pdf", C.POINTER(C.c_float)),
        ("n_moving_spheres", C.c_int32), ("moving_spheres", C.POINTER(MovingSphere)),
        ("n_volume_spheres", C.c_int32), ("volume_spheres", C.POINTER(VolumeSphere)),
        ("n_volume_boxes", C.c_int32), ("volume_boxes", C.POINTER(VolumeBox)),
    ]


class BvhInfo(C.Structure):
    _fields_ = [("n_prims", C.c_int32), ("n_bvh2_nodes", C.c_int32), ("n_bvh8_nodes", C.c_int32),
                ("n_leaf_records", C.c_int32), ("sah_cost_lbvh", C.c_float),
                ("sah_cost_refined", C.c_float), ("build_ms", C.c_float),
                ("node_bytes", C.c_int64), ("leaf_bytes", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("nodes_fetched", C.c_uint64), ("leaves_fetched", C.c_uint64),
                ("iterations", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_total", C.c_float), ("ms_trace_closest", C.c_float),
                ("ms_trace_shadow", C.c_float), ("ms_shade", C.c_float),
                ("ms_generate", C.c_float), ("ms_reduce", C.c_float), ("drained_paths", C.c_uint32)]


RAY_DTYPE = np.dtype([("origin", np.float32, 3), ("tmin", np.float32),
                      ("dir", np.float32, 3), ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("prim", np.int32), ("u", np.float32), ("v", np.float32)])

# every symbol include/ptb200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "pt_last_error", "pt_abi_version", "pt_scene_create", "pt_scene_destroy",
    "pt_scene_add_spheres", "pt_scene_add_rects", "pt_scene_add_boxes", "pt_scene_add_mesh",
    "pt_scene_add_moving_spheres", "pt_scene_add_volume_spheres", "pt_scene_add_volume_boxes",
    "pt_scene_set_materials", "pt_scene_set_textures", "pt_scene_set_texture_vector",
    "pt_scene_set_texture_data",
    "pt_scene_add_image", "pt_scene_set_lights", "pt_scene_set_camera", "pt_scene_set_miss",
    "pt_scene_commit", "pt_scene_get_desc", "pt_scene_bvh_info", "pt_trace_rays",
    "pt_trace_rays_device", "pt_renderer_create", "pt_renderer_destroy", "pt_render",
    "pt_renderer_sync", "pt_read_accum", "pt_accum_device_ptr", "pt_get_stats",
    "pt_renderer_set_flags", "pt_renderer_set_pool", "pt_read_display", "pt_write_accum",
    # ABI 3
    "pt_release_caches", "pt_renderer_create_multi", "pt_split_frames", "pt_comm_unique_id", "pt_comm_create",
    "pt_comm_destroy", "pt_comm_reduce", "pt_split_frames_weighted",
]
MAX_PIXELS, BAND_PIXELS, MAX_DEPTH, COMM_ID_BYTES = 1 << 28, 1 << 23, 255, 128


class PtError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """Load csrc/libptb200.so.  Fails loudly when the CUDA library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("PTB200_LIB") or LIB_PATH   # PTB200_LIB: dev override (kernel variants)
    if not os.path.exists(p):
        raise PtError(f"CUDA library not built: {p} (run `python -c 'import __graft_entry__ as g; "
                      f"g.build()'` or `make -C optix-path-tracer_b200/csrc`)")
    lib = C.CDLL(p)
    lib.pt_last_error.restype = C.c_char_p
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    P = C.POINTER
    sig = {
        "pt_scene_create": [P(vp)],
        "pt_scene_destroy": [vp],
        "pt_scene_add_spheres": [vp, vp, vp, vp, i32, P(i32)],
        "pt_scene_add_rects": [vp, vp, vp, vp, i32, P(i32)],
        "pt_scene_add_boxes": [vp, vp, vp, vp, i32, P(i32)],
        "pt_scene_add_moving_spheres": [vp, vp, vp, vp, i32, P(i32)],
        "pt_scene_add_volume_spheres": [vp, vp, vp, vp, i32, P(i32)],
        "pt_scene_add_volume_boxes": [vp, vp, vp, vp, i32, P(i32)],
        "pt_scene_add_mesh": [vp, vp, vp, vp, vp, i32, vp, i32, P(i32)],
        "pt_scene_set_materials": [vp, vp, i32],
        "pt_scene_set_textures": [vp, vp, i32],
        "pt_scene_set_texture_vector": [vp, vp, i32],
        "pt_scene_set_texture_data": [vp, vp, i32],
        "pt_scene_add_image": [vp, vp, i32, i32, P(i32)],
        "pt_scene_set_lights": [vp, vp, i32],
        "pt_scene_set_camera": [vp, P(Camera)],
        "pt_scene_set_miss": [vp, P(Miss)],
        "pt_scene_commit": [vp],
        "pt_scene_get_desc": [vp, P(SceneDesc)],
        "pt_scene_bvh_info": [vp, P(BvhInfo)],
        "pt_trace_rays": [vp, vp, i64, vp, vp],
        "pt_trace_rays_device": [vp, vp, i64, vp, vp],
        "pt_renderer_create": [vp, i32, i32, vp, P(vp)],
        "pt_renderer_destroy": [vp],
        "pt_render": [vp, i32, i32, i32, i32, i32],
        "pt_renderer_sync": [vp],
        "pt_read_accum": [vp, vp],
        "pt_accum_device_ptr": [vp, P(vp)],
        "pt_read_display": [vp, vp],
        "pt_write_accum": [vp, vp],
        "pt_get_stats": [vp, P(Stats)],
        "pt_renderer_set_flags": [vp, i32],
        "pt_renderer_set_pool": [vp, i64],
        "pt_release_caches": [],
        "pt_renderer_create_multi": [vp, i32, i32, P(i32), i32, P(vp)],
        "pt_split_frames": [i32, i32, i32, i32, P(i32), P(i32)],
        "pt_split_frames_weighted": [i32, i32, i32, vp, i32, P(i32), P(i32)],
        "pt_comm_unique_id": [vp],
        "pt_comm_create": [vp, i32, i32, P(vp)],
        "pt_comm_destroy": [vp],
        "pt_comm_reduce": [vp, vp, i32],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        msg = lib.pt_last_error()
        raise PtError(f"ptb200 error {rc}: {msg.decode() if msg else '?'}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _struct_array(cls, items):
    arr = (cls * max(1, len(items)))()
    for i, it in enumerate(items):
        arr[i] = it
    return arr


class Scene:
    """pt_scene wrapper.  Methods map 1:1 onto the C ABI."""

    def __init__(self, lib: Optional[C.CDLL] = None):
        self.lib = lib or load()
        self.h = C.c_void_p()
        _check(self.lib, self.lib.pt_scene_create(C.byref(self.h)))
        self._keep = []

    def close(self):
        if self.h:
            self.lib.pt_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _add(self, fn, cls, prims, xforms, materials):
        n = len(prims)
        arr = _struct_array(cls, prims)
        xf = _f32(xforms, (n, 12)) if xforms is not None else None
        mat = np.ascontiguousarray(materials, dtype=np.int32).reshape(n)
        first = C.c_int32(-1)
        _check(self.lib, fn(self.h, C.cast(arr, C.c_void_p), _ptr(xf), _ptr(mat), n, C.byref(first)))
        return first.value

    def add_spheres(self, prims: Sequence[Sphere], materials, xforms=None) -> int:
        return self._add(self.lib.pt_scene_add_spheres, Sphere, prims, xforms, materials)

    def add_rects(self, prims: Sequence[Rect], materials, xforms=None) -> int:
        return self._add(self.lib.pt_scene_add_rects, Rect, prims, xforms, materials)

    def add_boxes(self, prims: Sequence[Box], materials, xforms=None) -> int:
        return self._add(self.lib.pt_scene_add_boxes, Box, prims, xforms, materials)

    def add_moving_spheres(self, prims: Sequence[MovingSphere], materials, xforms=None) -> int:
        return self._add(self.lib.pt_scene_add_moving_spheres, MovingSphere, prims, xforms, materials)

    def add_volume_spheres(self, prims: Sequence[VolumeSphere], materials, xforms=None) -> int:
        return self._add(self.lib.pt_scene_add_volume_spheres, VolumeSphere, prims, xforms, materials)

    def add_volume_boxes(self, prims: Sequence[VolumeBox], materials, xforms=None) -> int:
        return self._add(self.lib.pt_scene_add_volume_boxes, VolumeBox, prims, xforms, materials)

    def add_mesh(self, v, material: int, n=None, uv=None, tex_index=None, xform=None) -> int:
        v = _f32(v).reshape(-1, 9)
        nt = v.shape[0]
        n = _f32(n, (nt, 9)) if n is not None else None
        uv = _f32(uv, (nt, 6)) if uv is not None else None
        ti = np.ascontiguousarray(tex_index, dtype=np.int32).reshape(nt) if tex_index is not None else None
        xf = _f32(xform, (12,)) if xform is not None else None
        first = C.c_int32(-1)
        _check(self.lib, self.lib.pt_scene_add_mesh(self.h, _ptr(v), _ptr(n), _ptr(uv), _ptr(ti), nt,
                                                    _ptr(xf), material, C.byref(first)))
        return first.value

    def set_materials(self, mats: Sequence[Material]):
        arr = _struct_array(Material, mats)
        _check(self.lib, self.lib.pt_scene_set_materials(self.h, C.cast(arr, C.c_void_p), len(mats)))

    def set_textures(self, texs: Sequence[Texture]):
        arr = _struct_array(Texture, texs)
        _check(self.lib, self.lib.pt_scene_set_textures(self.h, C.cast(arr, C.c_void_p), len(texs)))

    def set_texture_vector(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        _check(self.lib, self.lib.pt_scene_set_texture_vector(self.h, _ptr(a), a.size))

    def set_texture_data(self, data):
        a = np.ascontiguousarray(data, dtype=np.float32)
        _check(self.lib, self.lib.pt_scene_set_texture_data(self.h, _ptr(a), a.size))

    def add_image(self, rgba) -> int:
        a = _f32(rgba)
        assert a.ndim == 3 and a.shape[2] == 4
        idx = C.c_int32(-1)
        _check(self.lib, self.lib.pt_scene_add_image(self.h, _ptr(a), a.shape[1], a.shape[0], C.byref(idx)))
        return idx.value

    def set_lights(self, lights: Sequence[Light]):
        arr = _struct_array(Light, lights)
        _check(self.lib, self.lib.pt_scene_set_lights(self.h, C.cast(arr, C.c_void_p), len(lights)))

    def set_camera(self, cam: Camera):
        _check(self.lib, self.lib.pt_scene_set_camera(self.h, C.byref(cam)))

    def set_miss(self, miss: Miss):
        _check(self.lib, self.lib.pt_scene_set_miss(self.h, C.byref(miss)))

    def commit(self):
        _check(self.lib, self.lib.pt_scene_commit(self.h))

    def desc(self) -> SceneDesc:
        d = SceneDesc()
        _check(self.lib, self.lib.pt_scene_get_desc(self.h, C.byref(d)))
        return d

    def bvh_info(self) -> BvhInfo:
        b = BvhInfo()
        _check(self.lib, self.lib.pt_scene_bvh_info(self.h, C.byref(b)))
        return b

    def trace(self, rays: np.ndarray, count_fetches: bool = False):
        """pt_trace_rays: closest hits for a host ray set (RAY_DTYPE array)."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        counters = np.zeros(2, dtype=np.uint64) if count_fetches else None
        _check(self.lib, self.lib.pt_trace_rays(self.h, _ptr(rays), rays.shape[0], _ptr(hits), _ptr(counters)))
        return (hits, counters) if count_fetches else hits


class Renderer:
    """pt_renderer wrapper."""

    def __init__(self, scene: Scene, width: int, height: int, d_accum: int = 0, devices: Optional[Sequence[int]] = None):
        """devices: None = the scene's device (pt_renderer_create); a list = one renderer per device behind this
        handle, frames split, buffers merged by one ncclReduce on read (pt_renderer_create_multi)."""
        self.scene = scene
        self.lib = scene.lib
        self.W, self.H = width, height
        self.h = C.c_void_p()
        if devices is None:
            _check(self.lib, self.lib.pt_renderer_create(scene.h, width, height, C.c_void_p(d_accum or None),
                                                         C.byref(self.h)))
        else:
            ids = (C.c_int32 * len(devices))(*devices)
            _check(self.lib, self.lib.pt_renderer_create_multi(scene.h, width, height, ids, len(devices), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.pt_renderer_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_flags(self, flags: int):
        _check(self.lib, self.lib.pt_renderer_set_flags(self.h, flags))

    def set_pool(self, paths: int):
        _check(self.lib, self.lib.pt_renderer_set_pool(self.h, paths))

    def render(self, first_frame: int, nframes: int, max_depth: int = 50, russian: bool = True,
               clear: bool = False, sync: bool = True):
        _check(self.lib, self.lib.pt_render(self.h, first_frame, nframes, max_depth, int(russian), int(clear)))
        if sync:
            self.sync()

    def sync(self):
        _check(self.lib, self.lib.pt_renderer_sync(self.h))

    def read_accum(self) -> np.ndarray:
        out = np.empty((self.H, self.W, 4), dtype=np.float32)
        _check(self.lib, self.lib.pt_read_accum(self.h, _ptr(out)))
        return out

    def read_display(self) -> np.ndarray:
        """RGBA8 preview (display_buffer of the reference), converted on the device."""
        out = np.empty((self.H, self.W, 4), dtype=np.uint8)
        _check(self.lib, self.lib.pt_read_display(self.h, _ptr(out)))
        return out

    def write_accum(self, accum: np.ndarray):
        """Restore a checkpointed accumulation buffer (continue with render(first_frame=samples done))."""
        a = _f32(accum, (self.H, self.W, 4))
        _check(self.lib, self.lib.pt_write_accum(self.h, _ptr(a)))

    def accum_device_ptr(self) -> int:
        p = C.c_void_p()
        _check(self.lib, self.lib.pt_accum_device_ptr(self.h, C.byref(p)))
        return p.value

    def stats(self) -> Stats:
        s = Stats()
        _check(self.lib, self.lib.pt_get_stats(self.h, C.byref(s)))
        return s


def split_frames(total_frames: int, rank: int, world: int, first_frame: int = 0, lib=None):
    """pt_split_frames: (first_frame, nframes) of participant `rank` of `world`."""
    lib = lib or load()
    a, b = C.c_int32(), C.c_int32()
    _check(lib, lib.pt_split_frames(total_frames, rank, world, first_frame, C.byref(a), C.byref(b)))
    return a.value, b.value


def split_frames_weighted(total_frames: int, rank: int, world: int, weights, first_frame: int = 0, lib=None):
    """pt_split_frames_weighted: shares proportional to `weights` (measured rates of the ranks)."""
    lib = lib or load()
    w = np.ascontiguousarray(weights, dtype=np.float32)
    a, b = C.c_int32(), C.c_int32()
    _check(lib, lib.pt_split_frames_weighted(total_frames, rank, world, _ptr(w), first_frame, C.byref(a), C.byref(b)))
    return a.value, b.value


class Comm:
    """pt_comm wrapper (one process per device): NCCL communicator inside the library."""

    def __init__(self, comm_id: bytes, rank: int, world: int, lib=None):
        self.lib = lib or load()
        self.h = C.c_void_p()
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(comm_id)
        _check(self.lib, self.lib.pt_comm_create(C.cast(buf, C.c_void_p), rank, world, C.byref(self.h)))

    @staticmethod
    def unique_id(lib=None) -> bytes:
        lib = lib or load()
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        _check(lib, lib.pt_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def reduce(self, renderer: "Renderer", root: int = 0):
        _check(self.lib, self.lib.pt_comm_reduce(self.h, renderer.h, root))

    def close(self):
        if self.h:
            self.lib.pt_comm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tonemap_png8(accum: np.ndarray, samples: int) -> np.ndarray:
    """host_includes/image_save.hpp:13-29 — sqrt gamma, clamp, x255.99 (host-side output step)."""
    col = accum[..., :3].astype(np.float32) * np.float32(1.0 / np.float32(samples))
    col = np.sqrt(col, dtype=np.float32)
    col = np.float32(255.99) * np.clip(col, 0.0, 1.0).astype(np.float32)
    return col.astype(np.int32).astype(np.uint8)
