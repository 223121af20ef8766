"""Shared test helpers: package loader, oracle / reference-probe / host-sim
bindings, and numpy scene builders (Cornell, RTOW-style spheres, C3 material
zoo, procedural mesh + HDRI).  The oracle is the CHECKER only."""
from __future__ import annotations

import ctypes as C
import importlib.util
import math
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "optix-path-tracer_b200")
ORACLE_DIR = os.path.join(ROOT, "oracle")


def load_package():
    name = "optix_path_tracer_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


pt = load_package()


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed:\n{r.stdout}\n{r.stderr}")


def build_product():
    if not os.path.exists(pt.LIB_PATH):
        _run(["make", "-j8"], os.path.join(PKG_DIR, "csrc"))
    return pt.LIB_PATH


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ---- oracle ------------------------------------------------------------------------------
class OrcStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("bounces", C.c_uint64), ("node_tests", C.c_uint64), ("prim_tests", C.c_uint64),
                ("seconds", C.c_double), ("threads", C.c_int32)]


_oracles = {}


def oracle(variant=""):
    """liboracle.so, or (variant="_r2l") the build whose two-draws-in-one-argument-list are taken right to left
    like g++ does with the reference's source (only for replaying vectors recorded from that build)."""
    if variant in _oracles:
        return _oracles[variant]
    so = os.path.join(ORACLE_DIR, f"liboracle{variant}.so")
    src_m = max(os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("oracle.cpp", "oracle.h"))
    if not os.path.exists(so) or os.path.getmtime(so) < src_m:
        _run(["make", f"liboracle{variant}.so"], ORACLE_DIR)
    o = C.CDLL(so)
    vp, P = C.c_void_p, C.POINTER
    o.orc_scene_create.restype = vp
    o.orc_scene_create.argtypes = [P(pt.SceneDesc), C.c_int]
    o.orc_scene_destroy.argtypes = [vp]
    o.orc_trace.argtypes = [vp, vp, C.c_int64, vp, P(OrcStats)]
    o.orc_camera_rays.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    o.orc_bounce_rays.argtypes = [vp, vp, C.c_int64, C.c_uint32, vp]
    o.orc_render.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, P(OrcStats)]
    o.orc_render_pixel.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    o.orc_tonemap_png8.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    o.orc_display_rgba8.argtypes = [vp, C.c_int, C.c_int, vp]
    o.orc_tea64.restype = C.c_uint32
    o.orc_tea64.argtypes = [C.c_uint32, C.c_uint32]
    o.orc_rnd.restype = C.c_float
    o.orc_rnd.argtypes = [P(C.c_uint32)]
    F3 = pt.F3
    fl = C.c_float
    o.orc_bsdf_sample.argtypes = [C.c_int, F3, F3, fl, fl, F3, F3, F3, fl, fl, F3]
    o.orc_bsdf_eval.argtypes = [C.c_int, F3, F3, fl, fl, F3, F3, F3, F3, F3, P(fl)]
    o.orc_schlick.restype = fl
    o.orc_schlick.argtypes = [fl, fl]
    o.orc_refract.argtypes = [F3, F3, fl, F3]
    o.orc_ggx_sample.argtypes = [F3, fl, fl, fl, fl, F3]
    o.orc_ggx_d.restype = fl
    o.orc_ggx_d.argtypes = [F3, fl, fl]
    o.orc_ggx_g.restype = fl
    o.orc_ggx_g.argtypes = [F3, F3, fl, fl]
    o.orc_rect_sample.argtypes = [P(pt.Light), F3, fl, fl, F3]
    o.orc_rect_pdf.restype = fl
    o.orc_rect_pdf.argtypes = [P(pt.Light), F3, F3]
    o.orc_intersect_prim.argtypes = [vp, C.c_int, P(pt.Ray), P(fl), P(fl), P(fl)]
    o.orc_intersect_prim_ex.argtypes = [vp, C.c_int, P(pt.Ray), fl, C.c_uint32, vp, vp]
    o.orc_closest_hit_probe.argtypes = [vp, P(pt.Ray), C.c_int, fl, fl, fl, vp, vp]
    o.orc_trace_radiance.argtypes = [vp, vp, vp]
    o.orc_sample_texture.argtypes = [vp, C.c_int, fl, fl, F3, C.c_int, F3]
    o.orc_miss_probe.argtypes = [vp, F3, F3, F3, vp]
    o.orc_prim_bounds.argtypes = [vp, C.c_int, vp]
    o.orc_light_sample.argtypes = [P(pt.Light), F3, P(C.c_uint32), F3]
    o.orc_light_pdf.restype = fl
    o.orc_light_pdf.argtypes = [P(pt.Light), F3, F3]
    o.orc_volume_seed.restype = C.c_uint32
    o.orc_volume_seed.argtypes = [C.c_uint32, C.c_uint32]
    for n in ("orc_random_in_unit_disk", "orc_random_in_unit_sphere", "orc_random_on_unit_sphere"):
        getattr(o, n).argtypes = [P(C.c_uint32), F3]
    _oracles[variant] = o
    return o


class OracleScene:
    """orc_scene over a pt.Scene's flat description (keeps the Scene alive)."""

    def __init__(self, scene, mode: int = 0, variant: str = ""):
        self.o = oracle(variant)
        self.scene = scene
        self.desc = scene.desc()
        self.h = self.o.orc_scene_create(C.byref(self.desc), mode)

    def close(self):
        if self.h:
            self.o.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, dtype=pt.RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=pt.HIT_DTYPE)
        st = OrcStats()
        self.o.orc_trace(self.h, rays.ctypes.data, rays.shape[0], hits.ctypes.data, C.byref(st))
        return hits, st

    def camera_rays(self, W, H, frame=0):
        rays = np.empty(W * H, dtype=pt.RAY_DTYPE)
        self.o.orc_camera_rays(self.h, W, H, frame, rays.ctypes.data)
        return rays

    def bounce_rays(self, rays, seed=1):
        rays = np.ascontiguousarray(rays, dtype=pt.RAY_DTYPE)
        out = np.empty_like(rays)
        self.o.orc_bounce_rays(self.h, rays.ctypes.data, rays.shape[0], seed, out.ctypes.data)
        return out

    def render(self, W, H, first_frame, nframes, max_depth=50, russian=True):
        acc = np.zeros((H, W, 4), dtype=np.float32)
        st = OrcStats()
        self.o.orc_render(self.h, W, H, first_frame, nframes, max_depth, int(russian), acc.ctypes.data, C.byref(st))
        return acc, st


def ref_probe():
    """Reference-header probe (oracle/_ref/libref_probe.so) or None."""
    so = os.path.join(ORACLE_DIR, "_ref", "libref_probe.so")
    if not os.path.exists(so) and os.path.isdir("/root/reference/OptiX-Path-Tracer/programs"):
        _run(["make", "ref"], ORACLE_DIR)
    if not os.path.exists(so):
        return None
    r = C.CDLL(so)
    F3, fl, P = pt.F3, C.c_float, C.POINTER
    r.ref_tea64.restype = C.c_uint32
    r.ref_tea64.argtypes = [C.c_uint32, C.c_uint32]
    r.ref_rnd.restype = fl
    r.ref_rnd.argtypes = [P(C.c_uint32)]
    r.ref_bsdf_sample.argtypes = [C.c_int, F3, F3, fl, fl, F3, F3, F3, P(C.c_uint32), F3]
    r.ref_bsdf_eval.argtypes = [C.c_int, F3, F3, fl, fl, F3, F3, F3, F3, F3, P(fl)]
    r.ref_schlick.restype = fl
    r.ref_schlick.argtypes = [fl, fl]
    r.ref_refract.argtypes = [F3, F3, fl, F3]
    r.ref_ggx_sample.argtypes = [F3, fl, fl, fl, fl, F3]
    r.ref_ggx_d.restype = fl
    r.ref_ggx_d.argtypes = [F3, fl, fl]
    r.ref_ggx_g.restype = fl
    r.ref_ggx_g.argtypes = [F3, F3, fl, fl]
    for n in ("ref_random_in_unit_disk", "ref_random_in_unit_sphere", "ref_random_on_unit_sphere"):
        getattr(r, n).argtypes = [P(C.c_uint32), F3]
    return r


RADIANCE_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


def ref_ray_probe():
    """The reference's raygen.cu and miss.cu compiled unmodified as host code (oracle/_ref/libref_ray.so,
    oracle/ref_ray_probe.cpp) or None."""
    so = os.path.join(ORACLE_DIR, "_ref", "libref_ray.so")
    if not os.path.exists(so) and os.path.isdir("/root/reference/OptiX-Path-Tracer/programs"):
        _run(["make", "ref"], ORACLE_DIR)
    if not os.path.exists(so):
        return None
    r = C.CDLL(so)
    vp, ci = C.c_void_p, C.c_int
    r.refray_render.argtypes = [vp, ci, ci, ci, ci, ci, ci, RADIANCE_HOOK, vp, vp]
    r.refmiss.argtypes = [ci, ci, pt.F3, pt.F3, pt.F3, pt.F3, pt.F3, pt.F3, vp, C.POINTER(ci)]
    return r


class ReferenceRenderer:
    """The reference itself on the CPU: its own programs (compiled unmodified) inside an emulated OptiX pipeline
    (oracle/ref_render.cpp -> oracle/_ref/libref_render.so).  One rendering thread per process."""

    @staticmethod
    def lib():
        so = os.path.join(ORACLE_DIR, "_ref", "libref_render.so")
        if not os.path.exists(so) and os.path.isdir("/root/reference/OptiX-Path-Tracer/programs"):
            _run(["make", "ref"], ORACLE_DIR)
        if not os.path.exists(so):
            return None
        r = C.CDLL(so)
        ci, vp = C.c_int, C.c_void_p
        r.refrender_create.restype = vp
        r.refrender_create.argtypes = [C.POINTER(pt.SceneDesc)]
        r.refrender_destroy.argtypes = [vp]
        r.refrender_render.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp]
        return r

    def __init__(self, scene):
        self.r = self.lib()
        if self.r is None:
            raise RuntimeError("oracle/_ref/libref_render.so is not built (needs /root/reference)")
        self.scene = scene
        self.desc = scene.desc()
        self.h = self.r.refrender_create(C.byref(self.desc))

    def render(self, W, H, first_frame, nframes, max_depth=50, russian=True, rows=None, acc=None):
        """frames [first, first + n) of rows [y0, y1) (0 = bottom) added into acc; returns (acc, display)"""
        if acc is None:
            acc = np.zeros((H, W, 4), dtype=np.float32)
        disp = np.zeros((H, W, 4), dtype=np.uint8)
        y0, y1 = rows if rows else (0, H)
        self.r.refrender_render(self.h, W, H, y0, y1, first_frame, nframes, max_depth, int(russian), acc.ctypes.data,
                                disp.ctypes.data)
        return acc, disp

    def __del__(self):
        try:
            self.r.refrender_destroy(self.h)
        except Exception:
            pass


def ref_tex_probe():
    """The reference's textures/*.cu compiled unmodified as host code (oracle/_ref/libref_tex.so) or None."""
    so = os.path.join(ORACLE_DIR, "_ref", "libref_tex.so")
    if not os.path.exists(so) and os.path.isdir("/root/reference/OptiX-Path-Tracer/programs"):
        _run(["make", "ref"], ORACLE_DIR)
    if not os.path.exists(so):
        return None
    r = C.CDLL(so)
    for name in ("constant", "checker", "gradient", "vector", "noise"):
        getattr(r, "reftex_" + name).argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, pt.F3, C.c_int, pt.F3]
    return r


def ref_mat_probe():
    """The reference's materials/*.cu closest-hit programs (with light_sample.cuh) compiled unmodified as host
    code (oracle/_ref/libref_mat.so, oracle/ref_mat_probe.cpp) or None."""
    so = os.path.join(ORACLE_DIR, "_ref", "libref_mat.so")
    if not os.path.exists(so) and os.path.isdir("/root/reference/OptiX-Path-Tracer/programs"):
        _run(["make", "ref"], ORACLE_DIR)
    if not os.path.exists(so):
        return None
    r = C.CDLL(so)
    vp = C.c_void_p
    for name in MAT_PROGRAMS.values():
        getattr(r, "refmat_" + name).argtypes = [vp, vp, C.c_float, vp, vp, vp, C.c_int, vp, vp]
    return r


MAT_PROGRAMS = {0: "lambertian", 1: "metal", 2: "dielectric", 3: "diffuse_light", 4: "oren_nayar",
                5: "torrance_sparrow", 6: "ashikhmin_shirley", 7: "normal_shader", 8: "isotropic"}


def ref_cu_probe():
    """The reference's hitables/*.cu and pdfs/*.cu compiled unmodified as host code
    (oracle/_ref/libref_cu.so, oracle/ref_cu_probe.cpp) or None."""
    so = os.path.join(ORACLE_DIR, "_ref", "libref_cu.so")
    if not os.path.exists(so) and os.path.isdir("/root/reference/OptiX-Path-Tracer/programs"):
        _run(["make", "ref"], ORACLE_DIR)
    if not os.path.exists(so):
        return None
    r = C.CDLL(so)
    vp, fl = C.c_void_p, C.c_float
    for kind in ("sphere", "rect", "box", "triangle", "moving_sphere", "volume_sphere", "volume_box"):
        getattr(r, "refcu_" + kind).argtypes = [vp, vp, fl, C.c_uint32, vp, vp, vp]
    r.refcu_rect_pdf.restype = fl
    r.refcu_rect_pdf.argtypes = [vp, C.c_int, pt.F3, pt.F3]
    r.refcu_rect_sample.argtypes = [vp, C.c_int, pt.F3, C.POINTER(C.c_uint32), pt.F3]
    r.refcu_spherelight_pdf.restype = fl
    r.refcu_spherelight_pdf.argtypes = [vp, pt.F3, pt.F3]
    r.refcu_spherelight_sample.argtypes = [vp, pt.F3, C.POINTER(C.c_uint32), pt.F3]
    return r


# ---- host simulation of the kernel cores (test-only) ----------------------------------------
_sim = None


def hostsim():
    global _sim
    if _sim is not None:
        return _sim
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libhostsim.so")
    srcs = [os.path.join(d, "sim.cpp")] + [os.path.join(PKG_DIR, "csrc", f)
                                            for f in ("hd.h", "bvh_core.h", "traverse.h", "shading.h", "scene_dev.h",
                                                      "sample_order.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        _run(["/usr/bin/g++", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-w",
              "-o", so, "sim.cpp"], d)
    s = C.CDLL(so)
    vp, P = C.c_void_p, C.POINTER
    s.sim_create.restype = vp
    s.sim_create.argtypes = [P(pt.SceneDesc), C.c_int]
    s.sim_destroy.argtypes = [vp]
    s.sim_validate.argtypes = [vp]
    s.sim_info.argtypes = [vp, P(C.c_int), P(C.c_double), P(C.c_double), P(C.c_int)]
    s.sim_trace.argtypes = [vp, vp, C.c_longlong, vp, vp]
    s.sim_render.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
    s.sim_perm8.restype = C.c_uint
    s.sim_perm8.argtypes = [C.c_uint, C.c_uint]
    s.sim_node_invariant_violations.argtypes = [vp]
    s.sim_sample_order.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_int, vp]
    _sim = s
    return s


class SimScene:
    def __init__(self, scene, refine=True):
        self.s = hostsim()
        self.scene = scene
        self.desc = scene.desc()
        self.h = self.s.sim_create(C.byref(self.desc), int(refine))

    def close(self):
        if self.h:
            self.s.sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def validate(self):
        return self.s.sim_validate(self.h)

    def info(self):
        n, a, b, t = C.c_int(), C.c_double(), C.c_double(), C.c_int()
        self.s.sim_info(self.h, C.byref(n), C.byref(a), C.byref(b), C.byref(t))
        return dict(n_nodes8=n.value, sah_before=a.value, sah_after=b.value, n_treelets=t.value)

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, dtype=pt.RAY_DTYPE)
        hits = np.empty(rays.shape[0], dtype=pt.HIT_DTYPE)
        cnt = np.zeros(2, dtype=np.uint64)
        self.s.sim_trace(self.h, rays.ctypes.data, rays.shape[0], hits.ctypes.data, cnt.ctypes.data)
        return hits, cnt

    def render(self, W, H, first_frame, nframes, max_depth=50, russian=True):
        acc = np.zeros((H, W, 4), dtype=np.float32)
        self.s.sim_render(self.h, W, H, first_frame, nframes, max_depth, int(russian), acc.ctypes.data)
        return acc


# ---- host-side scene helpers (float32 arithmetic like host_includes/*.hpp) -------------------
f32 = np.float32
PI_F = f32(3.141592654)


def _v(x, y, z):
    return np.array([x, y, z], dtype=np.float32)


_libm = C.CDLL("libm.so.6")
_libm.tanf.restype = C.c_float
_libm.tanf.argtypes = [C.c_float]


def _dot(a, b):
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def _cross(a, b):
    return np.array([f32(f32(a[1] * b[2]) - f32(a[2] * b[1])), f32(f32(a[2] * b[0]) - f32(a[0] * b[2])),
                     f32(f32(a[0] * b[1]) - f32(a[1] * b[0]))], dtype=np.float32)


def _unit(v):
    """programs/vec.hpp unit_vector: v / length(v) with optix's float3 / float = v * (1 / s)"""
    inv = f32(1.0) / f32(np.sqrt(_dot(v, v), dtype=np.float32))
    return (v * inv).astype(np.float32)


def make_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, t0=0.0, t1=1.0):
    """host_includes/camera.hpp:16-38, operation for operation in binary32 (pinned bit-exactly against the reference's
    own camera.hpp by tests/test_oracle_host_golden.py)"""
    lookfrom, lookat, vup = (np.asarray(a, dtype=np.float32) for a in (lookfrom, lookat, vup))
    theta = f32(f32(f32(vfov) * PI_F) / f32(180.0))
    half_height = f32(_libm.tanf(float(f32(theta / f32(2.0)))))
    half_width = f32(f32(aspect) * half_height)
    w = _unit((lookfrom - lookat).astype(np.float32))
    u = _unit(_cross(vup, w))
    v = _cross(w, u)
    fd = f32(focus_dist)
    llc = (((lookfrom - f32(half_width * fd) * u).astype(np.float32) - f32(half_height * fd) * v).astype(np.float32)
           - fd * w).astype(np.float32)
    cam = pt.Camera()
    cam.origin = pt.F3(*lookfrom)
    cam.lower_left_corner = pt.F3(*llc)
    cam.horizontal = pt.F3(*(f32(f32(f32(2.0) * half_width) * fd) * u).astype(np.float32))
    cam.vertical = pt.F3(*(f32(f32(f32(2.0) * half_height) * fd) * v).astype(np.float32))
    cam.u, cam.v, cam.w = pt.F3(*u), pt.F3(*v), pt.F3(*w)
    cam.lens_radius = f32(aperture) / f32(2.0)
    cam.time0, cam.time1 = t0, t1
    return cam


def mat_translate(t):
    m = np.eye(4, dtype=np.float64)
    m[:3, 3] = t
    return m


def mat_scale(s):
    return np.diag([s[0], s[1], s[2], 1.0]).astype(np.float64)


def mat_rotate(deg, axis):
    """optix::Matrix4x4::rotate(deg*PI_F/180, unit axis) (host_includes/transforms.hpp:103-135)"""
    rad = float(f32(deg) * PI_F / f32(180.0))
    a = np.zeros(3)
    a[axis] = 1.0
    c, s = math.cos(rad), math.sin(rad)
    ux, uy, uz = a
    m = np.eye(4)
    m[0, 0] = ux * ux + c * (1 - ux * ux); m[0, 1] = ux * uy * (1 - c) - uz * s; m[0, 2] = uz * ux * (1 - c) + uy * s
    m[1, 0] = ux * uy * (1 - c) + uz * s; m[1, 1] = uy * uy + c * (1 - uy * uy); m[1, 2] = uy * uz * (1 - c) - ux * s
    m[2, 0] = uz * ux * (1 - c) - uy * s; m[2, 1] = uy * uz * (1 - c) + ux * s; m[2, 2] = uz * uz + c * (1 - uz * uz)
    return m


def m34(m):
    return np.asarray(m, dtype=np.float64)[:3, :].astype(np.float32).reshape(12)


def tex_const(r, g=None, b=None):
    if g is None:
        g = b = r
    return pt.Texture(pt.TEX_CONSTANT, pt.F3(r, g, b), 0, 0)


def material(kind, tex0=-1, tex1=-1, p0=0.0, p1=0.0, flags=0):
    return pt.Material(kind, tex0, tex1, p0, p1, flags)


def oren_nayar_params(R):
    """host_includes/materials.hpp:192-198"""
    sigma = f32(min(max(R, 0.0), 1.0))
    div = f32(1.0) / (PI_F + ((f32(3.0) * PI_F - f32(4.0)) / f32(6.0)) * sigma)
    return float(f32(1.0) * div), float(sigma * div)


def ts_alpha(R):
    """host_includes/materials.hpp:235-238"""
    r = f32(max(R, 1e-3))
    return float(r * r)


def build_cornell(W=64, H=64, lib=None, extra=None):
    """host_includes/scenes.hpp:188-260 (Cornell): 6 rects + rotated tall box, rect light."""
    s = pt.Scene(lib)
    texs = [tex_const(0.65, 0.05, 0.05), tex_const(0.73), tex_const(0.12, 0.45, 0.15), tex_const(7.0)]
    mats = [material(pt.MAT_LAMBERTIAN, 0), material(pt.MAT_LAMBERTIAN, 1), material(pt.MAT_LAMBERTIAN, 2),
            material(pt.MAT_DIFFUSE_LIGHT, 3)]
    RED, WHITE, GREEN, LIGHT = 0, 1, 2, 3
    rects = [pt.Rect(0, 555, 0, 555, 555, pt.X_AXIS, 1), pt.Rect(0, 555, 0, 555, 0, pt.X_AXIS, 0),
             pt.Rect(213, 343, 227, 332, 554, pt.Y_AXIS, 1), pt.Rect(0, 555, 0, 555, 555, pt.Y_AXIS, 1),
             pt.Rect(0, 555, 0, 555, 0, pt.Y_AXIS, 0), pt.Rect(0, 555, 0, 555, 555, pt.Z_AXIS, 1)]
    rmat = [RED, GREEN, LIGHT, WHITE, WHITE, WHITE]
    if extra is not None:
        extra(s, texs, mats)
    s.set_textures(texs)
    s.set_materials(mats)
    s.add_rects(rects, rmat)
    # box0.translate(265,0,295); box0.rotate(15,Y) via addElementsTo => world = T * R * p
    xf = mat_translate([265.0, 0.0, 295.0]) @ mat_rotate(15.0, pt.Y_AXIS)
    s.add_boxes([pt.Box(pt.F3(0, 0, 0), pt.F3(165, 330, 165))], [GREEN], xforms=m34(xf))
    s.set_lights([pt.Light(0, pt.Y_AXIS, 213, 343, 227, 332, 554, pt.F3(7, 7, 7))])
    s.set_camera(make_camera([278, 278, -800], [278, 278, 0], [0, 1, 0], 40.0, W / H, 0.0, 10.0))
    s.set_miss(pt.Miss(pt.MISS_CONSTANT, pt.F3(0, 0, 0), pt.F3(0, 0, 0), -1, 1, 0))
    return s


def build_sphere_light_scene(W=48, H=48, lib=None, two=False):
    """Cornell walls lit by an emissive Sphere sampled through Sphere_PDF
    (host_includes/pdfs.hpp:82-109, programs/pdfs/sphere_pdf.cu); `two` adds the rect light back
    so that light picking covers both PDF kinds."""
    def extra(s, texs, mats):
        pass
    s = build_cornell(W, H, lib, extra)
    s.add_spheres([pt.Sphere(pt.F3(278, 400, 278), 60.0)], [3])
    sph = pt.Light(pt.LIGHT_SPHERE, 0, 278, 400, 278, 0, 60.0, pt.F3(5, 5, 5))
    rect = pt.Light(pt.LIGHT_RECT, pt.Y_AXIS, 213, 343, 227, 332, 554, pt.F3(7, 7, 7))
    s.set_lights([sph, rect] if two else [sph])
    return s


def build_motion_scene(W=48, H=32, lib=None, n=5, seed=3):
    """Moving_Sphere (programs/hitables/moving_sphere.cu): a ground sphere, a grid of small spheres
    bouncing upwards during the shutter interval [0, 1], one big sphere sliding sideways, static
    metal / glass spheres in between; gradient sky."""
    rng = np.random.default_rng(seed)
    s = pt.Scene(lib)
    texs = [tex_const(0.5, 0.5, 0.5), tex_const(0.4, 0.2, 0.1), tex_const(0.7, 0.6, 0.5), tex_const(1.0)]
    mats = [material(pt.MAT_LAMBERTIAN, 0), material(pt.MAT_LAMBERTIAN, 1), material(pt.MAT_METAL, 2, -1, 0.0),
            material(pt.MAT_DIELECTRIC, 3, 3, 1.5)]
    moving, mmat, static, smat = [], [], [pt.Sphere(pt.F3(0, -1000, 0), 1000.0)], [0]
    for a in range(-n, n):
        for b in range(-n, n):
            c = (a + 0.9 * rng.random(), 0.2, b + 0.9 * rng.random())
            k = rng.random()
            if k < 0.7:
                texs.append(tex_const(*rng.random(3)))
                mats.append(material(pt.MAT_LAMBERTIAN, len(texs) - 1))
                c2 = (c[0], c[1] + 0.5 * rng.random(), c[2])
                moving.append(pt.MovingSphere(pt.F3(*c), pt.F3(*c2), 0.2, 0.0, 1.0))
                mmat.append(len(mats) - 1)
            else:
                static.append(pt.Sphere(pt.F3(*c), 0.2))
                smat.append(2 if k < 0.9 else 3)
    moving.append(pt.MovingSphere(pt.F3(-3, 1, 0), pt.F3(-3, 1, 1.5), 1.0, 0.0, 1.0))
    mmat.append(1)
    static.append(pt.Sphere(pt.F3(3, 1, 0), 1.0))
    smat.append(2)
    s.set_textures(texs)
    s.set_materials(mats)
    s.add_spheres(static, smat)
    s.add_moving_spheres(moving, mmat)
    s.set_camera(make_camera([13, 2, 3], [0, 0, 0], [0, 1, 0], 20.0, W / H, 0.0, 10.0, 0.0, 1.0))
    s.set_miss(pt.Miss(pt.MISS_GRADIENT, pt.F3(1, 1, 1), pt.F3(0.5, 0.7, 1.0), -1, 1, 0))
    return s


def build_smoke_scene(W=48, H=48, lib=None, fog=True):
    """Volumes (programs/hitables/volume_box.cu, volume_sphere.cu + materials/isotropic.cu): Cornell
    walls, two rotated blocks of smoke, a fog ball and (fog=True) a thin fog filling the room."""
    def extra(s, texs, mats):
        base = len(texs)
        texs.extend([tex_const(0.0), tex_const(1.0), tex_const(0.2, 0.4, 0.9)])
        mats.extend([material(pt.MAT_ISOTROPIC, base), material(pt.MAT_ISOTROPIC, base + 1),
                     material(pt.MAT_ISOTROPIC, base + 2)])
    s = pt.Scene(lib)
    texs = [tex_const(0.65, 0.05, 0.05), tex_const(0.73), tex_const(0.12, 0.45, 0.15), tex_const(7.0)]
    mats = [material(pt.MAT_LAMBERTIAN, 0), material(pt.MAT_LAMBERTIAN, 1), material(pt.MAT_LAMBERTIAN, 2),
            material(pt.MAT_DIFFUSE_LIGHT, 3)]
    extra(s, texs, mats)
    s.set_textures(texs)
    s.set_materials(mats)
    rects = [pt.Rect(0, 555, 0, 555, 555, pt.X_AXIS, 1), pt.Rect(0, 555, 0, 555, 0, pt.X_AXIS, 0),
             pt.Rect(113, 443, 127, 432, 554, pt.Y_AXIS, 1), pt.Rect(0, 555, 0, 555, 555, pt.Y_AXIS, 1),
             pt.Rect(0, 555, 0, 555, 0, pt.Y_AXIS, 0), pt.Rect(0, 555, 0, 555, 555, pt.Z_AXIS, 1)]
    s.add_rects(rects, [0, 2, 3, 1, 1, 1])
    xf0 = mat_translate([265.0, 0.0, 295.0]) @ mat_rotate(15.0, pt.Y_AXIS)
    xf1 = mat_translate([130.0, 0.0, 65.0]) @ mat_rotate(-18.0, pt.Y_AXIS)
    s.add_volume_boxes([pt.VolumeBox(pt.F3(0, 0, 0), pt.F3(165, 330, 165), 0.01),
                        pt.VolumeBox(pt.F3(0, 0, 0), pt.F3(165, 165, 165), 0.01)], [4, 5],
                       xforms=np.stack([m34(xf0), m34(xf1)]))
    vs = [pt.VolumeSphere(pt.F3(400, 420, 150), 70.0, 0.02)]
    vm = [6]
    if fog:
        vs.append(pt.VolumeSphere(pt.F3(0, 0, 0), 5000.0, 0.0001))
        vm.append(5)
    s.add_volume_spheres(vs, vm)
    s.set_lights([pt.Light(0, pt.Y_AXIS, 113, 443, 127, 432, 554, pt.F3(7, 7, 7))])
    s.set_camera(make_camera([278, 278, -800], [278, 278, 0], [0, 1, 0], 40.0, W / H, 0.0, 10.0))
    s.set_miss(pt.Miss(pt.MISS_CONSTANT, pt.F3(0, 0, 0), pt.F3(0, 0, 0), -1, 1, 0))
    return s


def build_spheres(n_small=22, seed=0, W=64, H=64, lib=None, lights=False):
    """RTOW-style scene (host_includes/scenes.hpp:17-93): ground + grid of small spheres
    (Lambertian / metal / dielectric) + three big ones, gradient sky."""
    rng = np.random.default_rng(seed)
    s = pt.Scene(lib)
    texs, mats, sph, smat = [], [], [], []

    def add(center, radius, kind, col, p0=0.0, col2=None):
        texs.append(tex_const(*col))
        t0 = len(texs) - 1
        t1 = -1
        if col2 is not None:
            texs.append(tex_const(*col2))
            t1 = len(texs) - 1
        mats.append(material(kind, t0, t1, p0, 0.0))
        sph.append(pt.Sphere(pt.F3(*center), radius))
        smat.append(len(mats) - 1)

    add((0, -1000, -1), 1000, pt.MAT_LAMBERTIAN, (0.5, 0.5, 0.5))
    half = n_small // 2
    for a in range(-half, n_small - half):
        for b in range(-half, n_small - half):
            choose = rng.random()
            c = (a + rng.random(), 0.2, b + rng.random())
            if choose < 0.8:
                add(c, 0.2, pt.MAT_LAMBERTIAN, rng.random(3))
            elif choose < 0.95:
                add(c, 0.2, pt.MAT_METAL, 0.5 * (1 + rng.random(3)), 0.5 * rng.random())
            else:
                add(c, 0.2, pt.MAT_DIELECTRIC, (1, 1, 1), 1.5, rng.random(3))
    add((4, 1, 0), 1.0, pt.MAT_DIELECTRIC, (1, 1, 1), 1.5, (1, 1, 1))
    add((0, 1, 0.5), 1.0, pt.MAT_LAMBERTIAN, (0.4, 0.2, 0.1))
    add((-4, 1, 1), 1.0, pt.MAT_METAL, (0.7, 0.6, 0.5), 0.0)
    s.set_textures(texs)
    s.set_materials(mats)
    s.add_spheres(sph, smat)
    s.set_camera(make_camera([13, 2, 3], [0, 0, 0], [0, 1, 0], 20.0, W / H, 0.1, 10.0))
    s.set_miss(pt.Miss(pt.MISS_GRADIENT, pt.F3(1, 1, 1), pt.F3(0.5, 0.7, 1.0), -1, 1, 0))
    return s


def build_material_zoo(W=64, H=64, lib=None):
    """Config C3: Cornell walls + glass sphere + Oren-Nayar / Torrance-Sparrow /
    Ashikhmin-Shirley objects (placed on spheres: those BSDFs NaN when N || Y)."""
    def extra(s, texs, mats):
        base = len(texs)
        texs.extend([tex_const(1.0), tex_const(0.8, 0.6, 0.2), tex_const(0.8, 0.85, 0.88),
                     tex_const(0.3, 0.4, 0.8), tex_const(0.9, 0.9, 0.9)])
        rA, rB = oren_nayar_params(0.5)
        a = ts_alpha(0.2)
        mats.extend([material(pt.MAT_DIELECTRIC, base, base, 1.5, 0.0),
                     material(pt.MAT_OREN_NAYAR, base + 1, -1, rA, rB),
                     material(pt.MAT_TORRANCE_SPARROW, base + 2, -1, a, a),
                     material(pt.MAT_ASHIKHMIN_SHIRLEY, base + 3, base + 4, 100.0, 10.0),
                     material(pt.MAT_METAL, base + 2, -1, 0.1)])
    s = build_cornell(W, H, lib, extra)
    sph = [pt.Sphere(pt.F3(130, 90, 120), 90), pt.Sphere(pt.F3(420, 70, 150), 70),
           pt.Sphere(pt.F3(120, 320, 380), 80), pt.Sphere(pt.F3(300, 420, 300), 60),
           pt.Sphere(pt.F3(450, 250, 420), 55)]
    s.add_spheres(sph, [4, 5, 6, 7, 8])
    return s


def heightfield_mesh(n, size=10.0, amp=0.6, seed=1234, normals=False):
    """n x n grid of quads (2 triangles each) displaced by a smooth height function plus
    0.05*hash jitter.  Returns (v[ntris,9], n[ntris,9] or None)."""
    xs = np.linspace(-size / 2, size / 2, n + 1, dtype=np.float64)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    rng = np.random.default_rng(seed)
    Y = amp * (np.sin(1.3 * X) * np.cos(0.9 * Z) + 0.5 * np.sin(2.7 * X + 1.0) * np.sin(3.1 * Z)) \
        + 0.05 * (size / n) * rng.standard_normal(X.shape)
    P = np.stack([X, Y, Z], axis=-1).astype(np.float32)
    p00, p10, p01, p11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    t1 = np.stack([p00, p01, p11], axis=2)
    t2 = np.stack([p00, p11, p10], axis=2)
    tris = np.concatenate([t1.reshape(-1, 9), t2.reshape(-1, 9)], axis=0).astype(np.float32)
    nrm = None
    if normals:
        gy, gx = np.gradient(Y, xs, xs)
        N = np.stack([-gy, np.ones_like(Y), -gx], axis=-1)
        N = (N / np.linalg.norm(N, axis=-1, keepdims=True)).astype(np.float32)
        n00, n10, n01, n11 = N[:-1, :-1], N[1:, :-1], N[:-1, 1:], N[1:, 1:]
        nrm = np.concatenate([np.stack([n00, n01, n11], axis=2).reshape(-1, 9),
                              np.stack([n00, n11, n10], axis=2).reshape(-1, 9)], axis=0).astype(np.float32)
    return tris, nrm


def synthetic_hdri(w=256, h=128, peak=0.5):
    """smooth sky gradient + gaussian sun, values <= peak (SURVEY §7 hard part 4)"""
    v = (np.arange(h, dtype=np.float64) + 0.5) / h
    u = (np.arange(w, dtype=np.float64) + 0.5) / w
    U, V = np.meshgrid(u, v)
    sky = 0.15 + 0.25 * (1.0 - V)
    sun = np.exp(-(((U - 0.3) / 0.03) ** 2 + ((V - 0.25) / 0.04) ** 2))
    img = np.zeros((h, w, 4), dtype=np.float32)
    img[..., 0] = np.minimum(peak, sky * 0.9 + peak * sun)
    img[..., 1] = np.minimum(peak, sky * 1.0 + peak * sun * 0.95)
    img[..., 2] = np.minimum(peak, sky * 1.2 + peak * sun * 0.8)
    return img


def build_mesh_scene(grid=32, W=64, H=64, lib=None, normals=False, mat_kind=None, env=True, env_nee=False):
    """Config C4 shape: procedural triangle mesh lit by a synthetic HDRI environment map."""
    s = pt.Scene(lib)
    texs = [tex_const(0.7, 0.7, 0.7)]
    kind = pt.MAT_LAMBERTIAN if mat_kind is None else mat_kind
    mats = [material(kind, 0)]
    s.set_textures(texs)
    s.set_materials(mats)
    tris, nrm = heightfield_mesh(grid, normals=normals)
    s.add_mesh(tris, 0, n=nrm)
    s.set_camera(make_camera([0, 4.5, -9], [0, 0, 0], [0, 1, 0], 40.0, W / H, 0.0, 10.0))
    if env:
        img = s.add_image(synthetic_hdri())
        s.set_miss(pt.Miss(pt.MISS_HDR, pt.F3(0, 0, 0), pt.F3(0, 0, 0), img, 1, int(env_nee)))
    else:
        s.set_miss(pt.Miss(pt.MISS_GRADIENT, pt.F3(1, 1, 1), pt.F3(0.5, 0.7, 1.0), -1, 1, 0))
    return s


def build_backdrop(W, H, miss_type, spherical=1, lib=None):
    """a mirror sphere and a diffuse one in front of an image background (miss.cu image_background /
    environmental_mapping): the miss program's own (u, v) arithmetic decides every pixel"""
    s = pt.Scene(lib)
    s.set_textures([tex_const(0.8, 0.6, 0.2), tex_const(0.9)])
    s.set_materials([material(pt.MAT_LAMBERTIAN, 0), material(pt.MAT_METAL, 1, p0=0.0)])
    s.add_spheres([pt.Sphere(pt.F3(-1.1, 0, 0), 1.0), pt.Sphere(pt.F3(1.1, 0, 0), 1.0)], [0, 1])
    s.set_camera(make_camera([0, 0.5, -6], [0, 0, 0], [0, 1, 0], 40.0, W / H, 0.0, 6.0))
    img = s.add_image(synthetic_hdri(64, 32))
    s.set_miss(pt.Miss(miss_type, pt.F3(0, 0, 0), pt.F3(0, 0, 0), img, spherical, 0))
    return s


def build_noise_scene(W=32, H=32):
    """Spheres with Perlin-noise, gradient and checker textures on Lambertian / metal."""
    rng = np.random.default_rng(7)
    s = pt.Scene()
    g = rng.normal(size=(256, 3))
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    perms = [rng.permutation(256) for _ in range(3)]
    data = np.concatenate([g.reshape(-1), *[p.astype(np.float64) for p in perms],
                           [0.9, 0.1, 0.1, 0.1, 0.9, 0.1, 0.1, 0.1, 0.9]]).astype(np.float32)
    s.set_texture_data(data)
    texs = [pt.Texture(pt.TEX_NOISE, pt.F3(4.0, 0, 0), 0, 0), pt.Texture(pt.TEX_GRADIENT, pt.F3(0, 0, 0), 1536, 0),
            tex_const(0.2, 0.3, 0.1), tex_const(0.9, 0.9, 0.9), pt.Texture(pt.TEX_CHECKER, pt.F3(0, 0, 0), 2, 3),
            pt.Texture(pt.TEX_NOISE, pt.F3(0.7, 0, 0), 0, 1)]
    s.set_textures(texs)
    s.set_materials([material(pt.MAT_LAMBERTIAN, 0), material(pt.MAT_LAMBERTIAN, 1), material(pt.MAT_LAMBERTIAN, 4),
                     material(pt.MAT_METAL, 5, p0=0.05)])
    s.add_spheres([pt.Sphere(pt.F3(0, -1000, 0), 1000), pt.Sphere(pt.F3(-2.2, 1, 0), 1.0), pt.Sphere(pt.F3(0, 1, 0), 1.0),
                   pt.Sphere(pt.F3(2.2, 1, 0), 1.0)], [2, 0, 1, 3])
    s.set_camera(make_camera([0, 2.5, 7], [0, 1, 0], [0, 1, 0], 40, W / H, 0.0, 7.0))
    s.set_miss(pt.Miss(pt.MISS_GRADIENT, pt.F3(1, 1, 1), pt.F3(0.5, 0.7, 1.0), -1, 1, 0))
    return s


def build_random_prims(seed=0, n_sph=40, n_rect=30, n_box=30, n_tri=200, xforms=True, lib=None, W=32, H=32,
                       n_moving=0, n_vsph=0, n_vbox=0):
    """Random soup of all primitive kinds, some with rotate/scale/translate chains (n_moving / n_vsph / n_vbox
    add Moving_Spheres and volumes; they come last so the ids of the other primitives do not change)."""
    rng = np.random.default_rng(seed)
    s = pt.Scene(lib)
    s.set_textures([tex_const(0.6, 0.6, 0.6)])
    s.set_materials([material(pt.MAT_NORMAL, flags=0), material(pt.MAT_NORMAL, flags=1)])

    def rand_xf():
        if not xforms or rng.random() < 0.4:
            return np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
        m = mat_translate(rng.uniform(-3, 3, 3)) @ mat_rotate(rng.uniform(-90, 90), int(rng.integers(0, 3))) \
            @ mat_scale(rng.uniform(0.5, 1.5, 3))
        return m34(m)

    if n_sph:
        sph = [pt.Sphere(pt.F3(*rng.uniform(-5, 5, 3)), float(rng.uniform(0.2, 1.0))) for _ in range(n_sph)]
        s.add_spheres(sph, rng.integers(0, 2, n_sph), xforms=np.stack([rand_xf() for _ in range(n_sph)]))
    if n_rect:
        rc = []
        for _ in range(n_rect):
            a0, b0, k = rng.uniform(-5, 4, 3)
            rc.append(pt.Rect(a0, a0 + rng.uniform(0.3, 2), b0, b0 + rng.uniform(0.3, 2), k,
                              int(rng.integers(0, 3)), int(rng.integers(0, 2))))
        s.add_rects(rc, rng.integers(0, 2, n_rect), xforms=np.stack([rand_xf() for _ in range(n_rect)]))
    if n_box:
        bx = []
        for _ in range(n_box):
            lo = rng.uniform(-5, 4, 3)
            bx.append(pt.Box(pt.F3(*lo), pt.F3(*(lo + rng.uniform(0.2, 1.5, 3)))))
        s.add_boxes(bx, rng.integers(0, 2, n_box), xforms=np.stack([rand_xf() for _ in range(n_box)]))
    if n_tri:
        c = rng.uniform(-5, 5, (n_tri, 1, 3))
        v = (c + rng.uniform(-0.7, 0.7, (n_tri, 3, 3))).astype(np.float32).reshape(n_tri, 9)
        s.add_mesh(v, 0)
    if n_moving:
        ms = []
        for _ in range(n_moving):
            c0 = rng.uniform(-5, 5, 3)
            ms.append(pt.MovingSphere(pt.F3(*c0), pt.F3(*(c0 + rng.uniform(-1.5, 1.5, 3))), float(rng.uniform(0.2, 0.8)),
                                      float(rng.uniform(-0.5, 0.2)), float(rng.uniform(0.8, 1.5))))
        s.add_moving_spheres(ms, rng.integers(0, 2, n_moving), xforms=np.stack([rand_xf() for _ in range(n_moving)]))
    if n_vsph:
        vs = [pt.VolumeSphere(pt.F3(*rng.uniform(-5, 5, 3)), float(rng.uniform(0.5, 2.0)), float(rng.uniform(0.2, 3.0)))
              for _ in range(n_vsph)]
        s.add_volume_spheres(vs, rng.integers(0, 2, n_vsph), xforms=np.stack([rand_xf() for _ in range(n_vsph)]))
    if n_vbox:
        vb = []
        for _ in range(n_vbox):
            lo = rng.uniform(-5, 4, 3)
            vb.append(pt.VolumeBox(pt.F3(*lo), pt.F3(*(lo + rng.uniform(0.5, 2.5, 3))), float(rng.uniform(0.2, 3.0))))
        s.add_volume_boxes(vb, rng.integers(0, 2, n_vbox), xforms=np.stack([rand_xf() for _ in range(n_vbox)]))
    s.set_camera(make_camera([0, 1, -14], [0, 0, 0], [0, 1, 0], 50.0, W / H, 0.0, 10.0, 0.0, 1.0))
    s.set_miss(pt.Miss(pt.MISS_CONSTANT, pt.F3(0.1, 0.2, 0.3), pt.F3(0, 0, 0), -1, 1, 0))
    return s


def random_rays(n, seed=0, extent=6.0):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=pt.RAY_DTYPE)
    o = rng.uniform(-extent, extent, (n, 3))
    tgt = rng.uniform(-extent * 0.7, extent * 0.7, (n, 3))
    rays["origin"] = o
    rays["dir"] = (tgt - o)
    rays["tmin"] = 1e-3
    rays["tmax"] = 1e27
    return rays


def rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)))
