"""C-ABI surface (CPU): the library loads without a GPU, exports every symbol
include/ptb200.h declares, validates arguments, and refuses to compute without CUDA."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as h

pt = h.pt
HDR = os.path.join(h.ROOT, "include", "ptb200.h")


def declared_symbols():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert set(syms) == set(pt.ABI_SYMBOLS), set(syms) ^ set(pt.ABI_SYMBOLS)
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.pt_abi_version() == 3


def test_limits_match_header():
    """PT_MAX_PIXELS / PT_BAND_PIXELS / PT_MAX_DEPTH of the header are what the Python binding exports"""
    src = open(HDR).read()

    def define(name):
        m = re.search(r"#define\s+%s\s+\(?\s*(\d+)\s*(?:<<\s*(\d+))?\s*\)?" % name, src)
        assert m, name
        return int(m.group(1)) << int(m.group(2) or 0)

    assert define("PT_MAX_PIXELS") == pt.MAX_PIXELS == 1 << 28
    assert define("PT_BAND_PIXELS") == pt.BAND_PIXELS == 1 << 23
    assert define("PT_MAX_DEPTH") == pt.MAX_DEPTH == 255
    assert define("PTB200_ABI_VERSION") == 3


def test_struct_layout_matches_header():
    # sizes the C header implies (plain 4-byte fields, no padding)
    assert C.sizeof(pt.Sphere) == 16 and C.sizeof(pt.Rect) == 28 and C.sizeof(pt.Box) == 24
    assert C.sizeof(pt.Material) == 24 and C.sizeof(pt.Texture) == 24 and C.sizeof(pt.Light) == 40
    assert C.sizeof(pt.Camera) == 96 and C.sizeof(pt.Miss) == 40
    assert C.sizeof(pt.Ray) == 32 and C.sizeof(pt.Hit) == 16 and C.sizeof(pt.PrimRef) == 16
    assert C.sizeof(pt.Xform) == 96
    assert pt.RAY_DTYPE.itemsize == 32 and pt.HIT_DTYPE.itemsize == 16


def test_scene_building_and_desc(lib):
    s = h.build_cornell(16, 16, lib)
    d = s.desc()
    assert d.n_prims == 7 and d.n_rects == 6 and d.n_boxes == 1 and d.n_xforms == 1
    assert d.n_materials == 4 and d.n_lights == 1
    # prim ids follow insertion order; the rotated box carries the only transform
    assert [d.prims[i].type for i in range(7)] == [pt.PRIM_RECT] * 6 + [pt.PRIM_BOX]
    assert d.prims[6].xform == 0 and all(d.prims[i].xform == -1 for i in range(6))
    x = d.xforms[0]
    o2w = np.array(list(x.o2w)).reshape(3, 4)
    w2o = np.array(list(x.w2o)).reshape(3, 4)
    M = np.vstack([o2w, [0, 0, 0, 1]]) @ np.vstack([w2o, [0, 0, 0, 1]])
    assert np.allclose(M, np.eye(4), atol=1e-4)
    # T * R: the box origin maps to the translation (host_includes/hitables.hpp:518-523)
    assert np.allclose(o2w[:, 3], [265, 0, 295])
    assert abs(o2w[0, 0] - np.cos(np.radians(15))) < 1e-6 and abs(o2w[0, 2] - np.sin(np.radians(15))) < 1e-6


def test_mesh_transform_applied_on_host(lib):
    s = pt.Scene(lib)
    s.set_textures([h.tex_const(0.5)])
    s.set_materials([h.material(pt.MAT_LAMBERTIAN, 0)])
    v = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], dtype=np.float32)
    n = np.array([[0, 0, 1] * 3], dtype=np.float32)
    xf = h.m34(h.mat_translate([1, 2, 3]) @ h.mat_scale([2, 2, 2]))
    first = s.add_mesh(v, 0, n=n, xform=xf)
    assert first == 0
    d = s.desc()
    assert [d.tri_v[i] for i in range(9)] == [1, 2, 3, 3, 2, 3, 1, 4, 3]
    assert abs(d.tri_n[2] - 0.5) < 1e-7  # inverse-transpose of a 2x scale


def test_argument_validation(lib):
    s = pt.Scene(lib)
    with pytest.raises(pt.PtError):
        s.add_rects([pt.Rect(0, 1, 0, 1, 0, 7, 0)], [0])          # invalid axis
    with pytest.raises(pt.PtError):
        s.set_materials([pt.Material(99, 0, 0, 0, 0, 0)])          # unknown material
    with pytest.raises(pt.PtError):
        s.set_miss(pt.Miss(17, pt.F3(0, 0, 0), pt.F3(0, 0, 0), -1, 1, 0))
    with pytest.raises(pt.PtError):
        s.add_spheres([pt.Sphere(pt.F3(0, 0, 0), 1)], [0], xforms=np.zeros((1, 12), np.float32))  # singular
    with pytest.raises(pt.PtError):
        s.bvh_info()                                               # not committed


@pytest.mark.skipif(h.has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    s = h.build_cornell(8, 8, lib)
    with pytest.raises(pt.PtError) as e:
        s.commit()
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)
    rays = np.zeros(4, dtype=pt.RAY_DTYPE)
    with pytest.raises(pt.PtError):
        s.trace(rays)


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(pt.PtError):
        pt.load(str(tmp_path / "nope.so"))


def test_frame_splits(lib):
    """pt_split_frames / pt_split_frames_weighted (host logic of the multi-GPU sample split): every frame is assigned
    exactly once, shares are contiguous, equal shares differ by at most one, weighted shares follow the weights."""
    for world in (1, 2, 3, 8):
        for total in (0, 1, 7, 512, 4096 + 3):
            got = [pt.split_frames(total, r, world, 10, lib) for r in range(world)]
            frames = [f for a, n in got for f in range(a, a + n)]
            assert frames == list(range(10, 10 + total))
            ns = [n for _, n in got]
            assert max(ns) - min(ns) <= 1
    import importlib
    mg = importlib.import_module("optix_path_tracer_b200.multigpu")
    for world in (1, 2, 4, 8):
        for r in range(world):
            assert pt.split_frames(4099, r, world, 10, lib) == mg.split_frames(4099, r, world, 10)   # the Python twin
    w = [1.0, 0.97, 1.02, 0.99, 1.0, 1.01, 0.98, 1.03]
    got = [pt.split_frames_weighted(512, r, 8, w, 0, lib) for r in range(8)]
    assert [f for a, n in got for f in range(a, a + n)] == list(range(512))
    ns = [n for _, n in got]
    assert ns[7] == max(ns) and ns[1] == min(ns)
    assert all(abs(n - 512 * wi / sum(w)) < 1.0 for n, wi in zip(ns, w))
    assert [pt.split_frames_weighted(512, r, 4, [1, 1, 1, 1], 0, lib) for r in range(4)] == \
        [pt.split_frames(512, r, 4, 0, lib) for r in range(4)]
    assert [pt.split_frames_weighted(5, r, 2, [0.0, 1.0], 0, lib) for r in range(2)] == [(0, 3), (3, 2)]   # unusable weights
    with pytest.raises(pt.PtError):
        pt.split_frames(8, 2, 2, 0, lib)
