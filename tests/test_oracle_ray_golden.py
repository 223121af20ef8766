"""The oracle against the reference's OWN ray-generation program (programs/raygen.cu: camera ray and seed, the
color() loop with its scatter events / Russian roulette / depth cut-off, de_nan, accumulation, display pixel) and
miss programs (programs/miss.cu), compiled UNMODIFIED as host code (oracle/ref_ray_probe.cpp).

Inside the reference's loop every rtTrace is answered by the oracle's single-trace entry point
(orc_trace_radiance: closest hit + that material's closest-hit program, or the miss program) — those pieces are
pinned against the reference separately (test_oracle_cu_golden.py, test_oracle_mat_golden.py).  Required: the
accumulation buffer and the 8-bit display buffer of the reference's renderPixel equal the oracle's own render of
the same frames BIT FOR BIT.  (g++ argument order: the oracle build is liboracle_r2l.so, see
test_oracle_mat_golden.py.)
"""
import json
import os
import sys

import numpy as np
import pytest

import helpers as h

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_ray as G  # noqa: E402

pt = h.pt
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ray_vectors.json")


def nb(x):
    a = np.asarray(x, dtype=np.float32).reshape(-1).copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return a.view(np.uint32)


def check_render(acc_ref, disp_ref, name, W, H, frames, depth, russian):
    acc, disp = G.oracle_render(G.build(name, W, H), W, H, frames, depth, russian)
    assert (acc[..., 3] == frames).all()
    same = nb(acc) == nb(acc_ref)
    assert same.all(), (name, f"{(~same).sum()} of {same.size} accumulation values differ")
    assert np.array_equal(disp.reshape(-1), np.asarray(disp_ref, dtype=np.uint8).reshape(-1)), (name, "display buffer")
    assert acc[..., :3].max() > 0


def miss_scene(c):
    s = pt.Scene()
    s.set_textures([h.tex_const(0.5)])
    s.set_materials([h.material(pt.MAT_NORMAL)])
    s.add_spheres([pt.Sphere(pt.F3(0, 0, 0), 1.0)], [0])
    img = -1
    if c["type"] >= 2:
        t = G.unbits(c["texel"])
        img = s.add_image(np.asarray([[[t[0], t[1], t[2], 1.0]]], dtype=np.float32))
    s.set_miss(pt.Miss(c["type"], pt.F3(*G.unbits(c["c1"])), pt.F3(*G.unbits(c["c2"])), img, c["spherical"], 0))
    return s


def check_miss(c, thr_ref, uv_ref, event_ref):
    sc = miss_scene(c)
    orc = h.OracleScene(sc, mode=1)
    thr, uv = pt.F3(), np.zeros(2, np.float32)
    orc.o.orc_miss_probe(orc.h, pt.F3(*G.unbits(c["dir"])), pt.F3(*G.unbits(c["thr"])), thr, uv.ctypes.data)
    assert event_ref == 3                                              # rayMissed (prd.cuh:24-29)
    if c["type"] < 2:
        assert list(nb(list(thr))) == list(nb(thr_ref)), ("miss colour", c["type"])
    else:
        # the sampler's coordinates are what the program computes; the texel fetch itself is the sampler's
        assert list(nb(uv)) == list(nb(uv_ref)), ("miss texture coordinates", c["type"], c["spherical"], uv, uv_ref)


def test_oracle_reproduces_reference_raygen_and_miss_golden():
    gold = json.load(open(GOLD))
    for g in gold["renders"]:
        check_render(G.unbits(g["acc"]), g["disp"], g["scene"], g["W"], g["H"], g["frames"], g["max_depth"], g["russian"])
    for c in gold["misses"]:
        check_miss(c, G.unbits(c["thr_out"]), G.unbits(c["uv"]), c["event"])


def test_oracle_reproduces_reference_raygen_and_miss_live():
    r = h.ref_ray_probe()
    if r is None:
        pytest.skip("reference tree not present (the recorded vectors cover this)")
    for name, W, H, frames, depth, russian in G.RENDERS + [("zoo", 20, 14, 2, 11, 1), ("spheres", 20, 20, 1, 3, 0)]:
        acc, disp = G.reference_render(r, G.build(name, W, H), W, H, frames, depth, russian)
        check_render(acc, disp.reshape(-1), name, W, H, frames, depth, russian)
    for c in G.miss_cases(np.random.default_rng(99), 64):
        check_miss(c, *G.run_miss_reference(r, c))
