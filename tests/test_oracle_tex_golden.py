"""The oracle's texture evaluation against the reference's OWN texture programs (programs/textures/*.cu compiled
UNMODIFIED as host code, oracle/ref_tex_probe.cpp): constant, checker, gradient, vector (incl. out-of-range index)
and Perlin noise (turbulence, the axis switch that falls through) — bit-exact, from recorded vectors and live."""
import json
import os
import sys

import numpy as np
import pytest

import helpers as h

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_tex as G  # noqa: E402

pt = h.pt
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tex_vectors.json")


def oracle_value(c):
    """one-texture scene in the ABI's representation (include/ptb200.h PT_TEX_*), evaluated by the oracle"""
    prm = G.full_params(c)
    s = pt.Scene()
    texs, texvec, texdata, root = [], [], [], 0
    if c["tex"] == "constant":
        texs = [h.tex_const(*prm[:3])]
    elif c["tex"] == "checker":
        texs = [pt.Texture(pt.TEX_CHECKER, pt.F3(0, 0, 0), 1, 2), h.tex_const(*prm[:3]), h.tex_const(*prm[3:6])]
    elif c["tex"] == "gradient":
        texs = [pt.Texture(pt.TEX_GRADIENT, pt.F3(0, 0, 0), 0, 0)]
        texdata = list(prm[:9])
    elif c["tex"] == "vector":
        k = c["n"]
        texs = [pt.Texture(pt.TEX_VECTOR, pt.F3(0, 0, 0), 0, k)] + [h.tex_const(*prm[3 * i:3 * i + 3]) for i in range(k)]
        texvec = list(range(1, k + 1))
    else:
        texs = [pt.Texture(pt.TEX_NOISE, pt.F3(float(prm[0]), 0, 0), 0, int(prm[1]))]
        texdata = list(prm[2:])
    s.set_textures(texs)
    if texvec:
        s.set_texture_vector(texvec)
    if texdata:
        s.set_texture_data(np.asarray(texdata, dtype=np.float32))
    s.set_materials([h.material(pt.MAT_LAMBERTIAN, 0)])
    s.add_spheres([pt.Sphere(pt.F3(0, 0, 0), 1.0)], [0])
    orc = h.OracleScene(s, mode=1)
    q = G.unbits(c["uvp"])
    out = pt.F3()
    orc.o.orc_sample_texture(orc.h, root, float(q[0]), float(q[1]), pt.F3(*q[2:5]), c["index"], out)
    return list(out)


def test_oracle_reproduces_reference_texture_programs_golden():
    gold = json.load(open(GOLD))
    for c in gold["cases"]:
        assert G.bits(oracle_value(c)) == c["rgb"], (c["tex"], oracle_value(c), list(G.unbits(c["rgb"])))
    assert {c["tex"] for c in gold["cases"]} == {"constant", "checker", "gradient", "vector", "noise"}


def test_oracle_reproduces_reference_texture_programs_live():
    r = h.ref_tex_probe()
    if r is None:
        pytest.skip("reference tree not present (the recorded vectors cover this)")
    for c in G.make_cases(np.random.default_rng(5), 30):
        assert G.bits(oracle_value(c)) == G.bits(G.run_reference(r, c)), c["tex"]
