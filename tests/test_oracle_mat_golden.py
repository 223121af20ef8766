"""The oracle against the reference's OWN closest-hit programs and Direct_Light.

oracle/ref_mat_probe.cpp compiles programs/materials/{lambertian,metal,dielectric,diffuse_light,oren_nayar,
torrance_sparrow,ashikhmin_shirley,normal_shader,isotropic}.cu — each with light_sample.cuh's Direct_Light, the
BSDF headers and the light programs of programs/pdfs/ — UNMODIFIED as host code.  One case = one closest-hit
invocation on a real hit of a test scene with a random incoming payload; compared bit for bit: the payload after
the program (RNG state = number of draws consumed, throughput, radiance, next origin and direction, scatter event,
specular flag) and the shadow ray Direct_Light traced (or that it traced none).

Two things are inputs rather than outputs, and stated as such:
  * what the shadow ray answers (the visibility rule is the oracle's pin; the reference leaves it to OptiX);
  * the hit record, taken from the oracle - pinned against the reference's Get_HitRecord programs by
    tests/test_oracle_cu_golden.py.
The reference programs are compiled by g++, which evaluates two rnd(seed) inside one argument list right to left;
the oracle build that must match them is liboracle_r2l.so (-DORC_ARGS_R2L: same source, that one ordering flipped).
"""
import json
import os
import sys

import numpy as np
import pytest

import helpers as h

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_mat as G  # noqa: E402

pt = h.pt
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_mat_vectors.json")


def nb(x):
    a = np.asarray(x, dtype=np.float32).reshape(-1).copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return [int(v) for v in a.view(np.uint32)]


def compare(run, c, ref_prd, ref_sh):
    prd, sh = run.run_oracle(c)
    name = (c["scene"], h.MAT_PROGRAMS[c["mat"]])
    assert nb(prd[:1]) == nb(ref_prd[:1]), (name, "RNG state after the program (draws consumed)")
    assert nb(prd) == nb(ref_prd), (name, "payload", prd, ref_prd)
    assert int(sh[0]) == int(ref_sh[0]), (name, "number of shadow rays traced")
    if sh[0] > 0:
        assert int(sh[7]) == c["in_shadow"], (name, "the scene's shadow answer changed: regenerate the golden file")
        assert nb(sh[1:7]) == nb(ref_sh[1:7]), (name, "shadow ray")


def test_oracle_reproduces_reference_closest_hit_programs_golden():
    gold = json.load(open(GOLD))
    run = G.Runner()
    seen = set()
    for c in gold["cases"]:
        compare(run, c, G.unbits(c["prd_out"]), G.unbits(c["shadow_out"]))
        seen.add(c["mat"])
    assert seen == set(range(9)), seen           # all nine closest-hit programs are covered
    assert sum(1 for c in gold["cases"] if G.unbits(c["shadow_out"])[0] > 0) > 100


def test_oracle_reproduces_reference_closest_hit_programs_live():
    r = h.ref_mat_probe()
    if r is None:
        pytest.skip("reference tree not present (the recorded vectors cover this)")
    rng = np.random.default_rng(4242)
    run = G.Runner()
    cases = G.make_cases(rng, per_scene=90)
    n_shadow = 0
    for c in cases:
        _, sh = run.run_oracle(c)
        for answer in ((int(sh[7]),) if sh[0] > 0 else (0, 1)):   # no shadow ray: the answer must not matter
            c["in_shadow"] = answer
            ref_prd, ref_sh, c["mat"] = run.run_reference(r, c, answer)
            compare(run, c, ref_prd, ref_sh)
        n_shadow += int(sh[0] > 0)
    assert len(cases) > 400 and n_shadow > 150


def test_standard_and_r2l_oracle_differ_only_in_draw_order():
    """Same scene, same frames: the two oracle builds consume the same NUMBER of draws everywhere (identical
    ray counts), and produce the same image where no two-draw argument list is involved (Normal_Shader pass)."""
    sc = h.build_cornell(24, 24)
    a, sa = h.OracleScene(sc).render(24, 24, 0, 2)
    b, sb = h.OracleScene(sc, variant="_r2l").render(24, 24, 0, 2)
    assert (a[..., 3] == 2).all() and (b[..., 3] == 2).all()
    assert abs(a[..., :3].mean() - b[..., :3].mean()) < 0.05 * a[..., :3].mean()
    ns = h.build_random_prims(3)
    a, _ = h.OracleScene(ns).render(24, 24, 0, 1)
    b, _ = h.OracleScene(ns, variant="_r2l").render(24, 24, 0, 1)
    assert np.array_equal(a, b)
