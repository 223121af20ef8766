"""Generate tests/golden/ref_cu_vectors.json from the REFERENCE's own intersection / bounds /
hit-record / light programs (programs/hitables/*.cu, programs/pdfs/*.cu).

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden_cu.py
It loads oracle/_ref/libref_cu.so — those .cu files compiled UNMODIFIED as host code against
oracle/ref_shim/optix_device.h by oracle/Makefile — and records input -> output vectors as uint32 bit
patterns; tests/test_oracle_cu_golden.py replays them through the oracle bit-exactly, with no reference tree.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers as h  # noqa: E402

pt = h.pt
KINDS = ["sphere", "rect", "box", "triangle", "moving_sphere", "volume_sphere", "volume_box"]


def bits(x):
    return [int(v) for v in np.asarray(x, dtype=np.float32).reshape(-1).view(np.uint32)]


def unbits(b):
    return np.asarray(b, dtype=np.uint32).view(np.float32)


def make_cases(rng, n_per_kind=40):
    """Random primitives + rays aimed roughly at them (unnormalised directions, ~2/3 hit)."""
    cases = []
    for kind in KINDS:
        for i in range(n_per_kind):
            c = rng.uniform(-4, 4, 3)
            ext = rng.uniform(0.3, 2.0, 3)
            if kind == "sphere":
                prm = [*c, ext[0]]
            elif kind == "rect":
                prm = [c[0], c[0] + ext[0], c[1], c[1] + ext[1], c[2], float(rng.integers(0, 3)), float(rng.integers(0, 2))]
            elif kind == "box":
                prm = [*c, *(c + ext)]
            elif kind == "triangle":
                v = (c[None, :] + rng.uniform(-1.5, 1.5, (3, 3))).reshape(-1)
                has_n, has_uv = float(i % 2), float((i // 2) % 2)
                nn = rng.normal(size=(3, 3)).reshape(-1)
                prm = [*v, has_n, *nn, has_uv, *rng.uniform(0, 1, 6)]
            elif kind == "moving_sphere":
                prm = [*c, *(c + rng.uniform(-1, 1, 3)), ext[0], float(rng.uniform(-0.3, 0.2)), float(rng.uniform(0.8, 1.4))]
            elif kind == "volume_sphere":
                prm = [*c, ext[0] + 0.5, float(rng.uniform(0.2, 3.0))]
            else:
                prm = [*c, *(c + ext + 0.5), float(rng.uniform(0.2, 3.0))]
            # aim point inside / near the primitive
            if kind == "rect":
                ax = int(prm[5])
                ab = [rng.uniform(prm[0] - 0.3, prm[1] + 0.3), rng.uniform(prm[2] - 0.3, prm[3] + 0.3)]
                tgt = {0: [prm[4], ab[0], ab[1]], 1: [ab[0], prm[4], ab[1]], 2: [ab[0], ab[1], prm[4]]}[ax]
            elif kind == "triangle":
                w = rng.dirichlet([1, 1, 1]) * rng.choice([1.0, 1.0, 1.4])
                tgt = w @ np.asarray(prm[:9]).reshape(3, 3)
            elif kind in ("box", "volume_box"):
                tgt = rng.uniform(np.asarray(prm[:3]) - 0.2, np.asarray(prm[3:6]) + 0.2)
            else:
                tgt = np.asarray(prm[:3]) + rng.normal(size=3) * prm[6 if kind == "moving_sphere" else 3] * 0.7
            inside = kind.startswith("volume") and i % 4 == 0
            o = np.asarray(tgt) + (rng.normal(size=3) * 0.1 if inside else rng.normal(size=3) * rng.uniform(3, 9))
            d = (np.asarray(tgt) - o) * rng.uniform(0.3, 2.5) if not inside else rng.normal(size=3)
            tmax = 1e27 if i % 5 else float(rng.uniform(0.5, 2.0))
            cases.append({"kind": kind, "prm": bits(prm), "ray": bits([*o, 1e-3, *d, tmax]),
                          "time": bits([rng.uniform(0, 1)]), "seed": int(rng.integers(0, 2**32))})
    return cases


def run_reference(r, case):
    prm = unbits(case["prm"]).copy()
    ray = unbits(case["ray"]).copy()
    hit3, rec, bnd = np.zeros(3, np.float32), np.zeros(14, np.float32), np.zeros(6, np.float32)
    hit = getattr(r, "refcu_" + case["kind"])(prm.ctypes.data, ray.ctypes.data, float(unbits(case["time"])[0]), case["vseed"],
                                              hit3.ctypes.data, rec.ctypes.data, bnd.ctypes.data)
    return int(hit), hit3, rec, bnd


def light_cases(rng, n=48):
    out = []
    for i in range(n):
        a0, b0, k = rng.uniform(-3, 3, 3)
        prm = [a0, a0 + rng.uniform(0.5, 2), b0, b0 + rng.uniform(0.5, 2), k]
        axis = int(rng.integers(0, 3))
        P = rng.uniform(-5, 5, 3)
        out.append({"light": "rect", "prm": bits(prm), "axis": axis, "P": bits(P), "seed": int(rng.integers(0, 2**32)),
                    "Wi_rand": bits(rng.normal(size=3))})
    for i in range(n):
        c = rng.uniform(-3, 3, 3)
        rad = float(rng.uniform(0.3, 1.5))
        P = c + rng.normal(size=3) * rng.uniform(2, 6)
        out.append({"light": "sphere", "prm": bits([*c, rad]), "axis": 0, "P": bits(P), "seed": int(rng.integers(0, 2**32)),
                    "Wi_rand": bits(rng.normal(size=3))})
    return out


def run_light_reference(r, c):
    prm = unbits(c["prm"]).copy()
    P, wr = pt.F3(*unbits(c["P"])), pt.F3(*unbits(c["Wi_rand"]))
    s = C.c_uint32(c["seed"])
    wi = pt.F3()
    if c["light"] == "rect":
        r.refcu_rect_sample(prm.ctypes.data, c["axis"], P, C.byref(s), wi)
        pdf_s = r.refcu_rect_pdf(prm.ctypes.data, c["axis"], P, wi)
        pdf_r = r.refcu_rect_pdf(prm.ctypes.data, c["axis"], P, wr)
    else:
        r.refcu_spherelight_sample(prm.ctypes.data, P, C.byref(s), wi)
        pdf_s = r.refcu_spherelight_pdf(prm.ctypes.data, P, wi)
        pdf_r = r.refcu_spherelight_pdf(prm.ctypes.data, P, wr)
    return list(wi), int(s.value), pdf_s, pdf_r


def main():
    r = h.ref_cu_probe()
    assert r is not None, "reference tree / probe not available"
    o = h.oracle()
    rng = np.random.default_rng(20261018)
    cases = make_cases(rng)
    for i, c in enumerate(cases):
        # the volume programs draw rnd(prd.seed); the oracle's pinned uniform is rnd(tea<4>(ray seed, prim id)),
        # so the probe's payload seed is that hash (prim id = the case's index within its kind, as the test builds it)
        c["prim"] = i % 40
        c["vseed"] = int(o.orc_volume_seed(c["seed"], c["prim"]))
        hit, hit3, rec, bnd = run_reference(r, c)
        c["hit"] = hit
        c["hit3"] = bits(hit3) if hit else []
        c["rec"] = bits(rec) if hit else []
        c["bounds"] = bits(bnd)
    lights = light_cases(rng)
    for c in lights:
        wi, seed_after, pdf_s, pdf_r = run_light_reference(r, c)
        c.update({"Wi": bits(wi), "seed_after": seed_after, "pdf_sampled": bits([pdf_s]), "pdf_random": bits([pdf_r])})
    out = {"note": "floats are uint32 bit patterns; generated by tests/golden/make_golden_cu.py from the reference's "
                   "hitables/*.cu and pdfs/*.cu compiled unmodified as host code",
           "prims": cases, "lights": lights}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cu_vectors.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    n_hit = sum(c["hit"] for c in cases)
    print("wrote", path, os.path.getsize(path), "bytes;", n_hit, "of", len(cases), "primitive cases hit")


if __name__ == "__main__":
    main()
