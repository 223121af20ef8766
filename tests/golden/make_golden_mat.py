"""Generate tests/golden/ref_mat_vectors.json from the REFERENCE's own closest-hit programs
(programs/materials/*.cu with light_sample.cuh's Direct_Light).

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden_mat.py
oracle/_ref/libref_mat.so is those programs compiled UNMODIFIED as host code (oracle/ref_mat_probe.cpp).  A case
is one closest-hit invocation: a hit of a real ray in one of the test scenes, an incoming payload (seed, time,
throughput, radiance) and the answer of the shadow ray; the outputs are the payload after the program and the
shadow ray it traced.  tests/test_oracle_mat_golden.py replays the cases through the oracle (the build with
g++'s argument evaluation order, liboracle_r2l.so) bit-exactly.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers as h  # noqa: E402

pt = h.pt


def bits(x):
    return [int(v) for v in np.asarray(x, dtype=np.float32).reshape(-1).view(np.uint32)]


def unbits(b):
    return np.asarray(b, dtype=np.uint32).view(np.float32)


def two_light_cornell(W, H):
    sc = h.build_cornell(W, H)
    sc.add_rects([pt.Rect(100, 200, 100, 200, 1.0, pt.Y_AXIS, 0)], [3])
    sc.set_lights([pt.Light(0, pt.Y_AXIS, 213, 343, 227, 332, 554, pt.F3(7, 7, 7)),
                   pt.Light(0, pt.Y_AXIS, 100, 200, 100, 200, 1.0, pt.F3(3, 3, 3))])
    return sc


SCENES = {
    "cornell": lambda: h.build_cornell(32, 32),
    "zoo": lambda: h.build_material_zoo(32, 32),
    "two_lights": lambda: two_light_cornell(32, 32),
    "sphere_light": lambda: h.build_sphere_light_scene(32, 32, two=True),
    "smoke": lambda: h.build_smoke_scene(32, 32, fog=True),
    "spheres": lambda: h.build_spheres(8, W=32, H=32),
    "soup_normals": lambda: h.build_random_prims(3),
}


def scene_tables(sc):
    d = sc.desc()
    mats = [d.materials[i] for i in range(d.n_materials)]
    texs = [d.textures[i] for i in range(d.n_textures)]
    prims = [d.prims[i] for i in range(d.n_prims)]
    lights = [d.lights[i] for i in range(d.n_lights)]
    larr = [float(len(lights))]
    for L in lights:
        larr += [float(L.type), float(L.axis), L.a0, L.a1, L.b0, L.b1, L.k, *L.emission]
    return mats, texs, prims, np.asarray(larr + [0.0] * 10, dtype=np.float32)


def make_cases(rng, per_scene=60):
    """inputs only: scene name, ray, hit (prim, t, u, v), payload"""
    cases = []
    for name, build in SCENES.items():
        sc = build()
        orc = h.OracleScene(sc, mode=1, variant="_r2l")
        rays = orc.camera_rays(32, 32, 0)
        b1 = orc.bounce_rays(rays, 7)
        rays = np.concatenate([rays, b1, orc.bounce_rays(b1, 9)])
        rays = rays[rays["tmax"] > rays["tmin"]]
        hits, _ = orc.trace(rays)          # ray-set convention: time = camera.time0, seed = ray index
        idx = np.flatnonzero(hits["prim"] >= 0)
        # spread the picks over the materials that occur
        mats, _, prims, _ = scene_tables(sc)
        by_mat = {}
        for i in idx:
            by_mat.setdefault(mats[prims[hits["prim"][i]].material].type, []).append(i)
        picks = []
        for mt, lst in by_mat.items():
            take = max(6, per_scene // len(by_mat))
            picks += list(rng.choice(lst, size=min(take, len(lst)), replace=False))
        for i in picks:
            r = rays[i]
            cases.append({"scene": name, "ray": bits([*r["origin"], r["tmin"], *r["dir"], r["tmax"]]),
                          "ray_index": int(i), "prim": int(hits["prim"][i]),
                          "hit": bits([hits["t"][i], hits["u"][i], hits["v"][i]]),
                          "prd_in": [int(rng.integers(0, 2**32))] + bits([sc.desc().camera.time0, *rng.uniform(0.05, 1.0, 3),
                                                                        *rng.uniform(0.0, 2.0, 3)])})
    return cases


class Runner:
    """Runs cases through the r2l oracle and (when available) the reference programs."""

    def __init__(self):
        self.o = h.oracle("_r2l")
        self.scenes = {}

    def scene(self, name):
        if name not in self.scenes:
            sc = SCENES[name]()
            self.scenes[name] = (sc, h.OracleScene(sc, mode=1, variant="_r2l"), scene_tables(sc))
        return self.scenes[name]

    def inputs(self, c):
        sc, orc, (mats, texs, prims, larr) = self.scene(c["scene"])
        r = unbits(c["ray"])
        ray = pt.Ray(pt.F3(*r[:3]), r[3], pt.F3(*r[4:7]), r[7])
        t, u, v = unbits(c["hit"])
        prd = np.zeros(16, np.float32)
        prd.view(np.uint32)[0] = c["prd_in"][0]
        prd[1:8] = unbits(c["prd_in"][1:])
        return sc, orc, mats, texs, prims, larr, ray, float(t), float(u), float(v), prd

    def run_oracle(self, c):
        sc, orc, mats, texs, prims, larr, ray, t, u, v, prd = self.inputs(c)
        sh = np.zeros(8, np.float32)
        self.o.orc_closest_hit_probe(orc.h, C.byref(ray), c["prim"], t, u, v, prd.ctypes.data, sh.ctypes.data)
        return prd, sh

    def run_reference(self, r, c, in_shadow):
        sc, orc, mats, texs, prims, larr, ray, t, u, v, prd = self.inputs(c)
        m = mats[prims[c["prim"]].material]
        hit3, rec = np.zeros(3, np.float32), np.zeros(14, np.float32)
        # the hit record the program's Get_HitRecord callable returns: the oracle's, which
        # tests/test_oracle_cu_golden.py pins against the reference's own Get_HitRecord programs
        ok = self.o.orc_intersect_prim_ex(orc.h, c["prim"], C.byref(ray), float(prd[1]), c["ray_index"], hit3.ctypes.data,
                                          rec.ctypes.data)
        assert ok and bits(hit3) == c["hit"], "scene builders changed: regenerate the golden file"
        tex6 = np.zeros(6, np.float32)
        for k, tid in enumerate((m.tex0, m.tex1)):
            if tid >= 0:
                assert texs[tid].type == pt.TEX_CONSTANT, "material cases use constant textures"
                tex6[3 * k:3 * k + 3] = list(texs[tid].color)
        mp3 = np.asarray([m.p0, m.p1, float(m.flags & 1)], np.float32)
        ray8 = unbits(c["ray"]).copy()
        sh = np.zeros(8, np.float32)
        getattr(r, "refmat_" + h.MAT_PROGRAMS[m.type])(rec.ctypes.data, ray8.ctypes.data, t, tex6.ctypes.data, mp3.ctypes.data,
                                                       larr.ctypes.data, int(in_shadow), prd.ctypes.data, sh.ctypes.data)
        return prd, sh, m.type


def main():
    r = h.ref_mat_probe()
    assert r is not None, "reference tree / probe not available"
    rng = np.random.default_rng(20261019)
    run = Runner()
    cases = make_cases(rng)
    for c in cases:
        _, sh = run.run_oracle(c)                       # only to learn what the shadow ray answers in this scene
        c["in_shadow"] = int(sh[7]) if sh[0] > 0 else int(rng.integers(0, 2))
        prd, shr, mt = run.run_reference(r, c, c["in_shadow"])
        c["mat"] = int(mt)
        c["prd_out"] = bits(prd)
        c["shadow_out"] = bits(shr)
    out = {"note": "floats are uint32 bit patterns; generated by tests/golden/make_golden_mat.py from the reference's "
                   "materials/*.cu compiled unmodified as host code", "cases": cases}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_mat_vectors.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    from collections import Counter
    print("wrote", path, os.path.getsize(path), "bytes;", len(cases), "cases; per material", dict(Counter(c["mat"] for c in cases)),
          "shadow rays traced in", sum(1 for c in cases if unbits(c["shadow_out"])[0] > 0))


if __name__ == "__main__":
    main()
