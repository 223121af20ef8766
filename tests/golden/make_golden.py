"""Generate tests/golden/ref_vectors.json from the REFERENCE's own header code.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
It loads oracle/_ref/libref_probe.so (the reference's programs/*.cuh compiled
unmodified against oracle/ref_shim by oracle/Makefile) and records
input -> output vectors as uint32 bit patterns, so the comparison in
tests/test_oracle_golden.py is bit-exact and needs no reference tree.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers as h  # noqa: E402

pt = h.pt
F3 = pt.F3


def bits(x):
    return [int(v) for v in np.asarray(x, dtype=np.float32).reshape(-1).view(np.uint32)]


def nrm(v):
    return v / np.linalg.norm(v)


def main():
    r = h.ref_probe()
    assert r is not None, "reference tree / probe not available"
    rng = np.random.default_rng(20261017)
    out = {"arg_order_l2r": int(r.ref_arg_order()), "note": "floats are uint32 bit patterns"}
    out["tea64"] = [[a, b, int(r.ref_tea64(a, b))] for a, b in
                    [(5, 0), (0, 5), (0, 0), (123456, 7), (4294967295, 3), (512 * 300 + 17, 63)]]
    seqs = []
    for seed in (460419465, 1, 0, 4294967295):
        s = C.c_uint32(seed)
        seqs.append([seed, bits([r.ref_rnd(C.byref(s)) for _ in range(8)]), int(s.value)])
    out["rnd"] = seqs
    samp = []
    for fn in ("ref_random_in_unit_disk", "ref_random_in_unit_sphere", "ref_random_on_unit_sphere"):
        for seed in rng.integers(0, 2**32, 16):
            s = C.c_uint32(int(seed))
            o = F3()
            getattr(r, fn)(C.byref(s), o)
            samp.append([fn[4:], int(seed), bits(list(o)), int(s.value)])
    out["sampling"] = samp
    mats = [(0, 0.0, 0.0), (4, 0.2605, 0.1302), (5, 0.04, 0.04), (5, 0.04, 0.0225), (6, 100.0, 10.0),
            (6, 50.0, 50.0), (3, 0.0, 0.0)]
    ev, sm = [], []
    # Isotropic (materials/isotropic.cuh) was added later: it draws from its own generator, after the
    # original materials, so the earlier vectors stay byte-identical when the file is regenerated
    rng_main = rng
    for mat, p0, p1 in mats + [(8, 0.0, 0.0)]:
        rng = np.random.default_rng(8) if mat == 8 else rng_main
        for k in range(64 if mat != 8 else 32):
            N = nrm(rng.normal(size=3))
            Wo = nrm(rng.normal(size=3))
            if k % 2 == 0 and np.dot(Wo, N) < 0:
                Wo = -Wo
            Wi = nrm(rng.normal(size=3)) * rng.uniform(0.5, 3.0)
            col, col2 = rng.uniform(0, 1, 3), rng.uniform(0, 1, 3)
            f, pdf = F3(), C.c_float()
            args = (mat, F3(*col), F3(*col2), p0, p1, F3(0, 0, 0), F3(*Wo), F3(*Wi), F3(*N))
            r.ref_bsdf_eval(*args, f, C.byref(pdf))
            ev.append({"mat": mat, "p": bits([p0, p1]), "col": bits(col), "col2": bits(col2), "Wo": bits(Wo),
                       "Wi": bits(Wi), "N": bits(N), "f": bits(list(f)), "pdf": bits([pdf.value])})
            seed = int(rng.integers(0, 2**32))
            s = C.c_uint32(seed)
            wi = F3()
            r.ref_bsdf_sample(mat, F3(*col), F3(*col2), p0, p1, F3(0, 0, 0), F3(*Wo), F3(*N), C.byref(s), wi)
            sm.append({"mat": mat, "p": bits([p0, p1]), "col": bits(col), "col2": bits(col2), "Wo": bits(Wo),
                       "N": bits(N), "seed": seed, "Wi": bits(list(wi)), "seed_after": int(s.value)})
    rng = rng_main
    out["bsdf_eval"] = ev
    out["bsdf_sample"] = sm
    misc = []
    for _ in range(48):
        c, ri = float(rng.uniform(0, 1)), float(rng.uniform(1.0, 2.5))
        v, n = nrm(rng.normal(size=3)) * rng.uniform(0.5, 2), nrm(rng.normal(size=3))
        eta = float(rng.choice([1.5, 1 / 1.5, 1.0, 2.4]))
        ro = F3()
        ok = r.ref_refract(F3(*v), F3(*n), eta, ro)
        o = nrm(rng.normal(size=3))
        if o[1] < 0 and rng.random() < 0.7:
            o[1] = -o[1]
        u1, u2 = float(rng.random()), float(rng.random())
        nu, nv = float(rng.choice([0.04, 0.09, 0.25])), float(rng.choice([0.04, 0.0225]))
        go = F3()
        r.ref_ggx_sample(F3(*o), u1, u2, nu, nv, go)
        H = nrm(rng.normal(size=3))
        wi = nrm(rng.normal(size=3))
        misc.append({"c": bits([c, ri]), "schlick": bits([r.ref_schlick(c, ri)]), "v": bits(v), "n": bits(n),
                     "eta": bits([eta]), "refract_ok": int(ok), "refract": bits(list(ro)), "o": bits(o),
                     "u": bits([u1, u2]), "nunv": bits([nu, nv]), "ggx_sample": bits(list(go)), "H": bits(H),
                     "ggx_d": bits([r.ref_ggx_d(F3(*H), nu, nv)]), "wi": bits(wi),
                     "ggx_g": bits([r.ref_ggx_g(F3(*o), F3(*wi), nu, nv)])})
    out["misc"] = misc
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
