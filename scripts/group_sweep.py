#!/usr/bin/env python
"""A/B of the camera-sample enumeration (k_generate): frames per group F = 1 ... 32 (PTB200_FRAME_GROUP), one process,
scenes built once.  Prints Msamples/s per workload and F, and checks that the image is the F = 1 image up to the order
of the float additions.   usage (GPU box): [SWEEP_F="1 32 256"] python scripts/group_sweep.py [c4:128 c2:64 c1:64 c3:64 c4:512:16777216 (= pool)]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import helpers as h  # noqa: E402

pt = h.pt
lib = pt.load()
specs = sys.argv[1:] or ["c4:128", "c2:64", "c1:64", "c3:64"]
FS = [int(f) for f in os.environ.get("SWEEP_F", "1 2 4 8 16 32 64 128 256").split()]
VAR = os.environ.get("SWEEP_VAR", "PTB200_FRAME_GROUP")   # any switch the library reads per render
for spec in specs:
    name, spp, *rest = spec.split(":")
    spp = int(spp)
    pool = int(rest[0]) if rest else 0
    desc, W, H, _ = bench.WORKLOADS[name]
    sc = bench.build_scene(h, name, W, H, lib)
    sc.commit()
    r = pt.Renderer(sc, W, H)
    if pool:
        r.set_pool(pool)
    ref = None
    line = [f"{name:7s} {W}x{H}@{spp}" + (f" pool {pool}" if pool else "")]
    for F in FS:
        os.environ[VAR] = str(F)
        r.render(0, spp, clear=True)
        ms = []
        for _ in range(3):
            r.render(0, spp, clear=True)
            ms.append(r.stats().ms_total)
        a = r.read_accum()
        assert (a[..., 3] == spp).all(), "sample count"
        if ref is None:
            ref = a
        err = float(np.abs(a[..., :3] - ref[..., :3]).max() / max(1e-9, float(ref[..., :3].max())))
        line.append(f"F={F}: {W * H * spp / (min(ms) * 1e-3) / 1e6:7.1f} (d {err:.1e})")
    print("  ".join(line), flush=True)
    r.close()
    sc.close()
