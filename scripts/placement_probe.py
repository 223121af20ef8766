#!/usr/bin/env python
"""Does the placement of the path-pool slab change the render time?  Same GPU, same scene, same job: the renderer (and its
14 GB slab) is destroyed and re-created between repetitions, with the process-level caches released so that the slab is a
fresh allocation each time.  usage (GPU box): python scripts/placement_probe.py [reps=8]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h
pt = h.pt
lib = pt.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
W, H, SPP = 1920, 1080, 128
sc = h.build_mesh_scene(708, W, H, lib)
sc.commit()
import torch
hold = []
for k in range(reps):
    r = pt.Renderer(sc, W, H)
    r.render(0, SPP, clear=True)
    t = []
    for _ in range(3):
        r.render(0, SPP, clear=True)
        t.append(r.stats().ms_total)
    print(f"slab {k}: {min(t):.2f} ms (runs {[round(x, 2) for x in t]}) = {W * H * SPP / min(t) / 1e3:.1f} Msamples/s", flush=True)
    r.close()
    lib.pt_release_caches()
    hold.append(torch.empty((64 + 37 * k) << 20, dtype=torch.uint8, device="cuda"))   # shift where the next slab lands
