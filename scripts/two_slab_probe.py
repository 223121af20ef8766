#!/usr/bin/env python
"""Two renderers alive at once (two 14 GB path-pool slabs: the first and the second big allocation of the process), the
same scene and job rendered alternately: is one of the two places slower?  usage (GPU box): python scripts/two_slab_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h
pt = h.pt
lib = pt.load()
W, H, SPP = 1920, 1080, 512
sc = h.build_mesh_scene(708, W, H, lib)
sc.commit()
a = pt.Renderer(sc, W, H)
a.render(0, SPP, clear=True)
b = pt.Renderer(sc, W, H)
b.render(0, SPP, clear=True)
for rnd in range(3):
    for name, r in (("first slab ", a), ("second slab", b)):
        t = []
        for _ in range(2):
            r.render(0, SPP, clear=True)
            t.append(round(r.stats().ms_total, 1))
        print(name, t, flush=True)
