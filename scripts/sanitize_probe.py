#!/usr/bin/env python
"""Small renders that touch every kernel (builder incl. top-down partition and collapse DP, env alias tables, wavefront
iterations, drain, ray-set queries, display conversion) - meant to be run under compute-sanitizer:
  compute-sanitizer --tool memcheck  python scripts/sanitize_probe.py
  compute-sanitizer --tool racecheck python scripts/sanitize_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as h
pt = h.pt
lib = pt.load()
W = H = 48
for name, sc in (("cornell", h.build_cornell(W, H, lib)), ("spheres", h.build_spheres(22, W=W, H=H, lib=lib)),
                 ("mesh_env_nee", h.build_mesh_scene(40, W, H, lib, env_nee=True)), ("smoke", h.build_smoke_scene(W, H, lib)),
                 ("motion", h.build_motion_scene(W, H, lib)), ("zoo", h.build_material_zoo(W, H, lib))):
    sc.commit()
    r = pt.Renderer(sc, W, H)
    r.set_pool(4096)
    r.render(0, 3)
    a = r.read_accum()
    d = r.read_display()
    rays = h.random_rays(2000, 1)
    hits = sc.trace(rays)
    assert (a[..., 3] == 3).all() and np.isfinite(a).all()
    print(name, "ok", r.stats().iterations, "iterations, drained", r.stats().drained_paths, "hits", int((hits["prim"] >= 0).sum()))
    r.close(); sc.close()
print("done")
