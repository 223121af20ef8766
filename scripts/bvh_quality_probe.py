#!/usr/bin/env python
"""CPU-side probe of BVH quality (dev tool): builds the C4-shaped mesh scene through the host simulation of
the builder (tests/hostsim, the same bvh_core.h bodies the GPU runs) and reports SAH cost, BVH8 node count and the
mean node / leaf fetches per ray for camera rays and for two generations of cosine-bounced (incoherent) rays.
Usage: python scripts/bvh_quality_probe.py [grid=708] [nrays_side=256]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as h

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 708
side = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sc = h.build_mesh_scene(grid, side, side)
orc = h.OracleScene(sc, mode=1)
cam = orc.camera_rays(side, side, 0)
b1 = orc.bounce_rays(cam, 7)
b2 = orc.bounce_rays(b1, 8)
t0 = time.time()
sim = h.SimScene(sc)
tb = time.time() - t0
info = sim.info()
print(f"grid {grid}: prims {2*grid*grid}  build {tb:.1f}s  nodes8 {info['n_nodes8']}  SAH {info['sah_before']:.2f} -> {info['sah_after']:.2f}")
for name, rays in (("camera", cam), ("bounce1", b1), ("bounce2", b2)):
    hits, cnt = sim.trace(rays)
    n = len(rays)
    print(f"  {name:8s} rays {n:7d}  nodes/ray {cnt[0]/n:6.2f}  leaves/ray {cnt[1]/n:5.2f}  hit {np.mean(hits['prim']>=0):.3f}")
