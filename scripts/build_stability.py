#!/usr/bin/env python
"""BVH build and end-to-end stability at the top of configuration 4's range (GPU): N fresh scenes of the 10 M-triangle
mesh (and the 1 M one), each committed and rendered; prints GPU build time, wall time of the commit, wall time of
commit + render + read-back.  The first repetition pays the process's one-off costs (CUDA module load, the
build arena, the path-pool slab) and is listed separately.
usage (GPU box): python scripts/build_stability.py [reps=6] > gpurun_out/r02_build_stability.md"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as h
pt = h.pt
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
lib = pt.load()
W, H, SPP = 1920, 1080, 32
print("# Round 2: GPU BVH build and end-to-end stability (one B200)\n")
for grid, label in ((708, "1 M triangles"), (2237, "10 M triangles")):
    tris, _ = h.heightfield_mesh(grid)
    rows = []
    for k in range(reps):
        s = pt.Scene(lib)
        s.set_textures([h.tex_const(0.7, 0.7, 0.7)])
        s.set_materials([h.material(pt.MAT_LAMBERTIAN, 0)])
        s.add_mesh(tris, 0)
        s.set_camera(h.make_camera([0, 4.5, -9], [0, 0, 0], [0, 1, 0], 40.0, W / H, 0.0, 10.0))
        img = s.add_image(h.synthetic_hdri())
        s.set_miss(pt.Miss(pt.MISS_HDR, pt.F3(0, 0, 0), pt.F3(0, 0, 0), img, 1, 0))
        t0 = time.perf_counter()
        s.commit()
        t1 = time.perf_counter()
        r = pt.Renderer(s, W, H)
        r.render(0, SPP, clear=True)
        out = r.read_accum()
        t2 = time.perf_counter()
        info = s.bvh_info()
        rows.append((info.build_ms, (t1 - t0) * 1e3, (t2 - t0) * 1e3, r.stats().ms_total, info.sah_cost_refined, info.n_bvh8_nodes))
        r.close(); s.close()
    print(f"## {label} ({2 * grid * grid} triangles), {W}x{H} @ {SPP} spp\n")
    print("| run | GPU build ms | commit wall ms (upload + build) | commit + render + read-back wall ms | render ms | SAH | BVH8 nodes |")
    print("|---|---|---|---|---|---|---|")
    for k, row in enumerate(rows):
        print(f"| {k}{' (first in process)' if k == 0 else ''} | {row[0]:.1f} | {row[1]:.1f} | {row[2]:.1f} | {row[3]:.1f} | {row[4]:.2f} | {row[5]} |")
    a = np.array([r_[:4] for r_ in rows[1:]])
    spread = (a.max(axis=0) - a.min(axis=0)) / a.mean(axis=0) * 100
    print(f"\nruns 1..{reps - 1}: mean build {a[:,0].mean():.1f} ms (spread {spread[0]:.1f} %), commit {a[:,1].mean():.1f} ms "
          f"(spread {spread[1]:.1f} %), end to end {a[:,2].mean():.1f} ms (spread {spread[2]:.1f} %), "
          f"e2e {W * H * SPP / (a[:,2].mean() * 1e-3) / 1e6:.0f} Msamples/s\n")
