#!/usr/bin/env python
"""One process driving N devices through ONE handle (pt_renderer_create_multi): set-up time and steady-state rate of the
C4 scene, against a single-device renderer in the same process.  usage (GPU box, N GPUs visible): python scripts/multi_handle_probe.py N"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as h
pt = h.pt
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
W, H, SPP = 1920, 1080, 512
lib = pt.load()
sc = h.build_mesh_scene(708, W, H, lib)
sc.commit()
single = pt.Renderer(sc, W, H)
single.render(0, SPP, clear=True)
t0 = time.perf_counter(); single.render(0, SPP, clear=True); a = single.read_accum(); t1 = time.perf_counter()
print(f"1 device : {W*H*SPP/(t1-t0)/1e6:8.1f} Msamples/s wall incl. read-back ({single.stats().ms_total:.1f} ms on the device)")
t0 = time.perf_counter()
multi = pt.Renderer(sc, W, H, devices=list(range(N)))
t_setup = time.perf_counter() - t0
multi.render(0, SPP, clear=True); multi.read_accum()
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); multi.render(0, SPP, clear=True); b = multi.read_accum(); best = min(best, time.perf_counter() - t0)
st = multi.stats()
print(f"{N} devices: {W*H*SPP/best/1e6:8.1f} Msamples/s wall incl. merge + read-back (slowest device {st.ms_total:.1f} ms, "
      f"ncclReduce {st.ms_reduce:.3f} ms); set-up (replicas: upload + BVH build on {N-1} devices, {N} renderers, ncclCommInitAll) {t_setup:.2f} s")
print("sample counts ok:", bool((b[..., 3] == SPP).all()), " max |a-b| / spp:", float(np.abs(a - b).max() / SPP))
