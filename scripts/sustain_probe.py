#!/usr/bin/env python
"""Why is a single render after a pause ~3 % faster than back-to-back renders?  Same renderer, same job (C4 at 512 spp):
eight renders back to back, a pause, four more; then the scene re-committed (new BVH placement) and four more.
usage (GPU box): python scripts/sustain_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h
pt = h.pt
lib = pt.load()
W, H, SPP = 1920, 1080, 512
sc = h.build_mesh_scene(708, W, H, lib)
sc.commit()
r = pt.Renderer(sc, W, H)
r.render(0, 16, clear=True)
def run(n, label):
    t = []
    for _ in range(n):
        r.render(0, SPP, clear=True)
        t.append(round(r.stats().ms_total, 1))
    print(label, t, flush=True)
run(8, "back to back      ")
time.sleep(2.0)
run(4, "after a 2 s pause ")
for k in range(4):
    time.sleep(1.0)
    run(1, "1 s pause each    ")
sc.commit()
run(4, "re-committed scene")
# with an L2 flush (256 MiB write) between renders, like bench.py's timed loop
import torch
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
t = []
for _ in range(6):
    flush.zero_()
    torch.cuda.synchronize()
    r.render(0, SPP, clear=True)
    t.append(round(r.stats().ms_total, 1))
print("flush before each ", t, flush=True)
run(3, "no flush again    ")
