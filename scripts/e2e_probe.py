"""Dev: where the end-to-end time of one step goes (commit / renderer / render / read-back)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch, helpers as h
pt = h.pt
lib = pt.load()
which = sys.argv[1:] or ["c1", "c2", "c4"]
cfg = {"c1": (lambda: h.build_cornell(512, 512, lib), 512, 512, 64),
       "c2": (lambda: h.build_spheres(22, W=1200, H=800, lib=lib), 1200, 800, 256),
       "c4": (lambda: h.build_mesh_scene(708, 1920, 1080, lib), 1920, 1080, 64),
       "c4_10m": (lambda: h.build_mesh_scene(2237, 1920, 1080, lib), 1920, 1080, 64)}
for name in which:
    build, W, H, spp = cfg[name]
    pin = torch.empty(H, W, 4, pin_memory=True)
    for it in range(4):
        s = build()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.commit(); torch.cuda.synchronize(); t1 = time.perf_counter()
        acc = torch.zeros(H, W, 4, device="cuda"); torch.cuda.synchronize(); t1b = time.perf_counter()
        r = pt.Renderer(s, W, H, d_accum=acc.data_ptr()); torch.cuda.synchronize(); t2 = time.perf_counter()
        r.render(0, spp, clear=True); torch.cuda.synchronize(); t3 = time.perf_counter()
        pin.copy_(acc); torch.cuda.synchronize(); t4 = time.perf_counter()
        d = s.desc(); t5 = time.perf_counter()
        r.close(); t6 = time.perf_counter(); s.close(); t7 = time.perf_counter()
        print(f"{name} it{it}: commit {1e3*(t1-t0):.2f} (build {s.bvh_info().build_ms if False else 0:.2f}) zeros {1e3*(t1b-t1):.2f} renderer {1e3*(t2-t1b):.2f} "
              f"render {1e3*(t3-t2):.2f} (gpu {r.stats().ms_total if False else 0:.2f}) d2h pinned {1e3*(t4-t3):.2f} desc {1e3*(t5-t4):.2f} r.close {1e3*(t6-t5):.2f} s.close {1e3*(t7-t6):.2f}", flush=True)
