import sys, os, time
sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch, helpers as h
pt = h.pt
lib = pt.load()
for name, build, W, H, spp in (("c1", lambda: h.build_cornell(512, 512, lib), 512, 512, 64), ("c2", lambda: h.build_spheres(22, W=1200, H=800, lib=lib), 1200, 800, 256)):
    for it in range(3):
        s = build()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.commit(); torch.cuda.synchronize(); t1 = time.perf_counter()
        acc = torch.zeros(H, W, 4, device="cuda"); r = pt.Renderer(s, W, H, d_accum=acc.data_ptr()); torch.cuda.synchronize(); t2 = time.perf_counter()
        r.render(0, spp, clear=True); torch.cuda.synchronize(); t3 = time.perf_counter()
        out = acc.cpu().numpy(); t4 = time.perf_counter()
        pin = torch.empty(H, W, 4, pin_memory=True); t5 = time.perf_counter()
        pin.copy_(acc); torch.cuda.synchronize(); t6 = time.perf_counter()
        print(f"{name} it{it}: commit {1e3*(t1-t0):.2f} (build {s.bvh_info().build_ms:.2f}) renderer {1e3*(t2-t1):.2f} render {1e3*(t3-t2):.2f} (gpu {r.stats().ms_total:.2f}) d2h pageable {1e3*(t4-t3):.2f} pinned-alloc {1e3*(t5-t4):.2f} d2h pinned {1e3*(t6-t5):.2f}")
        r.close(); s.close()
