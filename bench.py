#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json metric: Msamples/s and
Mrays/s; one JSON line on rank 0).

A step = one full render of the workload (default C4: procedural 1 M-triangle mesh under a
synthetic HDRI, 1920x1080 @ 512 spp - the north-star configuration): camera-ray generation, BVH8
closest-hit + shadow traversal, material shading / MIS / Russian roulette, float4 accumulation
and, for N > 1, the NCCL reduce of the accumulation buffers to rank 0.  N > 1 is STRONG scaling:
the workload's spp are split over the ranks (frames are independent: seed = tea<64>(pixel, frame),
programs/raygen.cu:150), scene + BVH replicated.  --workload c1|c2|c3|c4_8k|c4_10m|c5 select the others.

  value    whole-job Msamples/s, scene and BVH resident in HBM (device-timed)
  e2e      same metric through the C ABI from HOST buffers: scene upload (H2D), GPU BVH
           build, render, accumulation-buffer read-back (D2H) all inside the timed region
  roofline closest-hit traversal kernel on the TIMED configuration: algorithmic bytes / CUDA-event time vs measured
           HBM peak, with what actually bounds the kernel (ncu counters of the current build, profiles/ncu_traffic.json)
  incoherent_secondary  Mrays/s of closest-hit queries on uniform random secondary rays (triangle scenes)
  other_configs  C1 / C2 / C3 device-timed at full size, C4 with HDRI importance sampling on (N = 1 only)
  cpu_baseline  the CPU oracle on the host cores, bounded sample of the same workload

`--impl reference` times the reference's own programs on the host cores: compiled unmodified inside the
emulated OptiX pipeline of oracle/ref_render.cpp (oracle/_ref/libref_render.so), one process per core; the
oracle port when that library is absent.  (The reference itself needs OptiX 6 + an RT-capable GPU.)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (description, W, H, spp)
    "c1": ("C1 Cornell box (6 rects + rotated box, Lambertian + area light, MIS) 512x512 @ 64 spp", 512, 512, 64),
    "c2": ("C2 RTOW final scene: 488 spheres (Lambertian/metal/dielectric) 1200x800 @ 256 spp", 1200, 800, 256),
    "c3": ("C3 Cornell + glass sphere + Oren-Nayar/Torrance-Sparrow/Ashikhmin-Shirley, RR, 1024x1024 @ 1024 spp",
           1024, 1024, 1024),
    "c4": ("C4 procedural 1M-triangle mesh + synthetic HDRI env map 1920x1080 @ 512 spp", 1920, 1080, 512),
    "c4_10m": ("C4 procedural 10M-triangle mesh + synthetic HDRI env map 1920x1080 @ 512 spp (scene arrays built "
               "directly, not through the OBJ file)", 1920, 1080, 512),
    # the same scene and the same number of samples as C4 on 16 times the pixels: 33 M pixels are four bands of rows
    # inside pt_render (the path record indexes 2^23 pixels); shows what banded rendering costs
    "c4_8k": ("C4 scene (1M-triangle mesh + HDRI) at 7680x4320 @ 32 spp: four bands of rows inside pt_render", 7680, 4320, 32),
    # the multi-GPU configuration: 4096 spp are SPLIT across the ranks (strong scaling)
    "c5": ("C5 procedural 1M-triangle mesh + HDRI 3840x2160 @ 4096 spp, sample-split across ranks", 3840, 2160, 4096),
}


FACADE_IDS = {"c1": 2, "c2": 0, "c3": 5, "c4": 6, "c4_8k": 6, "c5": 6}   # scene functions in host_includes/scenes.hpp


class _StdoutToStderr:
    """The facade's scene functions printf progress like the reference does; bench.py must print
    exactly one JSON line on stdout, so C-level stdout is pointed at stderr while they run."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def build_scene(h, name, W, H, lib=None):
    with _StdoutToStderr():
        return _build_scene(h, name, W, H, lib)


def _build_scene(h, name, W, H, lib=None):
    """Scenes come from the C++ host facade's scene functions (InOneWeekend, Cornell, ... — the
    reference's scene-building surface); the numpy builders are the fallback if it is not built."""
    pt = h.pt
    if name in FACADE_IDS and os.path.exists(pt.HOST_LIB_PATH) and not os.environ.get("PTB200_NO_FACADE"):
        import facade_dump
        return facade_dump.build(pt.HOST_LIB_PATH, FACADE_IDS[name], W, H, 0)
    if name == "c1":
        return h.build_cornell(W, H, lib)
    if name == "c2":
        return h.build_spheres(22, seed=0, W=W, H=H, lib=lib)
    if name == "c3":
        return h.build_material_zoo(W, H, lib)
    if name in ("c4", "c4_8k", "c5"):
        return h.build_mesh_scene(708, W, H, lib)
    if name == "c4_10m":
        return h.build_mesh_scene(2237, W, H, lib)
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz, self.power = index, False, [], set(), None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(float(os.environ.get("PTB200_CLOCK_PERIOD", "0.5")))

    def result(self):
        import statistics
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "sm_mhz_mean": statistics.fmean(self.sm) if self.sm else None, "sm_mhz_min": min(self.sm) if self.sm else None,
                "power_w_mean": statistics.fmean(self.power) if self.power else None,
                "power_w_max": max(self.power) if self.power else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def scene_host_bytes(d):
    """bytes of the flat scene description a commit uploads (H2D)"""
    n = d.n_prims * 16 + d.n_spheres * 16 + d.n_rects * 28 + d.n_boxes * 24 + d.n_tris * 36
    n += d.n_xforms * 96 + d.n_materials * 24 + d.n_textures * 24 + d.n_texvec * 4 + d.n_lights * 40
    if d.tri_n:
        n += d.n_tris * 36
    if d.tri_uv:
        n += d.n_tris * 24
    if d.tri_tex:
        n += d.n_tris * 4
    for i in range(d.n_images):
        n += d.images[i].width * d.images[i].height * 16
    return n


_JSON_FD = None


def emit_json(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, line)
    else:
        os.write(_JSON_FD, line)


_REF_WORKER = None  # (ReferenceRenderer, W, H) inherited by the forked band workers


def _ref_band(task):
    """one row band x frame range of the reference renderer; returns the band's radiance sum (so that the work
    cannot be skipped) -- runs in a forked worker, one rendering thread per process (the reference's program
    variables are process globals in the emulation)"""
    import numpy as np
    ren, W, H = _REF_WORKER
    y0, y1, first, n = task
    acc = np.zeros((H, W, 4), dtype=np.float32)
    ren.render(W, H, first, n, rows=(y0, y1), acc=acc)
    return float(acc[H - y1:H - y0, :, :3].sum(dtype=np.float64))   # row 0 is the bottom scanline


def run_reference(args, rank, world):
    """Reference arm.  The reference has no CPU code path of its own (OptiX 6 + an RT-capable GPU), so what is
    timed is its own programs compiled unmodified for the host inside the emulated pipeline of
    oracle/ref_render.cpp (oracle/_ref/libref_render.so, kind "reference"), one process per host core over row
    bands; when that library was not built (no /root/reference at build time) the oracle port is timed instead
    (kind "port", OpenMP over scanlines)."""
    if rank != 0:
        return
    import helpers as h
    desc, W, H, spp = WORKLOADS[args.workload]
    sc = build_scene(h, args.workload, W, H)
    base = {
        "impl": "reference", "metric": "Msamples/s", "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "width": W, "height": H, "spp": spp, "max_depth": 50, "russian": True},
    }
    budget = float(os.environ.get("PTB200_REFERENCE_SECONDS", "90")) / max(1, args.steps + args.warmup)

    # the oracle port, always (a few seconds): the stronger CPU baseline, reported next to the reference's own code
    orc = h.OracleScene(sc)
    t0 = time.perf_counter()
    _, st = orc.render(W, H, 100000, 1)
    t1 = time.perf_counter() - t0
    use_ref = h.ReferenceRenderer.lib() is not None and not os.environ.get("PTB200_REFERENCE_PORT")
    if not use_ref:
        spp_step = int(max(1, min(spp, budget / max(t1, 1e-3))))
        for w in range(args.warmup):
            orc.render(W, H, 200000 + w * spp_step, spp_step)
        tot_s, rays = 0.0, 0
        for k in range(args.steps):
            t0 = time.perf_counter()
            _, st = orc.render(W, H, k * spp_step, spp_step)
            tot_s += time.perf_counter() - t0
            rays += st.closest_rays + st.shadow_rays
        val = W * H * spp_step * args.steps / tot_s / 1e6
        sample = f"{W}x{H} @ {spp_step} spp per step (of {spp}), all {st.threads} host threads, OpenMP over scanlines"
        base.update({
            "value": val, "ms_per_step": tot_s / args.steps * 1e3, "mrays_per_s": rays / tot_s / 1e6,
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": st.threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "oracle/_ref/libref_render.so absent: the oracle port (CPU restatement of the reference's programs) is timed",
        })
        emit_json(base)
        return
    n_port = int(max(1, min(spp, 8.0 / max(t1, 1e-3))))
    t0 = time.perf_counter()
    _, st = orc.render(W, H, 300000, n_port)
    t_port = time.perf_counter() - t0
    port = {"value": W * H * n_port / t_port / 1e6, "unit": "Msamples/s", "cores": st.threads, "kind": "port",
            "sample": f"{W}x{H} @ {n_port} spp, OpenMP over scanlines"}

    import multiprocessing as mp
    global _REF_WORKER
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    _REF_WORKER = (h.ReferenceRenderer(sc), W, H)   # acceleration structure built once, inherited through fork
    band = max(1, H // (cores * 8))                  # ~8 bands per process: dynamic balance over uneven rows
    bands = [(y, min(H, y + band)) for y in range(0, H, band)]
    with mp.get_context("fork").Pool(cores) as pool:
        def frame_pass(first, n):
            t0 = time.perf_counter()
            tot = sum(pool.imap_unordered(_ref_band, [(y0, y1, first, n) for y0, y1 in bands]))
            return time.perf_counter() - t0, tot
        t1, _ = frame_pass(100000, 1)
        spp_step = int(max(1, min(spp, budget / max(t1, 1e-3))))
        for w in range(args.warmup):
            frame_pass(200000 + w * spp_step, spp_step)
        tot_s, checksum = 0.0, 0.0
        for k in range(args.steps):
            dt, c = frame_pass(k * spp_step, spp_step)
            tot_s += dt
            checksum += c
    val = W * H * spp_step * args.steps / tot_s / 1e6
    sample = (f"{W}x{H} @ {spp_step} spp per step (of {spp}); the reference's own programs compiled unmodified for the "
              f"host (oracle/ref_render.cpp), {cores} processes over {len(bands)} row bands")
    base.update({
        "value": val, "ms_per_step": tot_s / args.steps * 1e3,
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "reference", "sample": sample},
        "oracle_port": port,
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_radiance": checksum / (W * H * 3.0 * spp_step * args.steps),
        "note": "the reference needs OptiX 6 + an RT-capable GPU and has no CPU path; its programs run here inside an "
                "emulated pipeline (images bit-identical to the oracle, tests/test_reference_render.py); oracle_port "
                "is the faster CPU restatement on the same cores",
    })
    emit_json(base)


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _ncu_evidence(workload):
    """counters of the dominant kernel from the committed ncu capture of the current build
    (profiles/ncu_traffic.json, regenerated by scripts/ncu_traffic.py, stamped with the commit it was taken at)"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(workload)
    except Exception:
        return None


def measure_config(h, pt, lib, torch, name, steps=3, warm=1):
    """device-timed Msamples/s of one other BASELINE configuration on this GPU (scene resident, full size)"""
    desc, W, H, spp = WORKLOADS[name]
    sc = build_scene(h, name, W, H, lib)
    sc.commit()
    r = pt.Renderer(sc, W, H)
    for _ in range(warm):
        r.render(0, spp, clear=True)
    ms, rays = 0.0, 0
    for _ in range(steps):
        r.render(0, spp, clear=True)
        st = r.stats()
        ms += st.ms_total
        rays += st.closest_rays + st.shadow_rays
    out = {"workload": desc, "msamples_per_s": float(W) * H * spp * steps / (ms * 1e-3) / 1e6,
           "mrays_per_s": rays / (ms * 1e-3) / 1e6, "ms_per_step": ms / steps, "steps": steps}
    r.close()
    sc.close()
    return out


def incoherent_rays(sc, lib, torch):
    """north_star target: incoherent secondary rays on the triangle scene.  Uniformly random directions from the
    first-hit points of the camera rays (origins scattered over the whole mesh, directions uncorrelated), traced
    through pt_trace_rays_device on resident buffers."""
    g = torch.Generator(device="cuda").manual_seed(1)
    n_r = 1 << 22
    cam = sc.desc().camera
    o = torch.tensor(list(cam.origin), device="cuda")
    llc, hor, ver = (torch.tensor(list(v), device="cuda") for v in (cam.lower_left_corner, cam.horizontal, cam.vertical))
    uv = torch.rand(n_r, 2, device="cuda", generator=g)
    d = llc + uv[:, :1] * hor + uv[:, 1:] * ver - o
    cam_rays = torch.zeros(n_r, 8, device="cuda")
    cam_rays[:, 0:3], cam_rays[:, 3], cam_rays[:, 4:7], cam_rays[:, 7] = o, 1e-6, d, 1e27
    hits = torch.zeros(n_r, 4, device="cuda")
    lib.pt_trace_rays_device(sc.h, cam_rays.data_ptr(), n_r, hits.data_ptr(), None)
    torch.cuda.synchronize()
    t = hits[:, 0:1]
    hit = (hits.view(torch.int32)[:, 1] >= 0)
    P = (o + t * d)[hit]
    n_s = P.shape[0]
    dirs = torch.randn(n_s, 3, device="cuda", generator=g)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    sec = torch.zeros(n_s, 8, device="cuda")
    sec[:, 0:3], sec[:, 3], sec[:, 4:7], sec[:, 7] = P, 1e-3, dirs, 1e27
    hits2 = torch.zeros(n_s, 4, device="cuda")
    for _ in range(3):
        lib.pt_trace_rays_device(sc.h, sec.data_ptr(), n_s, hits2.data_ptr(), None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        lib.pt_trace_rays_device(sc.h, sec.data_ptr(), n_s, hits2.data_ptr(), None)
    e1.record()
    e1.synchronize()
    return {"rays": int(n_s), "mrays_per_s": n_s * reps / (e0.elapsed_time(e1) * 1e-3) / 1e6,
            "hit_fraction": float((hits2.view(torch.int32)[:, 1] >= 0).float().mean()),
            "what": "uniform random directions from first-hit points, closest-hit queries, resident buffers"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the workload's spp are SPLIT over the ranks (default); weak = every rank renders all of them")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (dev only; invalidates the metric)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--env-nee", action="store_true", help="C4/C5 with HDRI importance sampling switched on (not the reference estimator)")
    ap.add_argument("--pool", type=int, default=0, help="path-pool size (dev only; 0 = library default)")
    ap.add_argument("--static-split", action="store_true",
                    help="N > 1: keep the equal split of the frames (default: shares follow the ranks' warm-up rates)")
    args = ap.parse_args()
    # stdout carries exactly ONE line: the JSON.  Libraries write there on their own (NCCL prints its version
    # banner at communicator creation, the facade's scene functions printf progress like the reference), so
    # file descriptor 1 points at stderr for the whole run and the JSON goes to the saved descriptor.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # torchrun exports OMP_NUM_THREADS=1; the CPU baseline (rank 0, N = 1 only) uses every core it may run on
    try:
        host_cores = len(os.sched_getaffinity(0))
    except Exception:
        host_cores = os.cpu_count() or 1
    if rank == 0:
        os.environ["OMP_NUM_THREADS"] = str(host_cores)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import helpers as h
    pt = h.pt

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = pt.load()
    desc, W, H, spp = WORKLOADS[args.workload]
    if args.spp:
        spp = args.spp
    warm = max(3, args.warmup)
    mesh_scene = args.workload in ("c4", "c4_8k", "c4_10m", "c5")

    def make_scene():
        s_ = build_scene(h, args.workload, W, H, lib)
        if args.env_nee and mesh_scene:
            m = s_.desc().miss
            m.env_nee = 1
            s_.set_miss(m)
        return s_

    t0 = time.perf_counter()
    sc = make_scene()
    host_build_s = time.perf_counter() - t0
    sc.commit()
    info = sc.bvh_info()
    r = pt.Renderer(sc, W, H)
    if args.pool:
        r.set_pool(args.pool)
    # The only collective of the path - the sum-reduce of the float4 accumulation buffers - runs INSIDE the library
    # (pt_comm_reduce: ncclReduce on the renderer's stream); torch.distributed only carries the 128-byte id.
    comm = None
    if world > 1:
        idt = torch.zeros(pt.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(pt.Comm.unique_id(lib)), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        comm = pt.Comm(bytes(idt.cpu().numpy().tobytes()), rank, world, lib)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    strong = args.scaling == "strong"
    if strong:   # the spp of the workload are split over the ranks
        total_spp = spp
        first, spp = pt.split_frames(total_spp, rank, world, lib=lib)
    else:        # weak scaling: every rank renders the full spp with its own frame indices
        first, total_spp = rank * spp, spp * world

    def step():
        """one pass of the hot path over one batch; returns (render ms, reduce ms) on the device"""
        r.render(first, spp, clear=True)
        ms = r.stats().ms_total
        red = 0.0
        if comm is not None:
            comm.reduce(r, 0)
            red = r.stats().ms_reduce
        return ms, red

    for _ in range(warm):
        step()
        flush.zero_()
    # GPUs of one box differ by 2-3 % in speed and a sample split ends when its slowest rank does: after the warm-up
    # the shares follow the ranks' measured rates (pt_split_frames_weighted; same frames, same total, every frame
    # rendered exactly once).  --static-split keeps the equal split.
    shares = [spp] * world
    if world > 1:
        rate = torch.tensor([spp / max(r.stats().ms_total, 1e-6)], device="cuda", dtype=torch.float32)
        rates = [torch.zeros_like(rate) for _ in range(world)]
        dist.all_gather(rates, rate)
        weights = [float(x) for x in rates]
        if strong and not args.static_split:
            first, spp = pt.split_frames_weighted(total_spp, rank, world, weights, lib=lib)
            shares = [pt.split_frames_weighted(total_spp, k, world, weights, lib=lib)[1] for k in range(world)]
            step()                      # one more warm-up step with the final shares
            flush.zero_()
        elif strong:
            shares = [pt.split_frames(total_spp, k, world, lib=lib)[1] for k in range(world)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    wall0 = time.perf_counter()
    ms_dev, ms_red, launches, rays = 0.0, 0.0, 0, 0
    for _ in range(args.steps):
        a, b = step()
        ms_dev += a + b
        ms_red += b
        st = r.stats()
        launches += st.kernel_launches + (1 if comm is not None else 0)
        rays += st.closest_rays + st.shadow_rays
        flush.zero_()  # L2 flush between timed iterations (outside the device-timed interval)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    sampler.stop_flag = True
    sampler.join()
    t = torch.tensor([ms_dev, ms_red, float(rays)], device="cuda", dtype=torch.float64)
    ms_dev_min = ms_dev
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tmin = t.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_dev, ms_red, rays, ms_dev_min = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tmin[0])
    samples_total = float(W) * H * total_spp * args.steps
    value = samples_total / (ms_dev * 1e-3) / 1e6
    final = r.read_accum() if rank == 0 else None
    if rank == 0:
        from optix_path_tracer_b200 import multigpu
        assert multigpu.check_sample_count(final, total_spp), "accumulation .w integrity check failed"

    # ---- e2e: host buffers -> C ABI -> host result -------------------------------------------------
    e2e_steps = max(1, min(args.steps, 3))
    # the result lands in PINNED host memory (a pageable destination made this 15 MB read-back 7.4 ms instead of
    # 0.35 ms on C2: the driver stages pageable copies at ~2 GB/s)
    host_out = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)
    h2d = 0

    e2e_parts = {"commit_ms": 0.0, "create_ms": 0.0, "render_device_ms": 0.0, "reduce_and_readback_ms": 0.0}

    def e2e_step(record=False):
        nonlocal h2d
        s2 = make_scene()                                   # host-side flat arrays (outside the timed region)
        t0 = time.perf_counter()
        s2.commit()                                         # H2D of the scene + GPU BVH build
        t1 = time.perf_counter()
        r2 = pt.Renderer(s2, W, H)
        t2 = time.perf_counter()
        r2.render(first, spp, clear=True)
        t3 = time.perf_counter()
        if comm is not None:
            comm.reduce(r2, 0)
        if rank == 0:                                       # D2H of the float4 accumulation buffer
            _check_rc(lib, lib.pt_read_accum(r2.h, host_out.data_ptr()))
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if record:
            e2e_parts["commit_ms"] += (t1 - t0) * 1e3
            e2e_parts["create_ms"] += (t2 - t1) * 1e3
            e2e_parts["render_device_ms"] += r2.stats().ms_total
            e2e_parts["reduce_and_readback_ms"] += (t4 - t3) * 1e3
        h2d = scene_host_bytes(s2.desc())
        r2.close()
        s2.close()
        return t4 - t0

    e2e_step()
    if world > 1:
        dist.barrier()
    t_e2e = 0.0
    for _ in range(e2e_steps):
        t_e2e += e2e_step(record=True)
    te = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = float(W) * H * total_spp * e2e_steps / float(te[0]) / 1e6

    incoherent = incoherent_rays(sc, lib, torch) if rank == 0 and mesh_scene else None

    # ---- roofline of the dominant kernel (closest-hit traversal) on the TIMED configuration: one instrumented pass
    # (fetch counters) and one per-kernel-class timing pass, both at this rank's full spp ----
    roof = None
    cpu_base = None
    others = None
    if rank == 0:
        r.set_flags(pt.FLAG_COUNT_FETCHES)
        r.render(first, spp, clear=True)
        sc_ = r.stats()
        n_node = sc_.nodes_fetched / max(1, sc_.closest_rays + sc_.shadow_rays)
        n_leaf = sc_.leaves_fetched / max(1, sc_.closest_rays + sc_.shadow_rays)
        leaf_bytes = 48 if mesh_scene else 32     # 48 B triangle / 32 B analytic primitive
        b_ray = 32 + 16 + 80 * n_node + leaf_bytes * n_leaf
        r.set_flags(pt.FLAG_TIME_KERNELS)
        r.render(first, spp, clear=True)
        stt = r.stats()
        r.set_flags(0)
        cls_sum = stt.ms_trace_closest + stt.ms_trace_shadow + stt.ms_shade + stt.ms_generate
        share = {"trace_closest": stt.ms_trace_closest / cls_sum, "trace_shadow": stt.ms_trace_shadow / cls_sum,
                 "shade": stt.ms_shade / cls_sum, "generate": stt.ms_generate / cls_sum}
        ach = stt.closest_rays * b_ray / (stt.ms_trace_closest * 1e-3) / 1e9
        peaks = _peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ev = _ncu_evidence(args.workload) or {}
        rays_per_launch = stt.closest_rays / max(1, stt.iterations)
        traffic = ev["dram_bytes"] / ev["rays"] * rays_per_launch if ev.get("dram_bytes") and ev.get("rays") else None
        roof = {"bound": "alu_issue", "roof": "hbm", "kernel": "k_trace_deep<closest>" if info.n_prims >= 65536 else "k_trace<closest>",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "why_bound": "ncu of this build: the BVH is served from L1/L2, DRAM traffic is a fraction of the algorithmic "
                             "bytes; the ALU pipe and the issue slots of the node test limit the kernel, not HBM",
                "issue_active_pct": ev.get("issue_active_pct"), "alu_pipe_pct": ev.get("alu_pipe_pct"),
                "lanes_per_inst": ev.get("lanes_per_inst"), "dram_pct_of_peak": ev.get("dram_pct"),
                "ncu_commit": ev.get("commit"),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6.65 TB/s (of fallback)",
                "bytes_per_ray": b_ray, "nodes_per_ray": n_node, "leaves_per_ray": n_leaf,
                "traversal": ("flat: <= 16 primitives, closest-hit rays test every leaf record; a record is ONE warp-uniform "
                              "broadcast fetch for 32 rays (counted once per warp), so this kernel is compute-bound, not on the "
                              "memory roofline" if info.n_prims <= 16 else "bvh8"),
                "measured_on": f"the timed configuration ({W}x{H} @ {spp} spp on this rank), per-kernel-class CUDA events",
                "closest_rays_per_step": int(stt.closest_rays), "trace_closest_ms_per_step": stt.ms_trace_closest,
                "rays_per_launch": rays_per_launch, "launch_ms": stt.ms_trace_closest / max(1, stt.iterations),
                "kernel_share": share,
                "check": "achieved = closest_rays_per_step * bytes_per_ray / trace_closest_ms_per_step; "
                         "trace_closest_ms_per_step ~= ms_per_step * kernel_share.trace_closest (the timing pass "
                         "synchronises per iteration, so its sum of classes runs a few % above the graph-replayed step)"}
        if world == 1 and not args.no_other_configs:
            others = {}
            for nm in ("c1", "c2", "c3", "c4"):
                if nm != args.workload:
                    others[nm] = measure_config(h, pt, lib, torch, nm)
            if mesh_scene and not args.env_nee:   # the same workload with HDRI importance sampling on (NEW estimator)
                s3 = build_scene(h, args.workload, W, H, lib)
                m = s3.desc().miss
                m.env_nee = 1
                s3.set_miss(m)
                s3.commit()
                r3 = pt.Renderer(s3, W, H)
                r3.render(0, spp, clear=True)
                r3.render(0, spp, clear=True)
                st3 = r3.stats()
                others[args.workload + "_env_nee"] = {
                    "workload": desc + " with alias-table HDRI importance sampling + MIS (pt_miss.env_nee = 1)",
                    "msamples_per_s": float(W) * H * spp / (st3.ms_total * 1e-3) / 1e6,
                    "mrays_per_s": (st3.closest_rays + st3.shadow_rays) / (st3.ms_total * 1e-3) / 1e6,
                    "ms_per_step": st3.ms_total, "steps": 1}
                r3.close()
                s3.close()
        if world == 1 and not args.no_cpu_baseline:
            orc = h.OracleScene(sc)
            t0 = time.perf_counter()
            orc.render(W, H, 100000, 1)
            t1 = time.perf_counter() - t0
            n_spp = int(max(1, min(spp, 15.0 / max(t1, 1e-3))))
            t0 = time.perf_counter()
            ref_acc, ost = orc.render(W, H, first, n_spp)
            dt = time.perf_counter() - t0
            cpu_base = {"value": W * H * n_spp / dt / 1e6, "unit": "Msamples/s", "cores": ost.threads, "kind": "port",
                        "sample": f"{W}x{H} @ {n_spp} spp of the same workload (oracle, OpenMP over scanlines)",
                        "mrays_per_s": (ost.closest_rays + ost.shadow_rays) / dt / 1e6}
            # converged-image agreement on the sample (independent of the timing)
            gi = np.sqrt(np.clip(final[..., :3] / total_spp, 0, 1))
            oi = np.sqrt(np.clip(ref_acc[..., :3] / n_spp, 0, 1))
            cpu_base["rmse_vs_gpu_gamma"] = h.rmse(gi, oi)

    if rank == 0:
        out = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "width": W, "height": H, "spp": total_spp, "spp_per_rank": shares if world > 1 else spp,
                       "split": ("rate-weighted after warm-up (pt_split_frames_weighted)" if world > 1 and strong and not args.static_split
                                 else "equal (pt_split_frames)"), "max_depth": 50,
                       "russian": True, "env_nee": bool(args.env_nee),
                       "l2": "flushed between timed steps (256 MiB write); path pool >> L2",
                       "prims": info.n_prims, "bvh8_nodes": info.n_bvh8_nodes,
                       "parallelism": f"sample-split x{world}: frames split over the ranks, scene + BVH replicated, one "
                                      f"ncclReduce of the float4 accumulation buffers inside the library (pt_comm_reduce)"},
            "mrays_per_s": rays / (ms_dev * 1e-3) / 1e6,
            "incoherent_secondary": incoherent,
            "reduce_ms_per_step": ms_red / args.steps,
            "rank_ms_per_step": {"max": ms_dev / args.steps, "min": ms_dev_min / args.steps},
            "wall_ms_per_step": wall / args.steps * 1e3,
            "clocks": sampler.result(),
            "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(W * H * 16), "steps": e2e_steps,
                    "wall_ms_per_step": float(te[0]) / e2e_steps * 1e3,
                    "rank0_breakdown_ms_per_step": {k: v / e2e_steps for k, v in e2e_parts.items()},
                    "note": "wall clock around commit + renderer creation + render + reduce + read-back of a FRESH scene and "
                            "renderer per step; the render inside it is a single render after ~1 s of host-side scene "
                            "building, the device-timed `value` is the mean of back-to-back renders (sustained clocks)",
                    "includes": "scene upload (H2D), GPU alias-table + BVH build, render, NCCL reduce (N>1), accumulation "
                                "read-back (D2H) through the C ABI",
                    "excludes": f"host-side scene construction (the facade's scene function: OBJ / HDR parse and flat arrays, "
                                f"{host_build_s:.2f} s here), which the reference also does before its first launch"},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu_base, "other_configs": others,
            "bvh": {"build_ms": info.build_ms, "sah_lbvh": info.sah_cost_lbvh, "sah_refined": info.sah_cost_refined},
        }
        emit_json(out)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def _check_rc(lib, rc):
    if rc != 0:
        raise SystemExit(f"ptb200 error {rc}: {lib.pt_last_error().decode()}")


if __name__ == "__main__":
    main()
