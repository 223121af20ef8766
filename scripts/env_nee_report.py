#!/usr/bin/env python
"""HDRI importance sampling as a measured feature (GPU): variance at equal time of the NEW estimator
(pt_miss.env_nee = 1: alias-table env sampling + MIS, csrc/shading.h) against the reference's estimator
(BSDF sampling only, the miss program adds the environment; programs/miss.cu:64-95) on the C4-shaped scene.

For each environment map and estimator: K independent renders of n spp (disjoint frame ranges), per-pixel variance of
the K means, device time per render.  Efficiency = 1 / (mean variance x time); the ratio of the two efficiencies is the
variance reduction at equal time.  Also reported: alias-table build time (GPU, inside commit) and the agreement of the
two estimators' converged means.

usage (GPU box): python scripts/env_nee_report.py > gpurun_out/r02_env_nee.md"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import helpers as h

pt = h.pt
W, H, GRID = 960, 540, 708
K, N = 8, 32


def scene(img, env_nee, lib):
    s = pt.Scene(lib)
    s.set_textures([h.tex_const(0.7, 0.7, 0.7)])
    s.set_materials([h.material(pt.MAT_LAMBERTIAN, 0)])
    tris, _ = h.heightfield_mesh(GRID)
    s.add_mesh(tris, 0)
    s.set_camera(h.make_camera([0, 4.5, -9], [0, 0, 0], [0, 1, 0], 40.0, W / H, 0.0, 10.0))
    i = s.add_image(img)
    s.set_miss(pt.Miss(pt.MISS_HDR, pt.F3(0, 0, 0), pt.F3(0, 0, 0), i, 1, int(env_nee)))
    return s


def measure(img, env_nee, lib):
    s = scene(img, env_nee, lib)
    t0 = time.perf_counter()
    s.commit()
    commit_s = time.perf_counter() - t0
    r = pt.Renderer(s, W, H)
    r.render(0, N, clear=True)                      # warm-up (graph capture)
    means, ms = [], 0.0
    for k in range(K):
        r.render(1000 + k * N, N, clear=True)
        ms += r.stats().ms_total
        means.append(r.read_accum()[..., :3] / N)
    means = np.stack(means)
    var = means.var(axis=0, ddof=1).mean()           # variance of an n-spp pixel estimate, averaged over the image
    big = means.mean(axis=0)                         # K * N spp
    st = r.stats()
    out = dict(var=float(var), ms=ms / K, mean=float(big.mean()), image=big, commit_s=commit_s,
               rays_per_sample=(st.closest_rays + st.shadow_rays) / st.samples)
    r.close()
    s.close()
    return out


def main():
    lib = pt.load()
    base = h.synthetic_hdri(2048, 1024)
    sun = base.copy()
    # the same sky with a small, very bright sun: what importance sampling is for.  (The reference estimator clamps
    # throughput x background to 1 at a miss, raygen.cu:82 - with this map it converges to a darker image; the new
    # estimator drops that clamp)
    v = (np.arange(1024) + 0.5) / 1024
    u = (np.arange(2048) + 0.5) / 2048
    U, V = np.meshgrid(u, v)
    disc = np.exp(-(((U - 0.3) / 0.006) ** 2 + ((V - 0.25) / 0.008) ** 2))
    sun[..., :3] += (40.0 * disc)[..., None]
    print("# Round 2: HDRI importance sampling (alias tables + MIS) against the reference estimator\n")
    print(f"Scene: the C4 mesh (1 M triangles, Lambertian 0.7) at {W}x{H}; {K} independent renders of {N} spp per "
          f"estimator and map; variance = per-pixel variance of the {N}-spp estimate, averaged over the image; one B200.\n")
    print("| environment map | estimator | ms per render | rays / sample | mean variance | efficiency 1/(var x ms) | converged mean |")
    print("|---|---|---|---|---|---|---|")
    res = {}
    for name, img in (("benchmark map (sky + soft sun, values <= 0.5)", base), ("same sky + small sun of 40", sun)):
        for env in (0, 1):
            m = measure(img, env, lib)
            res[(name, env)] = m
            print(f"| {name} | {'env NEE + MIS (new)' if env else 'reference (BSDF sampling only)'} | {m['ms']:.2f} | "
                  f"{m['rays_per_sample']:.2f} | {m['var']:.3e} | {1.0 / (m['var'] * m['ms']):.3e} | {m['mean']:.4f} |")
    print()
    for name in ("benchmark map (sky + soft sun, values <= 0.5)", "same sky + small sun of 40"):
        a, b = res[(name, 0)], res[(name, 1)]
        gain = (1.0 / (b["var"] * b["ms"])) / (1.0 / (a["var"] * a["ms"]))
        print(f"* {name}: variance ratio at equal SAMPLES {a['var'] / b['var']:.2f}x, cost per sample {b['ms'] / a['ms']:.2f}x, "
              f"**variance reduction at equal TIME {gain:.2f}x**; converged means differ by "
              f"{abs(a['mean'] - b['mean']) / a['mean'] * 100:.2f} %")
    # the reference estimator clamps throughput x background at a miss (raygen.cu:82): with the bright sun it is BIASED.
    # bias^2 per pixel from the two 256-spp means (their own noise, var / K each, subtracted); from it the render
    # time beyond which the unbiased estimator has the lower mean squared error
    name = "same sky + small sun of 40"
    a, b = res[(name, 0)], res[(name, 1)]
    bias2 = float(((a["image"] - b["image"]) ** 2).mean() - (a["var"] + b["var"]) / K)
    if bias2 > 0:
        # MSE(T) = var * (ms / T) + bias^2, T in ms of render time, var and ms per N-spp render
        t_star = (b["var"] * b["ms"] - a["var"] * a["ms"]) / bias2
        print(f"* {name}: the reference estimator's clamp costs it a mean squared bias of {bias2:.3e} per pixel "
              f"(image mean {a['mean']:.4f} against {b['mean']:.4f}); that is the variance it reaches after {a['var'] * a['ms'] / bias2:.0f} ms "
              f"of rendering, and beyond **{t_star:.0f} ms** ({t_star / b['ms'] * N:.0f} spp of the new estimator at this "
              f"resolution) the new estimator has the lower mean squared error - the reference estimator cannot "
              f"get below its bias however long it renders.")
    s = scene(base, 1, lib)
    s.commit()
    t0 = time.perf_counter()
    for _ in range(5):
        s.commit()
    t_env = (time.perf_counter() - t0) / 5
    s2 = scene(base, 0, lib)
    s2.commit()
    t0 = time.perf_counter()
    for _ in range(5):
        s2.commit()
    t_plain = (time.perf_counter() - t0) / 5
    print(f"\nAlias-table construction on the GPU (2048x1024 map, 2.1 M texels, inside pt_scene_commit): "
          f"{(t_env - t_plain) * 1e3:.2f} ms (commit {t_env * 1e3:.1f} ms with, {t_plain * 1e3:.1f} ms without the tables).")


if __name__ == "__main__":
    main()
