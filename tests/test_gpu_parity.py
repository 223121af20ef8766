"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI
(include/ptb200.h via the ctypes binding), against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star):
  1. first-hit primitive ids / t / barycentrics: BIT-EXACT on fixed ray sets;
  2. per-pixel radiance at fixed seed for non-stochastic passes (Normal_Shader, depth-1):
     |diff| <= 2e-5 absolute (libm sinf/atan2f/... differ by <= 2 ulp between glibc and CUDA);
  3. stochastic passes at fixed seed: identical RNG streams, so images agree except where a
     1-ulp libm difference flips a branch: <= 0.4 % of pixels may differ by > 1e-3 (1.2 % on the
     material zoo), the image means agree to 1 %, and the GPU must not show more differing pixels than
     the oracle with ulp-nudged libm does (tests/test_gpu_round2.py);
  4. converged images: RMSE on the gamma-encoded [0,1] image <= 0.02 against the oracle
     rendered with DIFFERENT frame indices (independent samples).
"""
import os

import numpy as np
import pytest

import helpers as h

pt = h.pt
pytestmark = pytest.mark.gpu


def same_hits(a, b):
    return ((a["prim"] == b["prim"]) & (a["t"].view(np.uint32) == b["t"].view(np.uint32)) &
            (a["u"].view(np.uint32) == b["u"].view(np.uint32)) & (a["v"].view(np.uint32) == b["v"].view(np.uint32)))


def ray_set(orc, W, H, extra=4000, seed=3):
    rays = np.concatenate([orc.camera_rays(W, H, 0), h.random_rays(extra, seed)])
    b1 = orc.bounce_rays(rays, 7)
    b2 = orc.bounce_rays(b1, 8)
    return np.concatenate([rays, b1, b2])


SCENES = {
    "cornell": lambda lib: h.build_cornell(64, 64, lib),
    "zoo": lambda lib: h.build_material_zoo(64, 64, lib),
    "spheres488": lambda lib: h.build_spheres(22, W=64, H=64, lib=lib),
    "soup": lambda lib: h.build_random_prims(3, lib=lib),
    "soup_big": lambda lib: h.build_random_prims(8, n_sph=600, n_rect=500, n_box=500, n_tri=5000, lib=lib),
    "mesh8k": lambda lib: h.build_mesh_scene(64, 64, 64, lib),
    "single": lambda lib: h.build_random_prims(5, n_sph=1, n_rect=0, n_box=0, n_tri=0, lib=lib),
}


@pytest.mark.parametrize("name", list(SCENES))
def test_hit_ids_bit_exact(lib, name):
    sc = SCENES[name](lib)
    sc.commit()
    orc = h.OracleScene(sc, mode=1 if sc.desc().n_prims <= 1000 else 2)
    rays = ray_set(orc, 64, 64)
    ho, _ = orc.trace(rays)
    hg = sc.trace(rays)
    ok = same_hits(ho, hg)
    assert ok.all(), f"{(~ok).sum()} of {len(rays)} rays differ"
    assert (ho["prim"] >= 0).sum() > 0


def test_hit_ids_bit_exact_mesh_250k_incoherent(lib):
    """>= 1 M rays on a 250k-triangle mesh, mostly incoherent secondary rays."""
    sc = h.build_mesh_scene(354, 512, 512, lib)
    sc.commit()
    info = sc.bvh_info()
    assert info.n_prims == 2 * 354 * 354 and info.n_leaf_records == info.n_prims
    assert info.sah_cost_refined <= info.sah_cost_lbvh
    orc = h.OracleScene(sc, mode=2)
    cam = orc.camera_rays(512, 512, 0)
    b1 = orc.bounce_rays(cam, 1)
    b2 = orc.bounce_rays(b1, 2)
    b3 = orc.bounce_rays(b2, 3)
    rays = np.concatenate([cam, b1, b2, b3])
    assert len(rays) >= 1_000_000
    ho, _ = orc.trace(rays)
    hg, cnt = sc.trace(rays, count_fetches=True)
    ok = same_hits(ho, hg)
    assert ok.all(), f"{(~ok).sum()} of {len(rays)} rays differ"
    assert cnt[0] > 0 and cnt[1] > 0


def test_trace_edge_cases(lib):
    sc = h.build_cornell(16, 16, lib)
    sc.commit()
    assert len(sc.trace(np.zeros(0, dtype=pt.RAY_DTYPE))) == 0          # empty ray set
    rays = np.zeros(4, dtype=pt.RAY_DTYPE)
    rays["origin"] = [[278, 278, -800]] * 4
    rays["dir"] = [[0, 0, 1], [0, 0, 1], [0, 0, -1], [0, 1e-30, 1]]
    rays["tmin"] = [1e-6, 10.0, 1e-6, 1e-6]
    rays["tmax"] = [1e27, 1.0, 1e27, 1e27]
    hg = sc.trace(rays)
    ho, _ = h.OracleScene(sc, mode=1).trace(rays)
    assert same_hits(ho, hg).all()
    assert hg["prim"][0] == 6 and hg["prim"][1] == -1 and hg["prim"][2] == -1  # tall box (id 6), empty interval, miss


def test_normal_shader_radiance_tolerance(lib):
    """Non-stochastic pass (programs/materials/normal_shader.cu): radiance = N*0.5+0.5."""
    sc = h.build_random_prims(3, lib=lib, W=96, H=96)
    sc.commit()
    r = pt.Renderer(sc, 96, 96)
    r.render(0, 1)
    g = r.read_accum()
    a, st = h.OracleScene(sc).render(96, 96, 0, 1)
    assert np.array_equal(g[..., 3], a[..., 3])
    assert np.abs(g - a).max() <= 2e-5
    s = r.stats()
    assert s.closest_rays == st.closest_rays == 96 * 96 and s.shadow_rays == 0


@pytest.mark.parametrize("name", ["cornell", "zoo", "spheres488", "mesh8k"])
def test_fixed_seed_radiance(lib, name):
    sc = SCENES[name](lib)
    sc.commit()
    W = H = 64
    r = pt.Renderer(sc, W, H)
    r.render(0, 4)
    g = r.read_accum()
    a, st = h.OracleScene(sc).render(W, H, 0, 4)
    assert (g[..., 3] == 4).all()
    d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    frac = float((d > 1e-3).mean())
    # bars from what was observed on the B200 (round 2: <= 0.12 % on every scene but the material zoo, whose GGX /
    # Ashikhmin-Shirley lobes amplify last-bit differences: 0.8 % at 16 spp) with a factor of ~3 of headroom; that
    # these differences ARE last-bit libm flips is measured in tests/test_gpu_round2.py against the ulp-noise oracle
    bar = 0.012 if name == "zoo" else 0.004
    assert frac <= bar, f"{frac:.4f} of pixels differ"
    assert abs(g[..., :3].mean() - a[..., :3].mean()) <= 0.01 * a[..., :3].mean()
    s = r.stats()
    assert abs(int(s.closest_rays) - int(st.closest_rays)) <= 0.01 * st.closest_rays
    assert abs(int(s.shadow_rays) - int(st.shadow_rays)) <= 0.01 * st.shadow_rays + 8


def test_depth1_is_deterministic_and_exact(lib):
    """max_depth=1: only camera rays that hit the light contribute (raygen.cu:85-90,122)."""
    sc = h.build_cornell(64, 64, lib)
    sc.commit()
    r = pt.Renderer(sc, 64, 64)
    r.render(0, 2, max_depth=1)
    g = r.read_accum()
    a, _ = h.OracleScene(sc).render(64, 64, 0, 2, max_depth=1)
    assert np.abs(g - a).max() <= 2e-5


def test_converged_rmse(lib):
    sc = h.build_cornell(64, 64, lib)
    sc.commit()
    n = 256
    r = pt.Renderer(sc, 64, 64)
    r.render(0, n)
    g = r.read_accum()
    a, _ = h.OracleScene(sc).render(64, 64, 1000, n)  # independent samples
    gi = np.sqrt(np.clip(g[..., :3] / n, 0, 1))
    ai = np.sqrt(np.clip(a[..., :3] / n, 0, 1))
    assert h.rmse(gi, ai) <= 0.02
    assert (g[..., 3] == n).all()
    # same seeds: the 8-bit images agree almost everywhere
    b, _ = h.OracleScene(sc).render(64, 64, 0, n)
    p1, p2 = pt.tonemap_png8(g, n).astype(int), pt.tonemap_png8(b, n).astype(int)
    assert np.abs(p1 - p2).mean() <= 0.25


def test_material_zoo_converged_rmse(lib):
    sc = h.build_material_zoo(48, 48, lib)
    sc.commit()
    n = 192
    r = pt.Renderer(sc, 48, 48)
    r.render(0, n)
    g = r.read_accum()
    a, _ = h.OracleScene(sc).render(48, 48, 5000, n)
    gi = np.sqrt(np.clip(g[..., :3] / n, 0, 1))
    ai = np.sqrt(np.clip(a[..., :3] / n, 0, 1))
    assert h.rmse(gi, ai) <= 0.03


def test_accumulation_properties(lib):
    """Size-independent properties: .w == samples, frame-range additivity, determinism,
    pool-size independence."""
    sc = h.build_spheres(22, W=160, H=96, lib=lib)
    sc.commit()
    W, H = 160, 96
    r = pt.Renderer(sc, W, H)
    r.render(0, 8)
    full = r.read_accum()
    assert (full[..., 3] == 8).all()
    r.render(0, 4, clear=True)
    r.render(4, 4)
    two = r.read_accum()
    assert (two[..., 3] == 8).all()
    assert np.allclose(full, two, rtol=1e-5, atol=1e-5)   # fp32 summation order only
    # one frame: every pixel receives exactly one sample -> bit-reproducible, pool independent
    r.render(3, 1, clear=True)
    a = r.read_accum()
    r2 = pt.Renderer(sc, W, H)
    r2.set_pool(2048)
    r2.render(3, 1)
    b = r2.read_accum()
    assert np.array_equal(a, b)
    assert r2.stats().iterations > r.stats().iterations


def test_display_preview_and_checkpoint_resume(lib):
    """Progressive loop pieces (main.cpp:264-293): the RGBA8 display buffer (make_Color, raygen.cu:136-146)
    converted on the device equals the oracle's conversion of the same accumulation buffer bit for bit;
    a render checkpointed with pt_read_accum and resumed in a NEW renderer through pt_write_accum
    continues exactly where it stopped."""
    import ctypes as C
    sc = h.build_cornell(96, 64, lib)
    sc.commit()
    W, H = 96, 64
    r = pt.Renderer(sc, W, H)
    for done in (1, 5):                       # preview after 1 and after 5 samples
        r.render(0 if done == 1 else 1, 1 if done == 1 else 4)
        acc = r.read_accum()
        want = np.empty((H, W, 4), dtype=np.uint8)
        h.oracle().orc_display_rgba8(acc.ctypes.data, W, H, want.ctypes.data)
        got = r.read_display()
        assert np.array_equal(got, want) and (got[..., 3] == 255).all() and got[..., :3].max() > 0
    ckpt = r.read_accum()                     # 5 samples per pixel
    r.render(5, 3)
    straight = r.read_accum()
    r2 = pt.Renderer(sc, W, H)
    r2.write_accum(ckpt)
    r2.render(5, 3)
    resumed = r2.read_accum()
    assert (resumed[..., 3] == 8).all()
    assert np.allclose(resumed, straight, rtol=1e-6, atol=1e-6)   # per-pixel fp32 summation order only
    # a restored buffer is not cleared by a following frame-0 render either
    r3 = pt.Renderer(sc, W, H)
    r3.write_accum(ckpt)
    r3.render(0, 1)
    assert (r3.read_accum()[..., 3] == 6).all()


def test_russian_roulette_and_depth_variants(lib):
    sc = h.build_cornell(32, 32, lib)
    sc.commit()
    o = h.OracleScene(sc)
    r = pt.Renderer(sc, 32, 32)
    for depth, rr in [(3, False), (14, True), (50, False)]:
        r.render(0, 3, max_depth=depth, russian=rr, clear=True)
        g = r.read_accum()
        a, _ = o.render(32, 32, 0, 3, max_depth=depth, russian=rr)
        d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
        assert (d > 1e-3).mean() <= 0.02, (depth, rr)


def test_textures_and_env_map(lib):
    """checker (programs/textures/checkered_texture.cu), image (image_texture.cu), vector
    (vector.cu) textures and the HDR environment miss program (miss.cu:64-95)."""
    s = pt.Scene(lib)
    img = np.random.default_rng(5).uniform(0, 1, (32, 64, 4)).astype(np.float32)
    iid = s.add_image(img)
    env = s.add_image(h.synthetic_hdri())
    texs = [h.tex_const(0.2, 0.3, 0.1), h.tex_const(0.9, 0.9, 0.9), pt.Texture(pt.TEX_CHECKER, pt.F3(0, 0, 0), 0, 1),
            pt.Texture(pt.TEX_IMAGE, pt.F3(0, 0, 0), iid, 0), pt.Texture(pt.TEX_VECTOR, pt.F3(0, 0, 0), 0, 3)]
    s.set_textures(texs)
    s.set_texture_vector([0, 1, 3])
    s.set_materials([h.material(pt.MAT_LAMBERTIAN, 2), h.material(pt.MAT_LAMBERTIAN, 3),
                     h.material(pt.MAT_LAMBERTIAN, 4)])
    s.add_spheres([pt.Sphere(pt.F3(0, -100.5, -1), 100), pt.Sphere(pt.F3(0, 0, -1), 0.5)], [0, 1])
    tris, _ = h.heightfield_mesh(8, size=2.0, amp=0.1)
    tris = tris + np.array([1.5, 0.2, -1.0] * 3, dtype=np.float32)
    uv = np.tile(np.array([0, 0, 1, 0, 1, 1], dtype=np.float32), (len(tris), 1))
    s.add_mesh(tris, 2, uv=uv, tex_index=np.arange(len(tris)) % 4)   # index 3 is out of range -> black
    s.set_camera(h.make_camera([0, 1, 3], [0.5, 0, -1], [0, 1, 0], 50, 1.0, 0.0, 1.0))
    s.set_miss(pt.Miss(pt.MISS_HDR, pt.F3(0, 0, 0), pt.F3(0, 0, 0), env, 1, 0))
    s.commit()
    r = pt.Renderer(s, 64, 64)
    r.render(0, 4)
    g = r.read_accum()
    a, _ = h.OracleScene(s).render(64, 64, 0, 4)
    d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    assert (d > 1e-3).mean() <= 0.015
    assert g[..., :3].mean() > 0.05


def test_caller_owned_accum_buffer(lib):
    """pt_renderer_create(d_accum): a torch tensor serves as the accumulation / NCCL buffer."""
    import torch
    sc = h.build_cornell(32, 32, lib)
    sc.commit()
    t = torch.zeros(32, 32, 4, device="cuda", dtype=torch.float32)
    r = pt.Renderer(sc, 32, 32, d_accum=t.data_ptr())
    r.render(0, 2)
    assert r.accum_device_ptr() == t.data_ptr()
    assert np.array_equal(t.cpu().numpy(), r.read_accum())
    assert float(t[..., 3].min()) == 2.0


def test_instrumented_counters(lib):
    sc = h.build_mesh_scene(64, 64, 64, lib)
    sc.commit()
    r = pt.Renderer(sc, 64, 64)
    r.set_flags(pt.FLAG_COUNT_FETCHES | pt.FLAG_TIME_KERNELS)
    r.render(0, 2)
    s = r.stats()
    assert s.nodes_fetched >= s.closest_rays and s.leaves_fetched > 0
    assert s.ms_trace_closest > 0 and s.ms_total > 0
    base = r.read_accum()
    r.set_flags(0)
    r.render(0, 2, clear=True)
    assert np.allclose(base, r.read_accum(), rtol=1e-5, atol=1e-6)


def test_env_importance_sampling_parity(lib):
    """NEW feature (pt_miss.env_nee): alias-table HDRI sampling + MIS, against the oracle's
    implementation of the same estimator at fixed seed, and converged against the reference
    estimator."""
    W = H = 64
    sc = h.build_mesh_scene(64, W, H, lib, env_nee=True)
    sc.commit()
    r = pt.Renderer(sc, W, H)
    r.render(0, 4)
    g = r.read_accum()
    a, st = h.OracleScene(sc).render(W, H, 0, 4)
    d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    assert (d > 1e-3).mean() <= 0.015
    assert abs(int(r.stats().shadow_rays) - int(st.shadow_rays)) <= 0.01 * st.shadow_rays + 8
    n = 256
    r.render(0, n, clear=True)
    g = r.read_accum()[..., :3] / n
    ref = h.build_mesh_scene(64, W, H, lib, env_nee=False)
    ref.commit()
    r2 = pt.Renderer(ref, W, H)
    r2.render(0, n)
    b = r2.read_accum()[..., :3] / n
    assert abs(g.mean() - b.mean()) <= 0.02 * b.mean()


def test_noise_gradient_checker_textures(lib):
    """Perlin noise / gradient / checker textures (programs/textures/*.cu) on the GPU vs the oracle."""
    import test_hostsim
    sc = test_hostsim._noise_scene(64, 64)
    sc.commit()
    r = pt.Renderer(sc, 64, 64)
    r.render(0, 4)
    g = r.read_accum()
    a, _ = h.OracleScene(sc).render(64, 64, 0, 4)
    d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    assert (d > 1e-3).mean() <= 0.02
    assert abs(g[..., :3].mean() - a[..., :3].mean()) <= 0.01 * a[..., :3].mean()


def test_two_lights_nee_parity(lib):
    """numLights = 2: light picking (light_sample.cuh:48), NEE from a light surface
    (diffuse_light.cu:41 with isLight=true) and its shadow ray, against the oracle."""
    def extra(s, texs, mats):
        pass
    sc = h.build_cornell(48, 48, lib, extra)
    # second emitter on the floor, facing up, plus its sampler entry
    sc.add_rects([pt.Rect(100, 200, 100, 200, 1.0, pt.Y_AXIS, 0)], [3])
    sc.set_lights([pt.Light(0, pt.Y_AXIS, 213, 343, 227, 332, 554, pt.F3(7, 7, 7)),
                   pt.Light(0, pt.Y_AXIS, 100, 200, 100, 200, 1.0, pt.F3(3, 3, 3))])
    sc.commit()
    r = pt.Renderer(sc, 48, 48)
    r.render(0, 6)
    g = r.read_accum()
    a, st = h.OracleScene(sc).render(48, 48, 0, 6)
    d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    assert (d > 1e-3).mean() <= 0.02
    assert abs(g[..., :3].mean() - a[..., :3].mean()) <= 0.01 * a[..., :3].mean()
    assert abs(int(r.stats().shadow_rays) - int(st.shadow_rays)) <= 0.01 * st.shadow_rays + 8


@pytest.mark.parametrize("two", [False, True])
def test_sphere_light_nee_parity(lib, two):
    """Sphere_PDF lights (programs/pdfs/sphere_pdf.cu) through Direct_Light, against the oracle."""
    sc = h.build_sphere_light_scene(48, 48, lib, two=two)
    sc.commit()
    r = pt.Renderer(sc, 48, 48)
    r.render(0, 6)
    g = r.read_accum()
    a, st = h.OracleScene(sc).render(48, 48, 0, 6)
    d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    assert (d > 1e-3).mean() <= 0.02
    assert abs(g[..., :3].mean() - a[..., :3].mean()) <= 0.01 * a[..., :3].mean()
    # On the sphere light's own surface sphere_pdf.cu:67 takes sqrt(1 - r^2/d^2) with d ~ r: NaN
    # directions.  The oracle traces (and counts) those shadow rays like the reference would; the
    # GPU skips a shadow ray whose outcome cannot matter (dot(N, Wi) > 0 is false for NaN).
    gs, os_ = int(r.stats().shadow_rays), int(st.shadow_rays)
    assert gs <= os_ + 8 and os_ - gs <= 0.03 * os_ + 8


def test_moving_spheres_parity(lib):
    """Motion blur (moving_sphere.cu): the sphere centre at prd.time uses never-contracted arithmetic, so
    hit ids / t stay bit-exact (ray sets see the scene at the camera's time0) and fixed-seed images agree
    like the static scenes do; converged RMSE against independent oracle samples."""
    W, H = 72, 48
    sc = h.build_motion_scene(W, H, lib)
    sc.commit()
    orc = h.OracleScene(sc, mode=1)
    rays = ray_set(orc, W, H, extra=2000)
    ho, _ = orc.trace(rays)
    hg = sc.trace(rays)
    assert same_hits(ho, hg).all()
    r = pt.Renderer(sc, W, H)
    r.render(0, 4)
    g = r.read_accum()
    a, st = orc.render(W, H, 0, 4)
    d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    assert (g[..., 3] == 4).all() and (d > 1e-3).mean() <= 0.015
    assert abs(g[..., :3].mean() - a[..., :3].mean()) <= 0.01 * a[..., :3].mean()
    assert abs(int(r.stats().closest_rays) - int(st.closest_rays)) <= 0.01 * st.closest_rays
    n = 128
    r.render(0, n, clear=True)
    g = r.read_accum()
    a, _ = orc.render(W, H, 3000, n)
    assert h.rmse(np.sqrt(np.clip(g[..., :3] / n, 0, 1)), np.sqrt(np.clip(a[..., :3] / n, 0, 1))) <= 0.03


@pytest.mark.parametrize("fog", [False, True])
def test_volumes_parity(lib, fog):
    """Volumetric_Box / Volumetric_Sphere + Isotropic (volume_*.cu, isotropic.cu) with the pinned
    per-(ray, primitive) uniform.  The scattering distance goes through logf, which differs by an
    ulp or two between CUDA and glibc: t agrees to 1e-5 relative, ids on all but a handful of rays."""
    W = H = 48
    sc = h.build_smoke_scene(W, H, lib, fog=fog)
    sc.commit()
    orc = h.OracleScene(sc, mode=1)
    rays = ray_set(orc, W, H, extra=0)
    rays = np.concatenate([rays, orc.camera_rays(W, H, 5)])
    ho, _ = orc.trace(rays)
    hg = sc.trace(rays)
    same_id = ho["prim"] == hg["prim"]
    assert (~same_id).mean() <= 1e-3
    hit = same_id & (ho["prim"] >= 0)
    assert np.allclose(ho["t"][hit], hg["t"][hit], rtol=1e-5, atol=0)
    vol = np.isin(ho["prim"], np.arange(6, 10))
    assert vol.mean() > 0.05                      # the volumes are actually being hit
    r = pt.Renderer(sc, W, H)
    r.render(0, 6)
    g = r.read_accum()
    a, st = orc.render(W, H, 0, 6)
    d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    assert (g[..., 3] == 6).all() and (d > 1e-3).mean() <= 0.03
    assert abs(g[..., :3].mean() - a[..., :3].mean()) <= 0.015 * a[..., :3].mean()
    assert abs(int(r.stats().shadow_rays) - int(st.shadow_rays)) <= 0.01 * st.shadow_rays + 8
    # converged: smoke is noisy (two independent 192-spp oracle renders differ by RMSE 0.034 / 0.049), so the
    # GPU renders 2048 spp and the oracle 512 independent ones: expected RMSE ~ 0.015 / 0.021 from noise alone
    ng, no = 2048, 512
    r.render(0, ng, clear=True)
    g = r.read_accum()
    a, _ = orc.render(W, H, 7000, no)
    assert h.rmse(np.sqrt(np.clip(g[..., :3] / ng, 0, 1)), np.sqrt(np.clip(a[..., :3] / no, 0, 1))) <= 0.03


def test_final_next_week_through_the_facade(lib, tmp_path):
    """The reference's Final_Next_Week scene function (scenes.hpp:263-375: 400 boxes, 1000 instanced
    spheres, Perlin sphere, glass + two Volumetric_Spheres, rect light) built by the C++ facade,
    rendered on the GPU and compared with the oracle on the same flat scene."""
    import subprocess, sys, ctypes as C
    host = C.CDLL(pt.HOST_LIB_PATH)
    host.pth_build_scene.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    hnd = C.c_void_p()
    assert host.pth_build_scene(3, 64, 64, 0, C.byref(hnd)) == 0
    sc = pt.Scene.__new__(pt.Scene)
    sc.lib, sc.h, sc._keep = lib, hnd, []
    d = sc.desc()
    assert d.n_volume_spheres == 2 and d.n_boxes == 400 and d.n_spheres >= 1005
    sc.commit()
    orc = h.OracleScene(sc, mode=2)
    r = pt.Renderer(sc, 64, 64)
    r.render(0, 4)
    g = r.read_accum()
    a, _ = orc.render(64, 64, 0, 4)
    dd = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    assert (dd > 1e-3).mean() <= 0.03
    assert abs(g[..., :3].mean() - a[..., :3].mean()) <= 0.02 * a[..., :3].mean()


def test_textured_mtl_mesh_parity(lib):
    """OBJ + MTL + map_Kd PNG loaded through the facade's Mesh (no given material):
    Vector_Texture per-face albedo, image lookup with REPEAT wrap, vt / vn interpolation."""
    import facade_dump
    if not os.path.exists(pt.HOST_LIB_PATH):
        pytest.fail("libptb200_host.so missing: run __graft_entry__.build()")
    sc = facade_dump.build(pt.HOST_LIB_PATH, 8, 64, 64, 24)
    sc.commit()
    r = pt.Renderer(sc, 64, 64)
    r.render(0, 4)
    g = r.read_accum()
    a, st = h.OracleScene(sc).render(64, 64, 0, 4)
    d = np.abs(g[..., :3] - a[..., :3]).max(axis=-1)
    assert (d > 1e-3).mean() <= 0.02
    assert abs(g[..., :3].mean() - a[..., :3].mean()) <= 0.01 * a[..., :3].mean()
    rays = h.random_rays(20000, 11, extent=5.0)
    hg = sc.trace(rays)
    ho, _ = h.OracleScene(sc, mode=1).trace(rays)
    assert same_hits(ho, hg).all()


def test_cli_renders_png(lib, tmp_path):
    """Headless CLI (app/main.cpp: the reference's main.cpp flow without the GUI): scene function ->
    facade -> C ABI -> GPU -> Save_PNG; the image must match the oracle's tonemapped render."""
    import os
    import subprocess
    from PIL import Image
    exe = os.path.join(h.PKG_DIR, "csrc", "ptb200_render")
    if not os.path.exists(exe):
        subprocess.run(["make", "ptb200_render"], cwd=os.path.join(h.PKG_DIR, "csrc"), check=True)
    out = str(tmp_path / "cornell")
    res = subprocess.run([exe, "--scene", "2", "--width", "64", "--height", "64", "--samples", "32", "--file", out],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Render time" in res.stdout and "BVH Building Time" in res.stdout
    img = np.array(Image.open(out + ".png")).astype(int)
    assert img.shape == (64, 64, 3)
    sc = h.build_cornell(64, 64, lib)
    a, _ = h.OracleScene(sc).render(64, 64, 0, 32)
    ref = pt.tonemap_png8(a, 32).astype(int)
    assert np.abs(img - ref).mean() <= 0.5
    # one-launch-per-sample mode (the reference's loop, main.cpp:264-293) gives the same picture
    res = subprocess.run([exe, "--scene", "2", "--width", "64", "--height", "64", "--samples", "8", "--batch", "1",
                          "--file", out + "_b1"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    img1 = np.array(Image.open(out + "_b1.png")).astype(int)
    a8, _ = h.OracleScene(sc).render(64, 64, 0, 8)
    assert np.abs(img1 - pt.tonemap_png8(a8, 8).astype(int)).mean() <= 0.5


def test_cli_checkpoint_resume_preview_and_new_scenes(lib, tmp_path):
    """CLI progressive extras: --checkpoint after 16 of 32 samples, --resume to 32 gives the image of an
    uninterrupted 32-sample run (8-bit identical up to fp32 summation order); --preview writes the display
    buffer; scenes 3 / 9 / 10 (Final_Next_Week with volumes, Motion_Blur, Cornell_Smoke) render."""
    import os
    import subprocess
    from PIL import Image
    exe = os.path.join(h.PKG_DIR, "csrc", "ptb200_render")
    base = ["--scene", "2", "--width", "64", "--height", "64"]
    run = lambda *a: subprocess.run([exe, *a], capture_output=True, text=True)
    ck = str(tmp_path / "ck.bin")
    res = run(*base, "--samples", "16", "--file", str(tmp_path / "half"), "--checkpoint", ck,
              "--preview", str(tmp_path / "prev.png"))
    assert res.returncode == 0, res.stdout + res.stderr
    half = np.array(Image.open(tmp_path / "half.png")).astype(int)
    prev = np.array(Image.open(tmp_path / "prev.png")).astype(int)
    assert np.abs(half - prev).max() <= 1          # Save_PNG multiplies in float, make_Color in double
    res = run(*base, "--samples", "32", "--file", str(tmp_path / "resumed"), "--resume", ck)
    assert res.returncode == 0 and "Resumed at sample 16" in res.stdout, res.stdout + res.stderr
    res = run(*base, "--samples", "32", "--file", str(tmp_path / "straight"))
    assert res.returncode == 0
    a = np.array(Image.open(tmp_path / "resumed.png")).astype(int)
    b = np.array(Image.open(tmp_path / "straight.png")).astype(int)
    assert np.abs(a - b).max() <= 1 and (a != b).mean() < 0.01
    res = run(*base, "--samples", "32", "--file", str(tmp_path / "bad"), "--resume", str(tmp_path / "nope.bin"))
    assert res.returncode == 3 and "cannot resume" in res.stderr
    for scene in ("3", "9", "10"):
        res = run("--scene", scene, "--width", "96", "--height", "64", "--samples", "8", "--file", str(tmp_path / f"s{scene}"))
        assert res.returncode == 0, res.stdout + res.stderr
        img = np.array(Image.open(tmp_path / f"s{scene}.png"))
        assert img.shape == (64, 96, 3) and img.std() > 5


def test_full_size_c2_and_c4_properties(lib):
    """BASELINE.json configurations at their FULL sizes (C2: 488 spheres 1200x800 @ 256 spp; C4: 1 M
    triangles + HDRI 1920x1080 @ 512 spp), checked through size-independent properties: every pixel's
    weight equals the sample count, frame ranges add up, the image mean agrees with the oracle's mean over a
    bounded sample of the same workload, and hit ids stay bit-exact on 1 M triangles."""
    for name, build, W, H, spp, tol in (("c2", lambda: h.build_spheres(22, W=1200, H=800, lib=lib), 1200, 800, 256, 0.01),
                                        ("c4", lambda: h.build_mesh_scene(708, 1920, 1080, lib=lib), 1920, 1080, 512, 0.01)):
        sc = build()
        sc.commit()
        assert (name != "c4") or sc.desc().n_tris == 708 * 708 * 2
        r = pt.Renderer(sc, W, H)
        r.render(0, spp)
        full = r.read_accum()
        assert (full[..., 3] == spp).all()
        assert np.isfinite(full).all() and full[..., :3].min() >= 0.0
        st = r.stats()
        assert st.samples == W * H * spp and st.closest_rays >= st.samples
        # frame-range additivity: [0, spp/2) + [spp/2, spp) == [0, spp) up to fp32 summation order
        r.render(0, spp // 2, clear=True)
        r.render(spp // 2, spp - spp // 2)
        two = r.read_accum()
        assert (two[..., 3] == spp).all()
        assert np.allclose(two[..., :3], full[..., :3], rtol=2e-4, atol=2e-3)
        # oracle on a bounded sample of the same workload: 2 spp at quarter resolution rows (same camera)
        orc = h.OracleScene(sc, mode=2)
        a, _ = orc.render(W, H, 10000, 1)
        gm, am = full[..., :3].mean() / spp, a[..., :3].mean()
        assert abs(gm - am) <= tol * am + 0.01, (name, gm, am)
        if name == "c4":
            rays = orc.camera_rays(640, 360, 0)
            rays = np.concatenate([rays, orc.bounce_rays(rays, 7), h.random_rays(200000, 5, extent=5.0)])
            ho, _ = orc.trace(rays)
            hg = sc.trace(rays)
            ok = same_hits(ho, hg)
            assert ok.all(), f"{(~ok).sum()} of {len(rays)} rays differ"
            assert (ho["prim"] >= 0).mean() > 0.3


def test_soup_with_moving_and_volumetric_primitives_gpu(lib):
    """Every primitive kind at once (incl. transformed Moving_Spheres and volumes) on the GPU: ids agree with the
    oracle except where a 1-ulp logf difference moves a volume scattering point across another surface
    (<= 0.1 % of rays); non-volume hits are bit-exact."""
    sc = h.build_random_prims(22, n_sph=300, n_rect=200, n_box=200, n_tri=3000, n_moving=200, n_vsph=40, n_vbox=40, lib=lib)
    sc.commit()
    orc = h.OracleScene(sc, mode=2)
    rays = ray_set(orc, 64, 64)
    ho, _ = orc.trace(rays)
    hg = sc.trace(rays)
    same_id = ho["prim"] == hg["prim"]
    assert (~same_id).mean() <= 1e-3
    kinds = np.array([sc.desc().prims[i].type for i in range(sc.desc().n_prims)])
    vol = np.zeros(len(rays), bool)
    hit = ho["prim"] >= 0
    vol[hit] = kinds[ho["prim"][hit]] >= pt.PRIM_VOLUME_SPHERE
    exact = same_id & ~vol
    assert (ho["t"][exact].view(np.uint32) == hg["t"][exact].view(np.uint32)).all()
    both = same_id & vol
    assert both.sum() > 100 and np.allclose(ho["t"][both], hg["t"][both], rtol=1e-5, atol=0)
    assert (kinds[ho["prim"][hit]] == pt.PRIM_MOVING_SPHERE).sum() > 100
