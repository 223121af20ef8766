"""CPU logic tests of the kernel cores (builder, BVH8 traversal, shading) by host
simulation (tests/hostsim/sim.cpp) against the oracle.  The GPU parity tests proper
are in test_gpu_parity.py; these catch logic errors without a GPU."""
import numpy as np
import pytest

import helpers as h

pt = h.pt


def same_hits(a, b):
    return ((a["prim"] == b["prim"]) & (a["t"].view(np.uint32) == b["t"].view(np.uint32)) &
            (a["u"].view(np.uint32) == b["u"].view(np.uint32)) & (a["v"].view(np.uint32) == b["v"].view(np.uint32)))


SCENES = {
    "cornell": lambda: h.build_cornell(32, 32),
    "zoo": lambda: h.build_material_zoo(32, 32),
    "spheres": lambda: h.build_spheres(12, W=32, H=32),
    "soup": lambda: h.build_random_prims(3),
    "soup_noxf": lambda: h.build_random_prims(4, xforms=False),
    "mesh": lambda: h.build_mesh_scene(40, 32, 32),
    "single": lambda: h.build_random_prims(5, n_sph=1, n_rect=0, n_box=0, n_tri=0),
    "two": lambda: h.build_random_prims(6, n_sph=0, n_rect=0, n_box=0, n_tri=2),
}


@pytest.mark.parametrize("name", list(SCENES))
def test_bvh8_structure_and_hit_ids(name):
    sc = SCENES[name]()
    sim = h.SimScene(sc)
    assert sim.validate() == 0
    orc = h.OracleScene(sc, mode=1)  # brute force: the definition of a hit
    rays = np.concatenate([orc.camera_rays(32, 32, 0), h.random_rays(3000, 11)])
    rays = np.concatenate([rays, orc.bounce_rays(rays, 5)])
    ho, _ = orc.trace(rays)
    hs, _ = sim.trace(rays)
    assert same_hits(ho, hs).all()


def test_oracle_bvh_equals_bruteforce():
    sc = h.build_mesh_scene(48, 32, 32)
    a, b = h.OracleScene(sc, mode=1), h.OracleScene(sc, mode=2)
    rays = np.concatenate([a.camera_rays(64, 64, 1), h.random_rays(5000, 2)])
    rays = np.concatenate([rays, a.bounce_rays(rays, 9)])
    ha, _ = a.trace(rays)
    hb, _ = b.trace(rays)
    assert same_hits(ha, hb).all()


def test_refinement_lowers_sah():
    sc = h.build_mesh_scene(64, 32, 32)
    info = h.SimScene(sc).info()
    assert info["n_treelets"] > 0 and info["sah_after"] < info["sah_before"]


def test_topdown_partition_and_optimal_collapse_cut_node_visits():
    """Round-2 builder (bvh_core.h 2b and 6a): the top-down binned-SAH partition of the upper levels and the
    SAH-optimal collapse, each switched off in turn, on a 32 k-triangle mesh: same hits, lower SAH cost, fewer BVH8
    nodes fetched per ray and (collapse) far fewer nodes in memory."""
    import os
    sc = h.build_mesh_scene(128, 64, 64)
    orc = h.OracleScene(sc, mode=1)
    cam = orc.camera_rays(64, 64, 0)
    rays = np.concatenate([cam, orc.bounce_rays(cam, 7)])
    ho, _ = orc.trace(rays)

    def run(**env):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update({k: "1" for k in env})
        try:
            sim = h.SimScene(sc)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        assert sim.validate() == 0
        hs, cnt = sim.trace(rays)
        assert same_hits(ho, hs).all()
        return sim.info(), cnt[0] / len(rays), cnt[1] / len(rays)

    full, n_full, l_full = run()
    no_td, n_no_td, _ = run(PTSIM_NO_TOPDOWN=True)
    no_dp, n_no_dp, _ = run(PTSIM_NO_DP=True)
    plain, n_plain, l_plain = run(PTSIM_NO_TOPDOWN=True, PTSIM_NO_DP=True)
    assert full["sah_after"] < 0.93 * no_td["sah_after"]          # the partition lowers the SAH cost of the BVH2
    assert n_full < n_no_td and n_full < n_no_dp < n_plain         # each step cuts node visits; together the most
    assert n_full < 0.93 * n_plain
    assert full["n_nodes8"] < 0.8 * no_dp["n_nodes8"]              # the optimal collapse needs far fewer wide nodes
    assert l_full < 1.1 * l_plain                                  # without paying in leaf tests


def test_degenerate_inputs():
    # coincident primitives (equal Morton codes, equal t): ties resolve to the lowest id
    s = pt.Scene()
    s.set_textures([h.tex_const(0.5)])
    s.set_materials([h.material(pt.MAT_NORMAL)])
    s.add_spheres([pt.Sphere(pt.F3(0, 0, 0), 1.0)] * 9, [0] * 9)
    sim, orc = h.SimScene(s), h.OracleScene(s, mode=1)
    assert sim.validate() == 0
    rays = h.random_rays(500, 4, extent=3.0)
    hs, _ = sim.trace(rays)
    ho, _ = orc.trace(rays)
    assert same_hits(ho, hs).all()
    assert set(np.unique(hs["prim"])) <= {-1, 0}
    # empty interval / zero direction components
    rays = np.zeros(3, dtype=pt.RAY_DTYPE)
    rays["origin"] = [[0, 0, -5], [0, 0, -5], [0.5, 0, -5]]
    rays["dir"] = [[0, 0, 1], [0, 0, 1], [0, 0, 1]]
    rays["tmin"] = [1e-3, 5.0, 1e-3]
    rays["tmax"] = [1e27, 1.0, 1e27]
    hs, _ = sim.trace(rays)
    assert hs["prim"][0] == 0 and abs(hs["t"][0] - 4.0) < 1e-6 and hs["prim"][1] == -1 and hs["prim"][2] == 0


@pytest.mark.parametrize("name,frames", [("cornell", 3), ("zoo", 3), ("spheres", 3), ("mesh", 2), ("soup", 1)])
def test_wavefront_logic_equals_reference_loop(name, frames):
    """The speculative two-continuation NEE scheme (shading.h) must reproduce the reference's
    per-sample RNG stream: on one CPU (same libm) the images are bit-identical."""
    sc = SCENES[name]()
    W = H = 24
    a, _ = h.OracleScene(sc).render(W, H, 0, frames)
    b = h.SimScene(sc).render(W, H, 0, frames)
    assert (a[..., 3] == frames).all() and (b[..., 3] == frames).all()
    assert np.array_equal(a, b)


def test_russian_roulette_and_depth_cutoff():
    sc = h.build_cornell(16, 16)
    o, s = h.OracleScene(sc), h.SimScene(sc)
    for depth, rr in [(1, True), (3, False), (12, True), (50, False)]:
        a, _ = o.render(16, 16, 0, 2, max_depth=depth, russian=rr)
        b = s.render(16, 16, 0, 2, max_depth=depth, russian=rr)
        assert np.array_equal(a, b), (depth, rr)
    # programs/raygen.cu:122: depth exhaustion returns black, so depth=1 only sees the light
    a, _ = o.render(16, 16, 0, 1, max_depth=1)
    assert a[..., :3].max() <= 7.0 + 1e-6


def test_triangle_boundary_rays():
    """Rays aimed exactly at mesh vertices and edge midpoints (u=0, v=0, u+v=1 boundaries) and
    grazing directions: the division-free pre-rejection in isect_triangle must not change any
    accept/reject decision or value."""
    sc = h.build_mesh_scene(12, 16, 16)
    d = sc.desc()
    v = np.ctypeslib.as_array(d.tri_v, shape=(d.n_tris, 3, 3)).copy()
    targets = np.concatenate([v.reshape(-1, 3), (v[:, 0] + v[:, 1]) / 2, (v[:, 1] + v[:, 2]) / 2,
                              (v[:, 0] + v[:, 2]) / 2]).astype(np.float32)
    rays = np.zeros(len(targets) * 2, dtype=pt.RAY_DTYPE)
    for k, org in enumerate(([0.25, 6.0, -0.5], [3.0, 0.3, -7.0])):
        sl = slice(k * len(targets), (k + 1) * len(targets))
        rays["origin"][sl] = org
        rays["dir"][sl] = targets - np.array(org, dtype=np.float32)
    rays["tmin"] = 1e-3
    rays["tmax"] = 1e27
    ho, _ = h.OracleScene(sc, mode=1).trace(rays)
    hs, _ = h.SimScene(sc).trace(rays)
    assert same_hits(ho, hs).all()
    assert (ho["prim"] >= 0).mean() > 0.5


def test_env_alias_table_is_the_texel_distribution():
    """pt_miss.env_nee: the alias table must reproduce the texel pdf exactly
    (P(i) = (prob_i + sum_{j: alias_j = i} (1 - prob_j)) / n) and the pdf must follow
    luminance x sin(theta) of the environment image."""
    sc = h.build_mesh_scene(8, env_nee=True)
    d = sc.desc()
    n = d.env_n
    assert n == 256 * 128
    prob = np.ctypeslib.as_array(d.env_alias_prob, shape=(n,)).astype(np.float64)
    alias = np.ctypeslib.as_array(d.env_alias_idx, shape=(n,))
    pdf = np.ctypeslib.as_array(d.env_texel_pdf, shape=(n,)).astype(np.float64)
    assert abs(pdf.sum() - 1.0) < 1e-5 and (pdf > 0).all()
    rebuilt = prob.copy()
    np.add.at(rebuilt, alias, 1.0 - prob)
    assert np.allclose(rebuilt / n, pdf, rtol=2e-4, atol=1e-9)
    img = h.synthetic_hdri()
    lum = 2.0 * (0.2126 * img[..., 0] + 0.7152 * img[..., 1] + 0.0722 * img[..., 2])
    w = lum * np.sin(np.pi * (np.arange(128) + 0.5) / 128)[:, None]
    w = w + 0.01 * w.sum() / n * np.sin(np.pi * (np.arange(128) + 0.5) / 128)[:, None]
    assert np.allclose(pdf.reshape(128, 256), w / w.sum(), rtol=1e-4)
    # reference mode carries no tables
    assert h.build_mesh_scene(8, env_nee=False).desc().env_n == 0


def test_env_nee_logic_equals_oracle():
    sc = h.build_mesh_scene(24, 24, 24, env_nee=True)
    a, st = h.OracleScene(sc).render(24, 24, 0, 3)
    b = h.SimScene(sc).render(24, 24, 0, 3)
    assert np.array_equal(a, b)
    assert st.shadow_rays > 0


def test_env_nee_converges_to_reference_estimator():
    """Same integrand (the synthetic HDRI stays <= 0.5 so the reference's clamp at miss is
    inactive): the importance-sampled estimator and the reference estimator agree on average,
    and the new one is less noisy."""
    W = H = 20
    ref_sc = h.build_mesh_scene(16, W, H, env_nee=False)
    nee_sc = h.build_mesh_scene(16, W, H, env_nee=True)
    n = 384
    a, _ = h.OracleScene(ref_sc).render(W, H, 0, n)
    b, _ = h.OracleScene(nee_sc).render(W, H, 0, n)
    a, b = a[..., :3] / n, b[..., :3] / n
    assert abs(a.mean() - b.mean()) <= 0.02 * a.mean()
    assert h.rmse(a, b) <= 0.06 * a.mean() + 0.01


def _noise_scene(W=32, H=32):
    return h.build_noise_scene(W, H)


def test_noise_and_gradient_textures_equal_oracle():
    """Perlin turbulence (programs/textures/noise_texture.cu) and gradient textures: the device
    cores reproduce the oracle bit-identically on the CPU."""
    sc = _noise_scene()
    a, _ = h.OracleScene(sc).render(32, 32, 0, 3)
    b = h.SimScene(sc).render(32, 32, 0, 3)
    assert np.array_equal(a, b)
    # the noise texture really modulates the albedo: the left sphere is not uniformly shaded
    left = a[8:24, 4:12, 0]
    assert left.std() > 0.02


@pytest.mark.parametrize("two", [False, True])
def test_sphere_light_logic_equals_oracle(two):
    """Sphere_PDF lights (programs/pdfs/sphere_pdf.cu): the device cores reproduce
    the oracle's Direct_Light with a sphere sampler bit-identically on the CPU, alone and mixed
    with a rect light (light picking, light_sample.cuh:48)."""
    sc = h.build_sphere_light_scene(24, 24, two=two)
    a, st = h.OracleScene(sc).render(24, 24, 0, 3)
    b = h.SimScene(sc).render(24, 24, 0, 3)
    assert np.array_equal(a, b)
    assert st.shadow_rays > 0 and np.isfinite(a).all()


# ---- motion blur (Moving_Sphere) and volumes (Volumetric_* + Isotropic) ----------
def test_moving_spheres_and_volumes_equal_oracle():
    """Kernel cores (host-compiled) against the oracle: hit ids/t bit-exact on ray sets and the
    rendered images bit-identical (same libm on one CPU, logf in the volume distance included)."""
    for build, W, H in ((h.build_motion_scene, 36, 24), (h.build_smoke_scene, 24, 24)):
        sc = build(W, H)
        sim = h.SimScene(sc)
        assert sim.validate() == 0
        orc = h.OracleScene(sc, mode=1)
        rays = np.concatenate([orc.camera_rays(W, H, 0), orc.camera_rays(W, H, 3)])
        rays = np.concatenate([rays, orc.bounce_rays(rays, 5)])
        ho, _ = orc.trace(rays)
        hs, _ = sim.trace(rays)
        assert same_hits(ho, hs).all()
        assert (ho["prim"] >= 0).mean() > 0.5
        a, _ = orc.render(W, H, 0, 3)
        b = sim.render(W, H, 0, 3)
        assert (a[..., 3] == 3).all() and np.array_equal(a, b)
        # the oracle's BVH agrees with its brute-force loop for the new primitive kinds
        hb, _ = h.OracleScene(sc, mode=2).trace(rays)
        assert same_hits(ho, hb).all()


def test_motion_blur_uses_the_sample_time():
    """A sphere moving along x during the shutter: every sample sees it at centre(prd.time) with
    prd.time = time0 + rnd * (time1 - time0) (raygen.cu:69), so the blurred footprint is wider than the
    static one and narrower camera shutters narrow it again."""
    def footprint(t1, c1x):
        s = pt.Scene()
        s.set_textures([h.tex_const(0.5)])
        s.set_materials([h.material(pt.MAT_NORMAL)])
        s.add_moving_spheres([pt.MovingSphere(pt.F3(-1, 0, 0), pt.F3(c1x, 0, 0), 0.5, 0.0, 1.0)], [0])
        s.set_camera(h.make_camera([0, 0, -10], [0, 0, 0], [0, 1, 0], 30.0, 2.0, 0.0, 10.0, 0.0, t1))
        s.set_miss(pt.Miss(pt.MISS_CONSTANT, pt.F3(0, 0, 0), pt.F3(0, 0, 0), -1, 1, 0))
        a, _ = h.OracleScene(s, mode=1).render(64, 32, 0, 16, max_depth=2)
        b = h.SimScene(s).render(64, 32, 0, 16, max_depth=2)
        assert np.array_equal(a, b)
        return int((a[..., :3].sum(axis=-1) > 0).any(axis=0).sum())   # covered image columns
    static, blurred, half = footprint(1.0, -1.0), footprint(1.0, 1.0), footprint(0.5, 1.0)
    assert blurred > half > static


def test_volume_free_flight_is_exponential():
    """Beer-Lambert through the pinned per-(ray, primitive) uniform: rays entering a slab of density
    sigma reach a wall D behind the entry face with probability exp(-sigma * D); scattering
    distances are exponential with mean 1/sigma (volume_box.cu:52-56)."""
    sigma, D = 0.02, 60.0
    s = pt.Scene()
    s.set_textures([h.tex_const(1.0)])
    s.set_materials([h.material(pt.MAT_ISOTROPIC, 0), h.material(pt.MAT_NORMAL)])
    s.add_volume_boxes([pt.VolumeBox(pt.F3(-50, -50, 0), pt.F3(50, 50, 500), sigma)], [0])
    s.add_rects([pt.Rect(-50, 50, -50, 50, D, pt.Z_AXIS, 0)], [1])
    n = 200000
    rng = np.random.default_rng(5)
    rays = np.zeros(n, dtype=pt.RAY_DTYPE)
    rays["origin"][:, :2] = rng.uniform(-20, 20, (n, 2))
    rays["origin"][:, 2] = -10.0
    rays["dir"][:, 2] = 2.0        # unnormalised on purpose: distances are scaled by |d|
    rays["tmin"], rays["tmax"] = 1e-3, 1e27
    hs, _ = h.SimScene(s).trace(rays)
    ho, _ = h.OracleScene(s, mode=1).trace(rays)
    assert same_hits(ho, hs).all()
    wall = hs["prim"] == 1
    assert abs(wall.mean() - np.exp(-sigma * D)) < 4e-3
    dist = (hs["t"][~wall] * 2.0) - 10.0           # metric distance behind the entry face
    assert (dist > 0).all() and (dist < D).all()
    # truncated-exponential mean of the scattering distances
    expect = 1 / sigma - D * np.exp(-sigma * D) / (1 - np.exp(-sigma * D))
    assert abs(dist.mean() - expect) < 0.3


def test_quantisation_guard_for_huge_next_to_small():
    """BVH8 child boxes live on the parent's 8-bit grid.  The RTOW scenes put a 1000-unit ground sphere next
    to 0.2-unit spheres: opened into the root, the small-sphere subtrees would all be rounded out to the same
    ~8-unit cells and every ray would visit all of them (7.5 nodes per ray before the guard in collapse_node).
    With the guard the small spheres stay under one inner child with a grid of its own."""
    sc = h.build_spheres(22)
    sim = h.SimScene(sc)
    assert sim.validate() == 0
    orc = h.OracleScene(sc, mode=1)
    rays = orc.camera_rays(64, 64, 0)
    rays = np.concatenate([rays, orc.bounce_rays(rays, 7)])
    hs, cnt = sim.trace(rays)
    ho, _ = orc.trace(rays)
    assert same_hits(ho, hs).all()
    assert cnt[0] / len(rays) < 5.5, cnt[0] / len(rays)


def test_soup_with_moving_and_volumetric_primitives():
    """Random soup of every primitive kind incl. transformed Moving_Spheres / volumes: BVH structure valid, hit ids
    and t bit-identical to the oracle's brute-force loop on the CPU, Normal_Shader images identical."""
    sc = h.build_random_prims(21, n_sph=30, n_rect=20, n_box=20, n_tri=150, n_moving=25, n_vsph=12, n_vbox=12)
    sim, orc = h.SimScene(sc), h.OracleScene(sc, mode=1)
    assert sim.validate() == 0
    rays = np.concatenate([orc.camera_rays(48, 48, 0), h.random_rays(4000, 11)])
    rays = np.concatenate([rays, orc.bounce_rays(rays, 5)])
    ho, _ = orc.trace(rays)
    hs, _ = sim.trace(rays)
    assert same_hits(ho, hs).all()
    kinds = np.array([sc.desc().prims[i].type for i in range(sc.desc().n_prims)])
    hit_kinds = set(kinds[ho["prim"][ho["prim"] >= 0]])
    assert {pt.PRIM_MOVING_SPHERE, pt.PRIM_VOLUME_SPHERE, pt.PRIM_VOLUME_BOX} <= hit_kinds
    a, _ = orc.render(32, 32, 0, 3)
    assert np.array_equal(a, sim.render(32, 32, 0, 3))


def test_octant_table_of_the_node_step():
    """pt_perm8 (traverse.h): bit s of the child-slot mask goes to priority bit s ^ octinv; an involution, popcount
    preserving, identity for octinv 0."""
    s = h.hostsim()
    for o in range(8):
        for m in range(256):
            p = s.sim_perm8(m, o)
            assert p == sum(1 << (b ^ o) for b in range(8) if m >> b & 1)
            assert s.sim_perm8(p, o) == m
        assert s.sim_perm8(0xA5, 0) == 0xA5


@pytest.mark.parametrize("builder", ["spheres", "mesh", "soup", "cornell"])
def test_node_encoding_invariants_the_hit_mask_assembly_relies_on(builder):
    """Empty child slots carry meta 0, no imask bit and the inverted box (they can never pass the slab test and would
    add nothing if they did); inner slot s is encoded as (1 << 5) | (24 + s); leaf record ranges fit the 24-bit
    field.  Checked on every node of four kinds of scene."""
    sc = {"spheres": lambda: h.build_spheres(22, 0, 32, 32),
          "mesh": lambda: h.build_mesh_scene(48, 32, 32),
          "soup": lambda: h.build_random_prims(seed=5, n_moving=6, n_vsph=3, n_vbox=3),
          "cornell": lambda: h.build_cornell(32, 32)}[builder]()
    sim = h.SimScene(sc)
    assert sim.s.sim_node_invariant_violations(sim.h) == 0


# ---- order of the camera samples (csrc/sample_order.h, the function k_generate calls) ---------------------------
def _sample_order(W, H, tiled, kmax, nframes, chunks_seed=0, max_chunk=5000):
    """ids 0 .. W*H*nframes in launches of random sizes, as the renderer's iterations walk them"""
    sim = h.hostsim()
    total = W * H * nframes
    out = np.empty((total, 3), dtype=np.int32)
    rng = np.random.default_rng(chunks_seed)
    s = 0
    while s < total:
        n = int(min(total - s, rng.integers(1, max_chunk)))
        sim.sim_sample_order(W, H, tiled, kmax, nframes, s, n, out[s:].ctypes.data)
        s += n
    return out


@pytest.mark.parametrize("W,H,tiled,kmax,nframes", [
    (16, 8, 1, 0, 3), (16, 8, 1, 2, 8), (16, 8, 1, 2, 7), (24, 12, 1, 5, 77), (10, 7, 0, 3, 19), (8, 4, 1, 8, 300),
    (8, 4, 1, 5, 1), (32, 16, 1, 6, 63), (32, 16, 1, 6, 65), (40, 30, 0, 8, 5), (64, 32, 1, 8, 256), (8, 8, 1, 3, 2)])
def test_sample_order_generates_every_sample_once(W, H, tiled, kmax, nframes):
    """k_generate may emit the (pixel, frame) pairs of a render in any order (seed = tea<64>(W*y + x, frame),
    programs/raygen.cu:150; accumulation is a sum, :166) but must emit each exactly once - full groups of 2^kmax
    frames, the remainder in groups of the smaller powers of two, tiled and untiled, whatever the launch boundaries -
    and a pixel's frames must come in ascending order (the serial sum of a pixel then adds in the reference's order)."""
    o = _sample_order(W, H, tiled, kmax, nframes)
    x, y, f = o[:, 0].astype(np.int64), o[:, 1].astype(np.int64), o[:, 2].astype(np.int64)
    assert x.min() == 0 and x.max() == W - 1 and y.min() == 0 and y.max() == H - 1
    assert f.min() == 0 and f.max() == nframes - 1
    key = (f * H + y) * W + x
    assert len(np.unique(key)) == W * H * nframes == len(key)
    pix = y * W + x
    order = np.argsort(pix, kind="stable")          # ids of one pixel, in emission order
    fs = f[order].reshape(W * H, nframes)
    assert (np.diff(fs, axis=1) > 0).all()
    # the same ids whatever the launch boundaries
    assert np.array_equal(o, _sample_order(W, H, tiled, kmax, nframes, chunks_seed=7, max_chunk=97))


def test_sample_order_keeps_a_warp_on_one_pixel_and_a_tile_compact():
    """What the order is for: with F >= 32 frames per group, 32 consecutive ids inside a full group are samples of ONE
    pixel; with F = 1, 2^m consecutive positions of a tile are a compact block (2x1, 2x2, 4x2, 4x4, 8x4); with F = 4,
    32 consecutive ids are 8 pixels (a 4x2 block) x 4 frames."""
    W, H = 64, 32
    o = _sample_order(W, H, 1, 5, 64).reshape(-1, 32, 3)
    assert (o[:, :, 0] == o[:, :1, 0]).all() and (o[:, :, 1] == o[:, :1, 1]).all()
    assert (np.diff(o[:, :, 2], axis=1) == 1).all()
    o = _sample_order(W, H, 1, 0, 1)
    for m, (bw, bh) in enumerate([(2, 1), (2, 2), (4, 2), (4, 4), (8, 4)], start=1):
        b = o.reshape(-1, 1 << m, 3)
        assert ((b[:, :, 0].max(1) - b[:, :, 0].min(1)) == bw - 1).all()
        assert ((b[:, :, 1].max(1) - b[:, :, 1].min(1)) == bh - 1).all()
    o = _sample_order(W, H, 1, 2, 8).reshape(-1, 32, 3)
    assert ((o[:, :, 0].max(1) - o[:, :, 0].min(1)) == 3).all() and ((o[:, :, 1].max(1) - o[:, :, 1].min(1)) == 1).all()
    assert (o[:, :, 2].max(1) - o[:, :, 2].min(1) == 3).all()
