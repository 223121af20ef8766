"""The reference ITSELF on the CPU versus the oracle, whole images.

oracle/ref_render.cpp wraps the reference's own ray-generation, miss, intersection, closest-hit, any-hit,
texture and light programs (compiled UNMODIFIED from /root/reference, host_includes/programs.hpp's wiring)
in an emulated OptiX 6 pipeline: a BVH over the reference's own bounds programs, closest-hit dispatch,
any-hit shadow rays.  Its images must be BIT-identical to liboracle_r2l.so (the oracle built with g++'s
argument evaluation order, which is what the reference's source gets from g++), float for float and byte for
byte of the display buffer.  This is the end-to-end link of the pinning chain in DESIGN.md section 3:
reference programs -> (these tests) -> oracle -> (tests/test_gpu_parity.py) -> CUDA path.

Live tests need /root/reference (they skip on the GPU box); the golden replay at the bottom does not.
"""
import json
import os

import numpy as np
import pytest

import helpers as h

pt = h.load_package()

needs_ref = pytest.mark.skipif(h.ReferenceRenderer.lib() is None, reason="oracle/_ref/libref_render.so not built")

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_render_vectors.json")


def build_empty(W, H, constant=False):
    """no primitives at all: every camera ray runs the miss program (miss.cu) and nothing else"""
    s = pt.Scene(None)
    s.set_textures([h.tex_const(0.5)])
    s.set_materials([h.material(pt.MAT_LAMBERTIAN, 0)])
    s.set_camera(h.make_camera([0, 0, -5], [0, 0, 0], [0, 1, 0], 40.0, W / H, 0.1, 5.0))
    if constant:
        s.set_miss(pt.Miss(pt.MISS_CONSTANT, pt.F3(0.25, 0.5, 0.75), pt.F3(0, 0, 0), -1, 1, 0))
    else:
        s.set_miss(pt.Miss(pt.MISS_GRADIENT, pt.F3(1, 1, 1), pt.F3(0.5, 0.7, 1.0), -1, 1, 0))
    return s


build_backdrop = h.build_backdrop


def scene_table():
    """name -> (builder, W, H, frames, max_depth)"""
    return {
        "cornell": (lambda: h.build_cornell(40, 40), 40, 40, 3, 50),
        "zoo": (lambda: h.build_material_zoo(40, 40), 40, 40, 3, 50),
        "spheres_bvh": (lambda: h.build_spheres(22, 0, 40, 40), 40, 40, 2, 50),
        "spheres_lights": (lambda: h.build_spheres(22, 0, 40, 40, lights=True), 40, 40, 2, 50),
        "sphere_light": (lambda: h.build_sphere_light_scene(32, 32), 32, 32, 3, 50),
        "two_lights": (lambda: h.build_sphere_light_scene(32, 32, two=True), 32, 32, 3, 50),
        "smoke": (lambda: h.build_smoke_scene(32, 32), 32, 32, 3, 50),
        "smoke_nofog": (lambda: h.build_smoke_scene(32, 32, fog=False), 32, 32, 2, 50),
        "motion": (lambda: h.build_motion_scene(48, 32), 48, 32, 3, 50),
        "mesh_env": (lambda: h.build_mesh_scene(12, 32, 32), 32, 32, 2, 50),
        "mesh_normals": (lambda: h.build_mesh_scene(12, 32, 32, normals=True), 32, 32, 2, 50),
        "cornell_depth3": (lambda: h.build_cornell(32, 32), 32, 32, 2, 3),
        "noise_gradient_checker": (lambda: h.build_noise_scene(32, 32), 32, 32, 2, 50),
        "image_background": (lambda: build_backdrop(24, 16, pt.MISS_IMG), 24, 16, 2, 50),
        "hdr_flat_background": (lambda: build_backdrop(24, 16, pt.MISS_HDR, spherical=0), 24, 16, 2, 50),
        "empty_gradient": (lambda: build_empty(24, 16), 24, 16, 2, 50),
        "empty_constant": (lambda: build_empty(24, 16, True), 24, 16, 2, 50),
    }


def same_bits(a, b):
    """float images equal bit for bit (NaNs compare equal to NaNs)"""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


@needs_ref
@pytest.mark.parametrize("name", list(scene_table().keys()))
def test_reference_renderer_equals_oracle(name):
    build, W, H, frames, depth = scene_table()[name]
    sc = build()
    ref = h.ReferenceRenderer(sc)
    acc_ref, disp_ref = ref.render(W, H, 0, frames, max_depth=depth)
    orc = h.OracleScene(sc, 0, "_r2l")
    acc_orc, _ = orc.render(W, H, 0, frames, max_depth=depth)
    assert np.isfinite(acc_orc[..., :3]).all()
    assert acc_orc[..., :3].max() > 0
    assert same_bits(acc_ref, acc_orc), (
        f"{name}: {np.count_nonzero(acc_ref != acc_orc)} of {acc_ref.size} floats differ, "
        f"max |d| = {np.nanmax(np.abs(acc_ref - acc_orc))}")
    # display buffer: raygen.cu's sqrt/clamp/255.99 on the running mean, against the oracle's restatement
    disp_orc = np.zeros((H, W, 4), np.uint8)
    orc.o.orc_display_rgba8(acc_orc.ctypes.data, W, H, disp_orc.ctypes.data)
    assert np.array_equal(disp_ref, disp_orc)


# scene functions of the C++ facade (host_includes/scenes.hpp: the reference's scene-building surface), by id:
# 0 InOneWeekend, 1 MovingSpheres (moving spheres + checker + earth image), 2 Cornell, 4 Test_Scene (1 M-triangle
# mesh + HDRI), 5 material zoo, 6 mesh + env map, 7 sphere light, 8 MTL-textured mesh (per-triangle texture index
# through Vector_Texture), 9 motion blur, 10 Cornell smoke
FACADE_SCENES = [(0, 0), (1, 0), (2, 0), (4, 0), (5, 0), (6, 8), (7, 0), (8, 8), (9, 0), (10, 0)]


def facade_scene(scene_id, model, W, H):
    import facade_dump
    h.build_product()
    if not os.path.exists(pt.HOST_LIB_PATH):
        pytest.skip("facade library not built")
    return facade_dump.build(pt.HOST_LIB_PATH, scene_id, W, H, model)


@needs_ref
@pytest.mark.parametrize("scene_id,model", FACADE_SCENES)
def test_facade_scene_functions_render_identically(scene_id, model, capfd):
    """scene function -> facade -> flat scene description -> {the reference's programs, the oracle}: same image,
    bit for bit.  Covers what the hand-built scenes above do not: checker / noise / image / vector textures, the
    facade's transforms, a million-triangle mesh under an HDR environment."""
    W, H = 36, 24
    sc = facade_scene(scene_id, model, W, H)
    a, _ = h.ReferenceRenderer(sc).render(W, H, 0, 2)
    b, _ = h.OracleScene(sc, 0, "_r2l").render(W, H, 0, 2)
    assert b[..., :3].max() > 0
    assert same_bits(a, b), f"scene {scene_id}: {np.count_nonzero(a != b)} of {a.size} floats differ"


@needs_ref
def test_final_next_week_differs_only_through_transformed_spheres(capfd):
    """Final_Next_Week holds 1000 spheres under translate + rotate: the one pinned difference (see below); everything
    else in it (noise and image textures, both volumes, the moving sphere, 400 boxes) must not add to it."""
    W, H = 36, 24
    sc = facade_scene(3, 0, W, H)
    a, _ = h.ReferenceRenderer(sc).render(W, H, 0, 2)
    b, _ = h.OracleScene(sc, 0, "_r2l").render(W, H, 0, 2)
    differ = a != b
    assert np.count_nonzero(differ) < 0.05 * a.size
    if differ.any():   # last-bit normal differences; a path may fork on one, hence a quantile rather than the maximum
        rel = np.abs(a - b)[differ] / np.maximum(np.abs(b)[differ], 1e-3)
        assert np.quantile(rel, 0.9) < 1e-3


@needs_ref
def test_reference_renderer_row_bands_and_frame_ranges():
    """Rows and frames are independent in the reference (seed = tea<64>(W*y+x, frame)): rendering by bands and by
    frame ranges into one accumulation buffer gives the whole image -- what bench.py's reference arm relies on
    when it spreads the reference over the host cores."""
    W = H = 32
    sc = h.build_cornell(W, H)
    ref = h.ReferenceRenderer(sc)
    whole, _ = ref.render(W, H, 0, 4)
    acc = np.zeros((H, W, 4), np.float32)
    for y0 in range(0, H, 8):
        ref.render(W, H, 0, 2, rows=(y0, y0 + 8), acc=acc)
    for y0 in range(0, H, 16):
        ref.render(W, H, 2, 2, rows=(y0, y0 + 16), acc=acc)
    assert same_bits(whole, acc)


@needs_ref
def test_reference_renderer_no_russian_roulette():
    W = H = 24
    sc = h.build_material_zoo(W, H)
    ref = h.ReferenceRenderer(sc)
    a, _ = ref.render(W, H, 0, 2, max_depth=8, russian=False)
    orc = h.OracleScene(sc, 0, "_r2l")
    b, _ = orc.render(W, H, 0, 2, max_depth=8, russian=False)
    assert same_bits(a, b)


@needs_ref
def test_transformed_sphere_normal_is_the_one_known_difference():
    """A Sphere under a non-identity transform: the reference's hit-record program normalises in a different
    order than the oracle's (and the CUDA path's) single formulation (pinned difference);
    the images agree to a few ulp of the accumulated radiance, not bit for bit."""
    W = H = 24
    sc = h.build_random_prims(seed=3, n_sph=12, n_rect=6, n_box=6, n_tri=30, xforms=True, W=W, H=H)
    ref = h.ReferenceRenderer(sc)
    a, _ = ref.render(W, H, 0, 2)
    orc = h.OracleScene(sc, 0, "_r2l")
    b, _ = orc.render(W, H, 0, 2)
    d = np.abs(a - b)
    scale = np.maximum(np.abs(b), 1.0)
    # paths whose random walk forks on the last-bit difference are rare; the bulk is identical
    frac_same = np.count_nonzero(a == b) / a.size
    assert frac_same > 0.97, frac_same
    assert np.median(d / scale) == 0.0


# ---------------------------------------------------------------------------------------------------------------
# golden replay: images of the reference renderer recorded in this container (tests/golden/make_golden_render.py),
# replayed against the oracle wherever the tests run (no /root/reference needed)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="no recorded reference images")
def test_oracle_matches_recorded_reference_images():
    with open(GOLDEN) as f:
        gold = json.load(f)
    table = scene_table()
    assert gold["cases"], "empty golden file"
    for case in gold["cases"]:
        build, W, H, frames, depth = table[case["scene"]]
        assert (W, H, frames, depth) == (case["W"], case["H"], case["frames"], case["max_depth"])
        orc = h.OracleScene(build(), 0, "_r2l")
        acc, _ = orc.render(W, H, 0, frames, max_depth=depth)
        bits = np.ascontiguousarray(acc, np.float32).view(np.uint32)
        # whole-image checksum plus a strided sample of raw bit patterns
        import zlib
        assert zlib.crc32(bits.tobytes()) == case["crc32"], case["scene"]
        sample = bits.reshape(-1)[:: case["stride"]]
        assert sample.tolist() == case["bits"], case["scene"]
