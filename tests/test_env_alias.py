"""Alias tables of the HDRI importance sampler (csrc/env_alias_core.h; NEW relative to the reference, whose miss
program programs/miss.cu:64-95 only evaluates the map).

The construction is the closed-form "sweep" in 32.32 fixed point: given the float pdf of every texel, the table is an
EXACT integer function of it.  So the checks are exact too:
  * CPU: the table the library describes for an uncommitted scene (serial host mirror of the device bodies) equals an
    independent numpy restatement of the same rule, is a valid alias table, and reproduces the distribution;
  * GPU (-m gpu): the table built on the device at commit equals the host mirror bit for bit (pdf, probability, alias).
"""
import numpy as np
import pytest

import helpers as h

pt = h.pt
ONE = 1 << 32


def tables(sc):
    d = sc.desc()
    n = d.n_images and d.env_n
    arr = lambda p, dt: np.ctypeslib.as_array(p, shape=(n,)).view(dt).copy() if n else np.zeros(0, dt)
    return arr(d.env_alias_prob, np.float32), arr(d.env_alias_idx, np.int32), arr(d.env_texel_pdf, np.float32), d


def env_scene(img, lib=None):
    s = pt.Scene(lib)
    s.set_textures([h.tex_const(0.7, 0.7, 0.7)])
    s.set_materials([h.material(pt.MAT_LAMBERTIAN, 0)])
    s.add_spheres([pt.Sphere(pt.F3(0, 0, 0), 1.0)], [0])
    s.set_camera(h.make_camera([0, 0.5, -6], [0, 0, 0], [0, 1, 0], 40.0, 1.0, 0.0, 6.0))
    i = s.add_image(img)
    s.set_miss(pt.Miss(pt.MISS_HDR, pt.F3(0, 0, 0), pt.F3(0, 0, 0), i, 1, 1))
    return s


def images():
    rng = np.random.default_rng(5)
    sun = h.synthetic_hdri(64, 32)
    spiky = np.zeros((16, 32, 4), np.float32)
    spiky[..., :3] = rng.random((16, 32, 3)) ** 8 * 50.0        # a few very heavy texels
    flat = np.full((8, 16, 4), 0.25, np.float32)
    black = np.zeros((4, 8, 4), np.float32)                      # only the 1 % floor remains
    one = np.zeros((4, 8, 4), np.float32)
    one[2, 3, :3] = 1000.0                                       # one texel carries almost everything
    return {"sun": sun, "spiky": spiky, "flat": flat, "black": black, "one_texel": one}


def numpy_alias(pdf):
    """independent restatement of env_alias_item: exact integer arithmetic on q = floor(pdf * n * 2^32)"""
    n = len(pdf)
    q = np.floor(pdf.astype(np.float64) * n * float(ONE)).astype(np.uint64)
    light = q <= ONE
    li, hi = np.nonzero(light)[0], np.nonzero(~light)[0]
    deficit = (np.uint64(ONE) - q[li])
    df = np.concatenate([[0], np.cumsum(deficit, dtype=np.uint64)]).astype(np.uint64)       # exclusive, a + 1 entries
    ex = np.cumsum(q[hi] - np.uint64(ONE), dtype=np.uint64)                                  # inclusive
    prob = np.ones(n, np.float32)
    alias = np.arange(n, dtype=np.int32)
    j = np.searchsorted(ex, df[:-1], side="right")                # first j with ex[j] > df[r]
    ok = j < len(hi)
    prob[li[ok]] = (q[li[ok]].astype(np.float64) / float(ONE)).astype(np.float32)
    alias[li[ok]] = hi[j[ok]]
    r = np.searchsorted(df, ex, side="left")                      # first r with df[r] >= ex[j]
    okh = (r <= len(li)) & (np.arange(len(hi)) + 1 < len(hi))
    rr = np.minimum(r, len(li))
    num = (np.uint64(ONE) + ex - df[rr]).astype(np.float64)
    prob[hi[okh]] = (num[okh] / float(ONE)).astype(np.float32)
    alias[hi[okh]] = hi[np.nonzero(okh)[0] + 1]
    return prob, alias, q


@pytest.mark.parametrize("name", list(images()))
def test_host_tables_are_exact_valid_and_reproduce_the_distribution(lib, name):
    img = images()[name]
    prob, alias, pdf, d = tables(env_scene(img, lib))
    n = img.shape[0] * img.shape[1]
    assert d.env_n == n and len(prob) == n
    # the pdf: luminance of the radiance the miss program returns (2 x texel) x sin(theta) + 1 % floor, normalised
    hh, ww = img.shape[:2]
    st = np.sin(np.pi * (np.arange(hh) + 0.5) / hh)[:, None]
    w = 2.0 * (0.2126 * img[..., 0].astype(np.float64) + 0.7152 * img[..., 1] + 0.0722 * img[..., 2]) * st
    fl = 0.01 * w.sum() / n if w.sum() > 0 else 1.0
    w = w + fl * st
    want = (w / w.sum()).reshape(-1)
    assert np.allclose(pdf, want, rtol=1e-5, atol=1e-12)
    assert abs(float(pdf.astype(np.float64).sum()) - 1.0) < 1e-5
    # exactness: the table is an integer function of the float pdf
    p2, a2, q = numpy_alias(pdf)
    assert np.array_equal(alias, a2)
    assert np.array_equal(prob.view(np.uint32), p2.view(np.uint32))
    # validity
    assert ((prob >= 0) & (prob <= 1)).all() and ((alias >= 0) & (alias < n)).all()
    # distribution: mass of texel k = prob[k] + sum over buckets aliased to k of (1 - prob)
    mass = prob.astype(np.float64).copy()
    np.add.at(mass, alias, 1.0 - prob.astype(np.float64))
    assert np.allclose(mass / n, pdf.astype(np.float64), rtol=2e-4, atol=2e-6 / n)
    # every light bucket keeps its own mass exactly, every aliased bucket points at a heavy texel
    light = q <= ONE
    moved = alias != np.arange(n)
    assert (~light[alias[moved]]).all()


def test_tables_absent_without_env_nee(lib):
    s = h.build_mesh_scene(8, 16, 16, lib, env_nee=False)
    assert s.desc().env_n == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sun", "spiky", "one_texel"])
def test_device_tables_equal_the_host_mirror_bit_for_bit(lib, name):
    img = images()[name]
    if name == "sun":
        img = h.synthetic_hdri(2048, 1024)      # the benchmark's map: 2 M texels
    host = tables(env_scene(img, lib))[:3]       # never committed: serial mirror
    sc = env_scene(img, lib)
    sc.commit()                                  # built on the device
    dev = tables(sc)[:3]
    for a, b, what in zip(host, dev, ("probability", "alias", "pdf")):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), what
