"""Host-side set-up against the reference's own host headers (host_includes/camera.hpp and host_common.hpp's rnd(),
compiled unmodified: oracle/ref_host_probe.cpp): the seven camera vectors + lens radius + shutter times that the
ray-generation program reads, bit-exact for the test helpers' make_camera (which the facade's camera.hpp is checked
against in test_facade.py), and the mt19937(0) stream the scene functions draw from."""
import json
import os
import sys

import numpy as np
import pytest

import helpers as h

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_host as G  # noqa: E402

pt = h.pt
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_host_vectors.json")


def mine(c):
    a = G.unbits(c["args"])
    cam = h.make_camera(a[0:3], a[3:6], a[6:9], *[float(x) for x in a[9:15]])
    return np.asarray([*cam.origin, *cam.lower_left_corner, *cam.horizontal, *cam.vertical, *cam.u, *cam.v, *cam.w,
                       cam.lens_radius, cam.time0, cam.time1], dtype=np.float32)


def test_camera_matches_reference_camera_hpp_golden():
    gold = json.load(open(GOLD))
    for c in gold["cameras"]:
        assert G.bits(mine(c)) == c["out"], (list(G.unbits(c["args"])), mine(c), G.unbits(c["out"]))
    # std::mt19937(0) + uniform_real_distribution<float>(0, 1): the stream InOneWeekend etc. consume
    r = G.unbits(gold["rnd_first_64"])
    assert ((r >= 0) & (r < 1)).all() and abs(float(r[0]) - 0.5488135) < 1e-6


def test_camera_matches_reference_camera_hpp_live():
    r = G.probe()
    if r is None:
        pytest.skip("reference tree not present (the recorded vectors cover this)")
    for c in G.camera_cases(np.random.default_rng(3)):
        assert G.bits(mine(c)) == G.bits(G.run_reference(r, c))
