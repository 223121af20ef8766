"""bench.py --impl reference: the CPU arm prints ONE JSON line with the contract's keys, for both of its
back ends (the reference's own programs in the emulated pipeline, kind "reference"; the oracle port, kind
"port").  Runs on the CPU; a few seconds per case."""
import json
import os
import subprocess
import sys

import pytest

import helpers as h

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_arm(env_extra):
    env = dict(os.environ, PTB200_REFERENCE_SECONDS="3", **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def check_line(d, kind):
    assert d["impl"] == "reference" and d["metric"] == "Msamples/s" and d["unit"] == "Msamples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C1 ")
    cb = d["cpu_baseline"]
    assert cb["kind"] == kind and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_oracle_port():
    check_line(run_arm({"PTB200_REFERENCE_PORT": "1"}), "port")


@pytest.mark.skipif(h.ReferenceRenderer.lib() is None, reason="oracle/_ref/libref_render.so not built")
def test_reference_arm_reference_programs():
    d = run_arm({})
    check_line(d, "reference")
    assert d["oracle_port"]["kind"] == "port" and d["oracle_port"]["value"] > 0
    # Cornell box: the mean linear radiance of the converged image is 0.039 (oracle); proves the bands rendered
    assert 0.03 < d["mean_radiance"] < 0.05
