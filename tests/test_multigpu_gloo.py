"""N>1 host logic on CPU: two gloo ranks split the frames, render their share (with the CPU
oracle standing in for the GPU renderer — this tests the sharding + reduce logic, not the
kernels) and sum-reduce the accumulation buffers; the result must equal one process
rendering all frames."""
import os
import subprocess
import sys

import numpy as np

import helpers as h

WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import helpers as h
mg = __import__("importlib").import_module("optix_path_tracer_b200.multigpu")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
W = H = 24
sc = h.build_cornell(W, H)
orc = h.OracleScene(sc)
first, n = mg.frame_range(rank, world, 3)
acc, _ = orc.render(W, H, first, n)
t = torch.from_numpy(acc)
mg.reduce_accum(t, dst=0)
first2, n2 = mg.split_frames(7, rank, world)          # strong split of 7 frames
acc2, _ = orc.render(W, H, first2, n2)
t2 = torch.from_numpy(acc2)
mg.reduce_accum(t2, dst=0)
if rank == 0:
    assert mg.check_sample_count(t, 3 * world)
    assert mg.check_sample_count(t2, 7)
    np.save(sys.argv[2], np.stack([t.numpy(), t2.numpy()]))
dist.destroy_process_group()
'''


def test_two_rank_sample_split_and_reduce(tmp_path):
    h.oracle()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "merged.npy")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), h.ROOT, out],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    merged = np.load(out)
    sc = h.build_cornell(24, 24)
    orc = h.OracleScene(sc)
    single, _ = orc.render(24, 24, 0, 6)
    assert np.allclose(merged[0], single, rtol=1e-6, atol=1e-6)
    single7, _ = orc.render(24, 24, 0, 7)
    assert np.allclose(merged[1], single7, rtol=1e-6, atol=1e-6)


def test_frame_partitions():
    mg = __import__("importlib").import_module("optix_path_tracer_b200.multigpu")
    for world in (1, 2, 4, 8):
        got = []
        for rnk in range(world):
            f, n = mg.split_frames(4096 + 3, rnk, world, first_frame=10)
            got.extend(range(f, f + n))
        assert got == list(range(10, 10 + 4099))
        assert [mg.frame_range(rnk, world, 256)[0] for rnk in range(world)] == [256 * k for k in range(world)]
