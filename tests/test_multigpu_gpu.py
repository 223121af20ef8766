"""Multi-GPU below the C ABI (needs >= 2 devices; run with `gpurun --gpus 2`):
  * one process, several devices: pt_renderer_create_multi (replicas + one renderer per device + ncclReduce on read);
  * one process per device: pt_comm_* under torchrun;
  * the CLI's --gpus.
Samples are independent (seed = tea<64>(pixel, frame), programs/raygen.cu:150) and accumulation is a sum
(raygen.cu:166): the merged buffer equals the single-device buffer up to fp32 summation order, .w == spp."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as h

pt = h.pt
pytestmark = pytest.mark.gpu


def n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(n_devices() < 2, reason="needs two CUDA devices")


def test_multi_handle_over_one_device(lib):
    """pt_renderer_create_multi with a single device: the same fan-out code path (handle -> sub-renderer, split, merge
    state, statistics) without a second GPU and without NCCL; must equal the plain renderer bit for bit."""
    W, H = 128, 96
    sc = h.build_spheres(22, W=W, H=H, lib=lib)
    sc.commit()
    plain = pt.Renderer(sc, W, H)
    multi = pt.Renderer(sc, W, H, devices=[0])
    plain.render(2, 1)
    multi.render(2, 1)
    a, b = plain.read_accum(), multi.read_accum()
    assert (b[..., 3] == 1).all() and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    multi.render(3, 5)
    plain.render(3, 5)
    assert (multi.read_accum()[..., 3] == 6).all()
    assert np.allclose(plain.read_accum(), multi.read_accum(), rtol=1e-5, atol=1e-5)
    st = multi.stats()
    assert st.samples == W * H * 5 and st.ms_reduce == 0 and st.closest_rays == plain.stats().closest_rays
    assert np.array_equal(plain.read_display(), multi.read_display())


@needs2
def test_one_handle_two_devices_equals_one_device(lib):
    W, H = 160, 96
    sc = h.build_spheres(22, W=W, H=H, lib=lib)
    sc.commit()
    single = pt.Renderer(sc, W, H)
    multi = pt.Renderer(sc, W, H, devices=[0, 1])
    # two frames, one per device: every pixel is the sum of exactly two samples on both sides -> identical bits
    single.render(0, 2)
    multi.render(0, 2)
    a, b = single.read_accum(), multi.read_accum()
    assert (b[..., 3] == 2).all()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # an odd number of frames, then a progressive continuation (merge in between)
    single.render(0, 7, clear=True)
    multi.render(0, 7, clear=True)
    a, b = single.read_accum(), multi.read_accum()
    assert (b[..., 3] == 7).all() and np.allclose(a, b, rtol=1e-5, atol=1e-5)
    st = multi.stats()
    assert st.samples == W * H * 7 and st.closest_rays == single.stats().closest_rays
    assert st.ms_reduce > 0
    single.render(7, 6)
    multi.render(7, 6)
    a, b = single.read_accum(), multi.read_accum()
    assert (b[..., 3] == 13).all() and np.allclose(a, b, rtol=1e-5, atol=1e-5)
    assert np.array_equal(single.read_display(), multi.read_display()) or \
        np.abs(single.read_display().astype(int) - multi.read_display().astype(int)).max() <= 1
    # checkpoint / resume through the multi handle
    multi.write_accum(a)
    multi.render(13, 3)
    single.render(13, 3)
    assert np.allclose(single.read_accum(), multi.read_accum(), rtol=1e-5, atol=1e-5)


@needs2
def test_replicas_follow_the_scene(lib):
    """camera changes and re-commits of the caller's scene reach the replicas on the other devices"""
    W, H = 96, 64
    sc = h.build_cornell(W, H, lib)
    sc.commit()
    multi = pt.Renderer(sc, W, H, devices=[0, 1])
    multi.render(0, 2)
    first = multi.read_accum()
    cam = sc.desc().camera
    cam2 = pt.Camera.from_buffer_copy(bytes(cam))
    cam2.origin[0] += 60.0
    sc.set_camera(cam2)
    multi.render(0, 2, clear=True)
    moved = multi.read_accum()
    single = pt.Renderer(sc, W, H)
    single.render(0, 2)
    assert np.array_equal(moved.view(np.uint32), single.read_accum().view(np.uint32))
    assert not np.array_equal(first, moved)
    sc.add_spheres([pt.Sphere(pt.F3(278, 100, 278), 60.0)], [1])
    with pytest.raises(pt.PtError):
        multi.render(0, 2, clear=True)
    sc.commit()
    multi.render(0, 2, clear=True)
    single.render(0, 2, clear=True)
    assert np.array_equal(multi.read_accum().view(np.uint32), single.read_accum().view(np.uint32))


@needs2
def test_multi_argument_errors(lib):
    sc = h.build_cornell(32, 32, lib)
    sc.commit()
    with pytest.raises(pt.PtError):
        pt.Renderer(sc, 32, 32, devices=[0, 0])
    with pytest.raises(pt.PtError):
        pt.Renderer(sc, 32, 32, devices=[1])           # the scene's own device is missing
    with pytest.raises(pt.PtError):
        pt.Renderer(sc, 32, 32, devices=[0, 99])


WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import helpers as h
pt = h.pt
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
lib = pt.load()
W, H, spp = 160, 96, 9
sc = h.build_spheres(22, W=W, H=H, lib=lib)
sc.commit()
r = pt.Renderer(sc, W, H)
idt = torch.zeros(pt.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(pt.Comm.unique_id(lib)), dtype=torch.uint8))
dist.broadcast(idt, src=0)
comm = pt.Comm(bytes(idt.cpu().numpy().tobytes()), rank, world, lib)
first, n = pt.split_frames(spp, rank, world, lib=lib)
r.render(first, n, clear=True)
comm.reduce(r, 0)
if rank == 0:
    assert r.stats().ms_reduce > 0
    np.save(sys.argv[2], r.read_accum())
comm.close()
dist.destroy_process_group()
'''


@needs2
def test_one_process_per_device_with_pt_comm(lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "merged.npy")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script), h.ROOT, out],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    merged = np.load(out)
    W, H, spp = 160, 96, 9
    sc = h.build_spheres(22, W=W, H=H, lib=lib)
    sc.commit()
    single = pt.Renderer(sc, W, H)
    single.render(0, spp)
    assert (merged[..., 3] == spp).all()
    assert np.allclose(merged, single.read_accum(), rtol=1e-5, atol=1e-5)


@needs2
def test_cli_two_gpus_png_equals_one_gpu_png(tmp_path):
    """ptb200_render --gpus 2 (Context::setDevices in the facade): the PNG equals the single-GPU PNG within 1/255"""
    PIL = pytest.importorskip("PIL.Image")
    exe = os.path.join(h.PKG_DIR, "csrc", "ptb200_render")
    outs = []
    for gpus in (1, 2):
        name = str(tmp_path / f"g{gpus}")
        r = subprocess.run([exe, "--scene", "2", "--width", "128", "--height", "128", "--samples", "32", "--file", name,
                            "--gpus", str(gpus)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(np.asarray(PIL.open(name + ".png")).astype(int))
    assert outs[0].shape == outs[1].shape
    assert np.abs(outs[0] - outs[1]).max() <= 1
