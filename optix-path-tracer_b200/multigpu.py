"""Multi-GPU host logic: one process per GPU, scene + BVH replicated, samples split by
frame index, ONE collective (sum-reduce of the float4 accumulation buffers to rank 0).

Samples are independent — seed = tea<64>(W*y+x, frame) (programs/raygen.cu:150) — and the
accumulation is a sum (raygen.cu:166), so rank g of G renders frames
[first + g*spp, first + (g+1)*spp) and the reduced buffer equals a single-GPU render of
G*spp frames up to fp32 summation order.  `.w` sums to the total sample count, which is the
integrity check after the reduce.
"""
from __future__ import annotations


def frame_range(rank: int, world: int, spp_per_rank: int, first_frame: int = 0):
    """(first_frame, nframes) rendered by `rank` (weak scaling: fixed work per rank)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return first_frame + rank * spp_per_rank, spp_per_rank


def split_frames(total_frames: int, rank: int, world: int, first_frame: int = 0):
    """(first_frame, nframes) for a strong-scaling split of `total_frames` (config C5)."""
    base, rem = divmod(total_frames, world)
    n = base + (1 if rank < rem else 0)
    start = first_frame + rank * base + min(rank, rem)
    return start, n


def reduce_accum(accum, dst: int = 0):
    """Sum-reduce the accumulation tensors to `dst` (NCCL on GPUs; gloo in CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst)
    return accum


def check_sample_count(accum, expected_samples: float, rtol: float = 1e-6) -> bool:
    """`.w` of every pixel must equal the number of samples that were merged."""
    w = accum[..., 3]
    return bool(abs(float(w.min()) - expected_samples) <= rtol * expected_samples and
                abs(float(w.max()) - expected_samples) <= rtol * expected_samples)
