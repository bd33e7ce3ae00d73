"""Multi-GPU plumbing: one process per GPU (torchrun), the scene replicated, the samples of every
pixel dealt out as contiguous ranges, one framebuffer sum-reduce at the end.

The path itself has no exchange step, so the only collective is `reduce(sum)` of the W*H*4 float
framebuffer to rank 0 — NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def sample_range(total_spp: int, rank: int, world: int):
    """Contiguous sample range [first, first + count) of `rank` (strong scaling of a fixed spp)."""
    base, rem = divmod(total_spp, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def weak_sample_range(spp_per_rank: int, rank: int):
    """Weak scaling: every rank renders the full per-GPU workload, on its own slice of the sequence."""
    return rank * spp_per_rank, spp_per_rank


def reduce_framebuffer(fb: torch.Tensor, dst: int = 0):
    """Sum the partial framebuffers (each already divided by the TOTAL spp) onto rank `dst`.
    Alpha is restored to 1 afterwards."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(fb, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() == dst:
            fb.view(-1, 4)[:, 3] = 1.0
    return fb
