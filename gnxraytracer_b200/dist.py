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


def reduce_filtered_sums(sums: torch.Tensor, dst: int = 0):
    """Gaussian film across ranks: every rank renders its sample range with GNX_FILM_GAUSSIAN_SUMS, the
    (sum L f, sum f) buffers are summed onto rank `dst`, which then resolves rgb = max(0, sum(L f) / sum(f)), alpha 1
    (the division the single-GPU GNX_FILM_GAUSSIAN film does on the device)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sums, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return sums
    v = sums.view(-1, 4)
    w = v[:, 3:4]
    v[:, :3] = torch.where(w != 0, (v[:, :3] * (1.0 / torch.where(w != 0, w, torch.ones_like(w)))).clamp_min(0.0), torch.zeros_like(v[:, :3]))
    v[:, 3] = 1.0
    return sums
