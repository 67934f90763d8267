"""Multi-GPU plumbing: chains are independent, so ranks own contiguous chain ranges and the hot loop
has no communication.  The only collective is the end-of-run gather of traced scalars for
R-hat / ESS (NCCL all-gather on GPUs; gloo in the CPU tests)."""

import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous, balanced chain range [lo, hi) owned by `rank`."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_chains(local, device=None):
    """Gather per-rank arrays [chains_local, ...] (equal trailing shape, ragged first axis allowed)
    into [chains_total, ...] on every rank, ordered by rank."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[dist.get_rank()] = local.shape[0]
    dist.all_reduce(counts)
    counts = counts.cpu().tolist()
    mx = max(counts)
    pad = np.zeros((mx,) + local.shape[1:])
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(outs, counts)], axis=0)
