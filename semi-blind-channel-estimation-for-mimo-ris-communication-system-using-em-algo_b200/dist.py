"""Multi-GPU plumbing: trials are sharded across ranks (one process per GPU,
disjoint seeds), there is NO data-path collective; the only exchange is one sum
all-reduce of the per-sweep-point accumulators at the end of a sweep
(SURVEY.md section 8e).  Works with any initialised torch.distributed backend
(nccl on the GPU box, gloo in the CPU tests) and degrades to a no-op when the
process group is not initialised."""
from __future__ import annotations

import os

import numpy as np


def world():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (RANK,
    WORLD_SIZE, MASTER_ADDR, MASTER_PORT, LOCAL_RANK).  Returns (rank, world, local_rank)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=ws)
    return rank, ws, local


def shard_trials(n_trials: int, rank: int, world_size: int):
    """Contiguous block partition of trial indices [0, n_trials): returns (start, stop)."""
    base, rem = divmod(n_trials, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_sum(acc: np.ndarray, device=None) -> np.ndarray:
    """Sum a small float64 accumulator array over all ranks (latency-only payload)."""
    rank, ws = world()
    if ws == 1:
        return acc
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(acc, dtype=np.float64))
    if dist.get_backend() == "nccl":
        # the rank's own GPU: an int is a CUDA ordinal; None falls back to the current device
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        elif isinstance(device, int):
            device = torch.device("cuda", device)
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def barrier():
    rank, ws = world()
    if ws > 1:
        import torch.distributed as dist

        dist.barrier()
