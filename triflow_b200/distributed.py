"""Multi-GPU: one process per GPU, ensembles sharded by member, no data-path
collective.  ``torch.distributed`` is used for the rendezvous,
the barrier around timed regions and the final gather only."""

import os

import numpy as np


def world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard(n_members, rank, world_size):
    """Contiguous block of members ``[lo, hi)`` owned by ``rank``."""
    base, rem = divmod(int(n_members), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init(backend=None):
    """Join the process group described by the torchrun environment."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1 or dist.is_initialized():
        return rank, ws
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        kw["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend=backend, rank=rank, world_size=ws, **kw)
    return rank, ws


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_members(local, n_members):
    """Final gather of per-member rows ``local`` (n_local, width) to rank 0 in
    member order.  Returns the full array on rank 0, ``None`` elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(local)
    rank, ws = dist.get_rank(), dist.get_world_size()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    width = local.shape[1]
    sizes = [shard(n_members, r, ws) for r in range(ws)]
    nmax = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((nmax, width), dtype=torch.float64, device=dev)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    out = [torch.empty_like(buf) for _ in range(ws)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return np.concatenate([o[: hi - lo].cpu().numpy() for o, (lo, hi) in zip(out, sizes)])
