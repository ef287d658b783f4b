"""torch.distributed plumbing shared by bench.py and the multi-rank tests: rendezvous, shipping
the NCCL unique id of the C-ABI communicator, max-over-ranks timing.  Works over `gloo` on CPU
(host-logic tests) and `nccl` on GPUs; the data path itself never goes through torch."""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple


def env_rank() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: Optional[str] = None):
    """Join the process group described by RANK/WORLD_SIZE/MASTER_* (no-op for a single rank)."""
    import torch
    import torch.distributed as dist
    rank, world, local_rank = env_rank()
    if world == 1 or dist.is_initialized():
        return rank, world, local_rank
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        kw["device_id"] = torch.device("cuda", local_rank)
    dist.init_process_group(backend=backend, **kw)
    return rank, world, local_rank


def share_unique_id(make_id: Callable[[], bytes]) -> Optional[bytes]:
    """Rank 0 creates the 128-byte id (ocffm_comm_unique_id), everyone receives it."""
    import torch.distributed as dist
    rank, world, _ = env_rank()
    if world == 1:
        return None
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    _, world, _ = env_rank()
    if world == 1 or not dist.is_initialized():
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    _, world, _ = env_rank()
    if world == 1 or not dist.is_initialized():
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    _, world, _ = env_rank()
    if world > 1 and dist.is_initialized():
        dist.barrier()
