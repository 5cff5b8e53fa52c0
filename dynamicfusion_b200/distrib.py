"""Multi-GPU plumbing for the batched-sequences configuration (BASELINE.json config 5).

The path shards by SEQUENCE: every rank owns an independent KinFu (volume + pyramids + node table), there is no data-path
collective.  torch.distributed (NCCL on the GPU box, gloo in the CPU tests) is used only for the start/stop barrier, the
max-over-ranks timing and the sum of fused frames -- a few bytes per run."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def sequence_seed(rank: int, base_seed: int = 0) -> int:
    """sequence i -> rank i (one sequence per GPU); seeds are distinct so the ranks do not run identical data"""
    return base_seed + rank


def init(backend: str, device: torch.device | None = None) -> tuple[int, int]:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def barrier(device: torch.device | None = None) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def aggregate(local_ms: float, local_frames: int, device: torch.device | None = None) -> tuple[float, int, int]:
    """weak-scaling aggregate: (max elapsed ms over ranks, total fused frames, min fused frames per rank)"""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(local_ms), int(local_frames), int(local_frames)
    dev = device if device is not None else torch.device("cpu")
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    f = torch.tensor([local_frames], dtype=torch.int64, device=dev)
    fmin = f.clone()
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(fmin, op=dist.ReduceOp.MIN)
    return float(t.item()), int(f.item()), int(fmin.item())


def throughput(total_frames: int, max_ms: float) -> float:
    return total_frames / (max_ms * 1e-3)
