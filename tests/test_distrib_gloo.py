"""N > 1 path on the CPU: two gloo ranks run the bench's weak-scaling aggregation (one independent sequence per rank,
max-over-ranks timing, sum of frames) -- and two independent ORACLE sequences give different results per seed, i.e. the
ranks really process different data."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from dynamicfusion_b200 import distrib, synth
    from oracle import orc_pipe
    r, w = distrib.init("gloo")
    assert (r, w) == (rank, world)
    # each rank: its own short sequence through the CPU oracle pipeline (stand-in for the per-GPU KinFu)
    p = orc_pipe.default_params(0, dim=32, size=1.0)
    p.cloud_capacity = 100000
    p.flags = 1
    k = orc_pipe.KinFu(p)
    seed = distrib.sequence_seed(rank)
    fused = 0
    for t in range(3):
        fused += int(k(synth.umbrella_depth(t, seed=seed)))
    checksum = int(k.buffer("volume").astype(np.uint64).sum())
    k.close()
    distrib.barrier()
    local_ms = 100.0 * (rank + 1)                       # rank 1 is the slow one
    max_ms, total, fmin = distrib.aggregate(local_ms, fused)
    out[rank] = (max_ms, total, fmin, checksum, distrib.throughput(total, max_ms))
    torch.distributed.destroy_process_group()


def test_two_rank_weak_scaling_aggregate():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0[:3] == r1[:3] == (200.0, 4, 2)            # max over ranks, 2 fused frames per rank
    assert r0[3] != r1[3], "ranks must process different sequences (seed = rank)"
    assert abs(r0[4] - 4 / 0.2) < 1e-9                  # whole-job frames/s = total frames / slowest rank


def test_single_process_aggregate_is_identity():
    from dynamicfusion_b200 import distrib
    assert distrib.aggregate(12.5, 7) == (12.5, 7, 7)
