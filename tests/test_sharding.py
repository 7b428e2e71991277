"""N > 1 host logic on CPU: world_size-2 gloo group (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from igt_mpc_int_b200 import sharding


def test_shard_ranges_cover_batch():
    for n in (0, 1, 7, 32768, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = 10
    lo, hi = sharding.shard_range(n_total, rank, world)
    # every rank "solves" its block: result row i = problem index, status converged for even indices
    local = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1).repeat(1, 3)
    conv = int(((torch.arange(lo, hi) % 2) == 0).sum())
    t, c, it, nprob = sharding.reduce_counters(10.0 + rank, conv, 20 * (hi - lo), hi - lo)
    gathered = sharding.gather_solutions(local, n_total)
    q.put((rank, t, c, it, nprob, gathered.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, t, c, it, nprob, g in res:
        assert t == 11.0                      # max over ranks
        assert c == 5 and it == 200 and nprob == 10
        assert np.array_equal(g[:, 0], np.arange(10.0))
