"""N>1 host logic on CPU: two gloo ranks shard a batch by index, time-reduce with MAX and gather results."""
import importlib
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

shard = importlib.import_module("teaser-plusplus_b200.shard")
capi = importlib.import_module("teaser-plusplus_b200.capi")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_problems, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard.shard_indices(n_problems, rank, world)
    rec = np.zeros(len(idx), dtype=capi.SOLUTION_DTYPE)
    rec["clique_size"] = idx * 10 + 1  # stand-in for solved records (no GPU here)
    rec["scale"] = 1.0
    t = shard.max_over_ranks(5.0 + 3.0 * rank, dist)
    full = shard.gather_solutions(rec, idx, n_problems, dist)
    dist.barrier()
    if rank == 0:
        q.put((t, full["clique_size"].tolist()))
    dist.destroy_process_group()


def test_two_rank_sharding():
    world, n_problems = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_problems, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, sizes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert t == 8.0  # MAX over ranks
    assert sizes == [b * 10 + 1 for b in range(n_problems)]


def test_shard_indices_partition():
    for world in (1, 2, 4, 8):
        allidx = np.concatenate([shard.shard_indices(4096, r, world) for r in range(world)])
        assert sorted(allidx.tolist()) == list(range(4096))
        sizes = [len(shard.shard_indices(4096, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_indices(10, 2, 2)
