"""The one-process-per-GPU plumbing of bench.py on CPU: world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import smallvcm_b200 as sv
from smallvcm_b200 import dist as D


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, results):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    assert D.env_rank() == (rank, rank, world)
    D.init_process_group("gloo")
    payload = bytes(range(128)) if rank == 0 else None
    got = D.broadcast_bytes(payload, 128, 0)
    assert got == bytes(range(128))
    assert D.max_over_ranks(1.0 + rank) == float(world)
    assert D.sum_over_ranks(1.0 + rank) == world * (world + 1) / 2
    D.barrier()
    # the rows this rank would trace under the library's shard map
    lib = sv.load_library()
    w, h = 12, 7
    rows = D.shard_rows(h, rank, world)
    assert lib.svcm_shard_local_paths(w, h, rank, world) == len(rows) * w
    for y in rows:
        assert lib.svcm_shard_owner(w, world, y * w + 3) == rank
    results[rank] = len(rows)
    dist.destroy_process_group()


def test_gloo_world_size_2():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert sum(results.values()) == 7
