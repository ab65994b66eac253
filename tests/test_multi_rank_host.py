"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the Line sharding and the
fan-in reduce plumbing (the GPU data path itself is covered by -m gpu tests and bench.py --gpus)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _oracle as orc
from pipe_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        n_lines, ch, frames = 4, 8, 256
        mine = shard.assign_lines(n_lines, world)[rank]
        # every rank produces its Lines' outputs (here: the synthetic source itself), sums them locally,
        # then the fan-in reduce adds the per-rank partial sums on rank 0
        partial = np.zeros((frames, ch))
        for line in mine:
            partial += orc.source_fill(0, frames * ch, line=line).reshape(frames, ch)
        t = torch.from_numpy(partial.copy())
        shard.fan_in_reduce(t, dst=0)
        if rank == 0:
            ref = orc.mix_sum([orc.source_fill(0, frames * ch, line=l).reshape(frames, ch) for l in range(n_lines)])
            q.put(("sum_err", float(np.abs(t.numpy() - ref).max())))
        counts = [None] * world
        dist.all_gather_object(counts, len(mine))
        if rank == 0:
            q.put(("counts", counts))
    finally:
        dist.destroy_process_group()


def test_assign_lines_round_robin():
    assert shard.assign_lines(8, 8) == [[i] for i in range(8)]
    assert shard.assign_lines(4, 2) == [[0, 2], [1, 3]]
    assert shard.assign_lines(3, 4) == [[0], [1], [2], []]
    with pytest.raises(ValueError):
        shard.assign_lines(1, 0)


def test_world_size_2_gloo_sharding_and_fan_in():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=10) for _ in range(2))
    assert got["counts"] == [2, 2]
    assert got["sum_err"] < 1e-12
