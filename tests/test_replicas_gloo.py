"""The N>1 host logic (instance sharding + the single all-gather of trajectories) on CPU with the gloo
backend, world_size 2 and 3 -- the GPU box runs the same code over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eqvio_b200_replicas_shim import gather_trajectories, shard_instances


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _traj(i, T):
    rng = np.random.default_rng(i)
    return rng.standard_normal((T, 11))


def _worker(rank, world, port, num_instances, T, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_instances(num_instances, world, rank)
    local = {i: _traj(i, T) for i in mine}
    out = gather_trajectories(local, num_instances)
    ok = all(np.array_equal(out[i], _traj(i, T)) for i in range(num_instances))
    q.put((rank, mine, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,num_instances", [(2, 5), (3, 7), (2, 2)])
def test_shard_and_gather(world, num_instances):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_instances, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = sorted(i for _, mine, _ in res for i in mine)
    assert covered == list(range(num_instances)), "every instance belongs to exactly one rank"
    assert all(ok for _, _, ok in res), "every rank holds every trajectory after the all-gather"


def test_shard_is_contiguous_and_balanced():
    for n in (0, 1, 7, 16, 128):
        for w in (1, 2, 3, 8):
            blocks = [shard_instances(n, w, r) for r in range(w)]
            flat = [i for b in blocks for i in b]
            assert flat == list(range(n))
            sizes = [len(b) for b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_instances(4, 2, 2)


def test_gather_without_process_group():
    local = {0: _traj(0, 3), 1: _traj(1, 3)}
    out = gather_trajectories(local, 2)
    assert np.array_equal(out[1], _traj(1, 3))
