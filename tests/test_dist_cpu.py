"""world_size-2 gloo test of the multi-GPU host logic (runs on CPU): one weight broadcast at
init, contiguous batch shards, max-over-ranks timing, and NO other collective on the data path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sfast_b200 import dist as sdist
from sfast_b200.unet_spec import param_shapes, random_state_dict, spec_from_config
from sfast_b200.synthetic import TINY


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = sdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    spec = spec_from_config(TINY)
    shapes = param_shapes(spec)
    sd0 = random_state_dict(spec, seed=42, dtype=torch.float32) if rank == 0 else None
    sd = sdist.broadcast_state_dict(shapes, sd0, torch.float32, "cpu")
    ref = random_state_dict(spec, seed=42, dtype=torch.float32)
    same = all(torch.equal(sd[k], ref[k]) for k in shapes)
    lo, hi = sdist.shard_batch(7, rank, world)
    mx = sdist.max_over_ranks(float(rank + 1), "cpu")
    q.put((rank, same, (lo, hi), mx, len(sd)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_broadcast_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "broadcast weights differ from rank 0's"
    assert [r[2] for r in res] == [(0, 4), (4, 7)]
    assert all(r[3] == 2.0 for r in res)


def test_shard_batch_covers_everything_once():
    for gb in (1, 7, 8, 64):
        for world in (1, 2, 4, 8):
            spans = [sdist.shard_batch(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
