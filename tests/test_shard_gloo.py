"""world_size-2 (and 3) gloo tests of the multi-GPU sharding logic on CPU: contiguous item ranges,
byte-aligned shard starts, all-gather of the per-shard verify bitmaps into the global bitmap."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

shard = importlib.import_module("ring-zk_b200.shard")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.shard_range(B, rank, world)
        # "verify" result of item i is a fixed pseudo-random function of i: every rank computes its shard only
        truth = (np.arange(B) * 2654435761 % 7) != 0
        local = truth[lo:hi]
        bm = torch.from_numpy(np.packbits(local, bitorder="little")) if hi > lo else torch.zeros(0, dtype=torch.uint8)
        full = shard.gather_bitmaps(bm, B)
        got = shard.bitmap_to_bool(full, B)
        q.put((rank, lo, hi, bool((got == truth).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 1 << 12), (2, 1003), (3, 77), (2, 5)])
def test_gather_bitmaps_gloo(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for *_, ok in res)
    # ranges tile [0, B) and start on byte boundaries
    assert res[0][1] == 0 and res[-1][2] == B
    for (r0, lo0, hi0, _), (r1, lo1, hi1, _) in zip(res, res[1:]):
        assert hi0 == lo1 and lo1 % 8 == 0 or lo1 == B


def test_shard_range_single():
    assert shard.shard_range(65536, 0, 1) == (0, 65536)
    assert shard.shard_range(65536, 7, 8) == (57344, 65536)
    assert shard.shard_range(10, 1, 2) == (8, 10)
