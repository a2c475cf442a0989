"""CPU, gloo, world_size 2: the host-side logic of the row partition (block ranges and the
row exchange in its even and ragged forms).  The kernels themselves need a GPU; the same
exchange code runs under NCCL in tests/multi_gpu_check.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deep_cbrs_amar_renaissance_b200.distributed import block_ranges, exchange_rows


def test_block_ranges_cover_every_row_once():
    for sizes, world in [([10, 5], 4), ([8, 8], 2), ([6040, 3706, 17554], 8), ([3], 4)]:
        ranges = block_ranges(sizes, world)
        owned = np.zeros(sum(sizes), int)
        for rank_ranges in ranges:
            assert len(rank_ranges) == len(sizes)
            for a, b in rank_ranges:
                owned[a:b] += 1
        assert (owned == 1).all()
        base = 0
        for t, n in enumerate(sizes):  # every block stays inside its node type
            for rank_ranges in ranges:
                a, b = rank_ranges[t]
                assert base <= a <= b <= base + n
            base += n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, sizes, width, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = sum(sizes)
        ranges = block_ranges(sizes, world)
        truth = torch.arange(n * width, dtype=torch.float32).reshape(n, width)
        x = torch.full((n, width), -1.0)
        for a, b in ranges[rank]:
            x[a:b] = truth[a:b]  # each rank authors only its own blocks
        exchange_rows(x, ranges)
        ok = torch.equal(x, truth)
        # partial exchange (items only): user rows of the other rank stay untouched
        y = torch.full((n, width), -1.0)
        for a, b in ranges[rank]:
            y[a:b] = truth[a:b]
        exchange_rows(y, [[rg[1]] for rg in ranges])
        ua, ub = ranges[1 - rank][0]
        ok = ok and torch.equal(y[sizes[0]:], truth[sizes[0]:]) and bool((y[ua:ub] == -1).all())
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sizes", [[8, 6], [9, 5]])  # even blocks -> all_gather; ragged -> broadcasts
def test_exchange_rows_gloo_world2(sizes):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, sizes, 3, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
