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

from deep_cbrs_amar_renaissance_b200.distributed import balanced_ranges, block_ranges, exchange_rows


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


def _adversarial_rowptr(n_users, n_items, seed=0):
    """users: ~uniform short rows; items in POPULARITY order (np.unique ids of a catalogue sorted by popularity):
    the first items hold most of the edges - the case equal row counts cannot balance"""
    rng = np.random.RandomState(seed)
    deg_u = rng.randint(20, 40, size=n_users)
    w = 1.0 / np.arange(1, n_items + 1)
    deg_i = np.maximum(1, np.floor(w / w.sum() * deg_u.sum())).astype(np.int64)
    return torch.from_numpy(np.concatenate([[0], np.cumsum(np.concatenate([deg_u, deg_i]))]).astype(np.int64))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_balanced_ranges_equalise_edges_on_an_unscattered_graph(world):
    n_users, n_items = 4000, 1000
    rp = _adversarial_rowptr(n_users, n_items)
    sizes = [n_users, n_items]
    bal = balanced_ranges(sizes, world, rp)
    owned = np.zeros(sum(sizes), int)
    for rank_ranges in bal:
        for a, b in rank_ranges:
            owned[a:b] += 1
    assert (owned == 1).all()

    def edges(ranges, t):
        return np.array([int(rp[b] - rp[a]) for a, b in (rg[t] for rg in ranges)], float)
    for t in range(2):  # per node type: the sparse kernel runs one launch per owned block
        e = edges(bal, t)
        heaviest_row = int((rp[1:] - rp[:-1])[sum(sizes[:t]):sum(sizes[:t + 1])].max())
        assert e.max() - e.min() <= 2 * heaviest_row + 1, (t, e)   # as even as whole rows allow
    items_by_rows = edges(block_ranges(sizes, world), 1)
    assert items_by_rows.max() / max(items_by_rows.min(), 1) > 3       # what equal row counts would have given
    # blocks stay inside their node type and ascend
    for rank_ranges in bal:
        assert 0 <= rank_ranges[0][0] <= rank_ranges[0][1] <= n_users <= rank_ranges[1][0] <= rank_ranges[1][1] <= n_users + n_items


def test_balanced_ranges_edge_cases():
    rp = torch.tensor([0, 0, 0, 5, 5], dtype=torch.int64)      # one non-empty row among four
    r = balanced_ranges([4], 3, rp)
    assert sorted(b - a for (a, b), in r) and sum(b - a for (a, b), in r) == 4
    assert balanced_ranges([0, 2], 2, torch.tensor([0, 1, 2], dtype=torch.int64))[0][0] == (0, 0)
    assert balanced_ranges([3], 1, torch.tensor([0, 1, 2, 3], dtype=torch.int64)) == [[(0, 3)]]


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
        if isinstance(sizes, tuple):      # (sizes, "balanced"): edge-balanced (ragged) cuts of an unscattered graph
            sizes = list(sizes[0])
            n = sum(sizes)
            ranges = balanced_ranges(sizes, world, _adversarial_rowptr(*sizes))
        else:
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


@pytest.mark.parametrize("sizes", [[8, 6], [9, 5], ((40, 12), "balanced")])  # even -> all_gather; ragged -> broadcasts
def test_exchange_rows_gloo_world2(sizes):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, sizes, 3, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_edge_cost_prefix_weighs_streamed_rows_only_where_blocked_rows_exist():
    from deep_cbrs_amar_renaissance_b200.distributed import COLD_EDGE_COST, edge_cost_prefix
    lens = torch.tensor([3, 4, 5, 100, 2, 50, 1], dtype=torch.int64)      # users: 3 short rows; items: 100, 2, 50, 1
    rp = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(lens, 0)])
    assert edge_cost_prefix(rp, [3, 4], 0) is rp                          # row-major schedule: plain edge counts
    cp = edge_cost_prefix(rp, [3, 4], block_min_len=40)
    want = [3, 4, 5, 100, 2 * COLD_EDGE_COST, 50, 1 * COLD_EDGE_COST]    # users have no blocked row: unweighted
    assert torch.allclose(cp[1:] - cp[:-1], torch.tensor(want, dtype=torch.float64))
    r = balanced_ranges([3, 4], 2, cp)
    assert r[0][0][0] == 0 and r[1][0][1] == 3 and r[0][1][0] == 3 and r[1][1][1] == 7
