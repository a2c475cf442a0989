"""Size-independent properties at BASELINE config 5's FULL size (10 M users x 1 M items x 1e9 edges, D = 128): the
oracle cannot replay this, so the kernels are checked through identities that hold at any size.  ~20 s on a B200."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_config5_full_size_properties():
    from deep_cbrs_amar_renaissance_b200 import _lib as L
    from deep_cbrs_amar_renaissance_b200 import ops
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    ops.check_device()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info()
    if free < 120 * (1 << 30):
        pytest.skip("needs ~120 GB of free HBM for the 2e9-entry sort")
    n_users, n_items, n_edges = 10_000_000, 1_000_000, 1_000_000_000
    n = n_users + n_items
    row, col = ops.synth_bipartite(n_users, n_items, n_edges, 42, dev)
    g = DeviceGraph(row, col, None, n)
    del row, col
    a = g.norm
    g.release_coo()
    torch.cuda.empty_cache()
    rowptr = a.rowptr
    # structure: offsets are monotone and end at nnz; nnz <= entries + self loops (duplicates were summed)
    assert int(rowptr[-1]) == a.nnz and bool((rowptr[1:] >= rowptr[:-1]).all())
    assert n < a.nnz <= 2 * n_edges + n
    # every row is strictly ascending (sampled: 4 blocks of 1e6 consecutive edges + the hottest item rows)
    for start in (0, a.nnz // 3, a.nnz // 2, a.nnz - 1_000_001):
        c = a.colidx[start:start + 1_000_000].long()
        r = torch.searchsorted(rowptr, torch.arange(start, start + 1_000_000, device=dev), right=True) - 1
        keys = r * n + c
        assert bool((keys[1:] > keys[:-1]).all())
    # user rows: the only column below U is the self loop, and it sorts first
    first = a.colidx[rowptr[:1000]]
    assert torch.equal(first.long(), torch.arange(1000, device=dev))
    out = torch.empty(n, 128, device=dev)
    ones = torch.ones(n, 128, device=dev)
    ops.spmm(g.norm, ones, out, agg=L.AGG_SUM)  # values ignored: row sums of the pattern = entries per row
    cnt = (rowptr[1:] - rowptr[:-1]).float()
    assert torch.equal(out[:, 0], cnt) and torch.equal(out[:, 0], out[:, 127])
    # A_hat = D^-1/2 (A+I) D^-1/2 has sqrt(deg) as a fixed vector, and (bipartite: A has no diagonal) A_hat_ii = 1/deg_i
    diag = a.vals[_diag_index(a, n, dev)].double()
    deg = 1.0 / diag
    assert bool((((deg.round() - deg).abs() / deg).max() < 1e-6) & (deg.min() > 1.5))   # integers >= 2 (fp32 A_hat_ii)
    sq = torch.sqrt(1.0 / diag).float()
    xs = sq.reshape(-1, 1).repeat(1, 4).contiguous()
    ys = torch.empty(n, 4, device=dev)
    ops.spmm(a, xs, ys)
    rel = ((ys[:, 0] - sq).abs() / sq).max().item()
    assert rel < 2e-5, rel   # A_hat sqrt(deg) == sqrt(deg), rows of up to ~1e6 terms in fp32
    # linearity and determinism on the full operator, D = 128
    x1 = torch.randn(n, 128, device=dev)
    y1 = torch.empty(n, 128, device=dev)
    y2 = torch.empty(n, 128, device=dev)
    ops.spmm(a, x1, y1)
    ops.spmm(a, x1, y2)
    assert torch.equal(y1, y2)
    x1.mul_(2.0)
    ops.spmm(a, x1, y2)
    assert torch.equal(y2, 2.0 * y1)  # scaling by a power of two is exact in every product and sum
    # a row slice (the multi-GPU unit of work) gives the same bits as the full launch
    lo, hi = n_users - 50_000, n_users + 50_000  # straddles the user / item boundary, includes the hottest items
    sl = a.row_slice(lo, hi)
    ops.spmm(sl, x1, out[:hi - lo])
    assert torch.equal(out[:hi - lo], y2[lo:hi])


def _diag_index(a, n, dev):
    """position of the (i,i) entry in every row: binary search per row over its ascending columns"""
    lo = a.rowptr[:-1].clone()
    hi = a.rowptr[1:].clone()
    target = torch.arange(n, device=dev, dtype=torch.int64)
    for _ in range(22):  # rows have < 2^21 entries
        mid = (lo + hi) // 2
        ok = mid < a.rowptr[1:]
        c = torch.where(ok, a.colidx[torch.clamp(mid, max=a.nnz - 1)].long(), torch.full_like(mid, n))
        go_right = c < target
        lo = torch.where(go_right & (lo < hi), mid + 1, lo)
        hi = torch.where(~go_right & (lo < hi), mid, hi)
    assert bool((a.colidx[lo].long() == target).all())
    return lo
