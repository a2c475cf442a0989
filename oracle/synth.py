"""Oracle (TEST INFRASTRUCTURE): numpy twin of cbrs_synth_bipartite (csrc/misc.cu), the
integer-only scaled-graph generator of SURVEY.md 8d config 5.  Bit-identical to the device."""
import math

import numpy as np


def _splitmix(x):
    x = x + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def synth_bipartite(n_users, n_items, n_edges, seed, scatter_items=True):
    """(row, col) int32 [2*n_edges]: (u, U+i) for every edge, then the transposed entries."""
    with np.errstate(over="ignore"):
        e = np.arange(n_edges, dtype=np.uint64)
        h0 = _splitmix(np.uint64(seed) ^ (e * np.uint64(0x2545F4914F6CDD1D)))
        h1 = _splitmix(h0)
        h2 = _splitmix(h1)
    c = max(n_items // 1024, 1)
    levels = 1
    while c * ((1 << levels) - 1) < n_items:
        levels += 1
    mult = 0x9E3779B1 % n_items or 1
    while math.gcd(mult, n_items) != 1:
        mult += 1
    if not scatter_items:
        mult = 1
    u = (h0 % np.uint64(n_users)).astype(np.int64)
    lvl = (h1 % np.uint64(levels)).astype(np.int64)
    span = (np.int64(c) << lvl).astype(np.uint64)
    rank = ((span - np.uint64(c) + (h2 % span)) % np.uint64(n_items)).astype(np.int64)
    it = (rank * mult) % n_items + n_users
    row = np.concatenate([u, it]).astype(np.int32)
    col = np.concatenate([it, u]).astype(np.int32)
    return row, col
