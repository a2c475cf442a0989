"""Multi-GPU parity (run on the GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Every rank builds the same graph and weights, runs the propagation once unpartitioned and once
row-partitioned (NCCL exchange), and requires the two to be BIT-identical for all four layer
families plus the relational extension; then checks user-sharded catalog top-k."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from deep_cbrs_amar_renaissance_b200.distributed import RowPartition  # noqa: E402
from deep_cbrs_amar_renaissance_b200.keras_like import set_seed  # noqa: E402
from deep_cbrs_amar_renaissance_b200.models import basic  # noqa: E402
from tests.helpers import random_bipartite  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_users, n_items, n_props = 3001, 1999, 500
    adj = random_bipartite(n_users, n_items, 150000, seed=21, n_props=n_props, n_links=4000, dup_links=300)
    u = np.arange(512) % n_users
    i = np.arange(512) % n_items + n_users
    for name in ("BasicGCN", "BasicGraphSage", "BasicGAT", "BasicLightGCN"):
        for even in (False, True):
            set_seed(7)
            sizes = [n_users, n_items, n_props]
            a = adj
            if even:  # a node count every type divides by the world size -> all_gather path
                sizes = [n_users + (-n_users) % world, n_items + (-n_items) % world, n_props + (-n_props) % world]
            model = getattr(basic, name)(a if not even else _pad(adj, n_users, n_items, n_props, sizes),
                                         n_hiddens=[16, 16], n_layers=2, embedding_dim=16, dense_units=[48, 48],
                                         clf_units=[64, 64])
            seq = model.gnn.gnn_layers
            model((u, i))
            full = model.gnn(None).clone()
            for exchange, pipeline in (("peer", "off"), ("peer", "kernel"), ("peer", "ce"), ("nccl", "off")):
                if pipeline != "off" and name != "BasicGCN":
                    continue  # the software pipeline exists for GCN stacks
                part = RowPartition(sizes, final_types=[0, 1, 2], exchange=exchange, pipeline=pipeline,
                                    row_blocks=3 if pipeline != "off" else 1).attach(seq)
                for rep in range(3):  # repeated calls reuse the symmetric buffers
                    got = model.gnn(None)
                    torch.cuda.synchronize()
                    assert torch.equal(got, full), "%s even=%s %s/%s rep %d: partitioned result differs" % (
                        name, even, exchange, pipeline, rep)
                if part.heap is not None:
                    part.heap.check()
                part.close()
                seq.partition = None
                # items only: user rows of other ranks are not exchanged, own rows and item rows must match
                part = RowPartition(sizes, final_types=[1], exchange=exchange).attach(seq)
                got = model.gnn(None)
                torch.cuda.synchronize()
                lo, hi = sizes[0], sizes[0] + sizes[1]
                assert torch.equal(got[lo:hi], full[lo:hi]), "%s %s: item rows differ" % (name, exchange)
                for a, b in part.mine:
                    assert torch.equal(got[a:b], full[a:b]), "%s %s: own rows differ" % (name, exchange)
                part.close()
                seq.partition = None
            if rank == 0:
                print("ok", name, "even" if even else "ragged", flush=True)
    # 128-wide GCN stack: the sparse kernel of layer l also produces layer l+1's transform and stores it into every
    # rank's copy (cbrs_spmm_gcn_fused); must still equal the single-GPU, unfused result bit for bit
    set_seed(11)
    model = basic.BasicGCN(adj, n_hiddens=[128, 128, 128], embedding_dim=128, dense_units=[48, 48], clf_units=[64, 64])
    seq = model.gnn.gnn_layers
    model((u, i))
    full = model.gnn(None).clone()
    for pipeline in ("fused", "off"):
        part = RowPartition([n_users, n_items, n_props], final_types=[0, 1, 2], exchange="peer", pipeline=pipeline).attach(seq)
        for rep in range(2):
            got = model.gnn(None)
            torch.cuda.synchronize()
            assert torch.equal(got, full), "128-wide GCN, pipeline=%s rep %d: partitioned result differs" % (pipeline, rep)
        part.heap.check()
        part.close()
        seq.partition = None
    if rank == 0:
        print("ok 128-wide GCN fused transform", flush=True)
    # user-sharded catalog top-k: each rank ranks its own users with replicated item rows
    set_seed(7)
    adj2 = random_bipartite(n_users, n_items, 150000, seed=22)
    model = basic.BasicGCN(adj2, n_hiddens=[16, 16], embedding_dim=16, dense_units=[48, 48], clf_units=[64, 64])
    model((u, i))
    model.cache_propagation = True
    ids_full, vals_full = model.recommend_top_k(n_users, n_items, 10)
    model.invalidate()
    part = RowPartition([n_users, n_items], final_types=[1]).attach(model.gnn.gnn_layers)
    a, b = part.ranges[rank][0]
    mine = torch.arange(a, b, device="cuda")
    ids, vals = model.recommend_top_k(n_users, n_items, 10, users=mine)
    assert torch.equal(ids, ids_full[a:b]) and torch.equal(vals, vals_full[a:b])
    if rank == 0:
        print("ok user-sharded top-k", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def _pad(adj, n_users, n_items, n_props, sizes):
    """Re-index nodes so each type has the padded size (extra nodes are isolated)."""
    from scipy import sparse
    def remap(x):
        x = x.astype(np.int64)
        out = x.copy()
        out[x >= n_users] += sizes[0] - n_users
        out[x >= n_users + n_items] += sizes[1] - n_items
        return out.astype(np.int32)
    n = sum(sizes)
    return sparse.coo_matrix((adj.data, (remap(adj.row), remap(adj.col))), shape=(n, n), dtype=np.float32)


if __name__ == "__main__":
    main()
