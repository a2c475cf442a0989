"""debug: where does the peer-exchanged GAT differ from the single-GPU run?"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_cbrs_amar_renaissance_b200.distributed import RowPartition
from deep_cbrs_amar_renaissance_b200.keras_like import set_seed
from deep_cbrs_amar_renaissance_b200.models import basic
from tests.helpers import random_bipartite

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_users, n_items, n_props = 3001, 1999, 500
adj = random_bipartite(n_users, n_items, 150000, seed=21, n_props=n_props, n_links=4000, dup_links=300)
u = np.arange(512) % n_users
i = np.arange(512) % n_items + n_users
set_seed(7)
model = basic.BasicGAT(adj, n_hiddens=[16, 16], n_layers=2, embedding_dim=16, dense_units=[48, 48], clf_units=[64, 64])
seq = model.gnn.gnn_layers
model((u, i))
full = model.gnn(None).clone()
part = RowPartition([n_users, n_items, n_props], final_types=[0, 1, 2], exchange="peer").attach(seq)
got = model.gnn(None)
torch.cuda.synchronize()
part.heap.check()
bad = (got != full)
print(rank, "ranges", part.ranges, flush=True)
for c0 in (0, 16, 32):
    rows = bad[:, c0:c0 + 16].any(1).nonzero().flatten()
    print(rank, "cols", c0, "bad rows", rows.numel(), rows[:5].tolist(), rows[-5:].tolist() if rows.numel() else [], flush=True)
    if rows.numel():
        r = rows[0].item()
        print(rank, " got", got[r, c0:c0 + 4].tolist(), "want", full[r, c0:c0 + 4].tolist(), flush=True)
# z / q check vs single-GPU
for l in range(2):
    z = part._sym[("z", l)][1]
    q = part._sym[("q", l)][1]
    print(rank, "layer", l, "z finite", torch.isfinite(z).all().item(), "q zeros", (q == 0).sum().item(), flush=True)
dist.barrier()
dist.destroy_process_group()
