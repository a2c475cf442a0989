import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_cbrs_amar_renaissance_b200.keras_like import set_seed
from deep_cbrs_amar_renaissance_b200.models import hybrid
from deep_cbrs_amar_renaissance_b200 import ops
from tests.helpers import random_bipartite
torch.cuda.set_device(0)
n_users, n_items = 6040, 3706
adj = random_bipartite(n_users, n_items, 572000, seed=42)
rng = np.random.RandomState(0)
u = rng.randint(0, n_users, size=1024); i = rng.randint(0, n_items, size=1024) + n_users; y = rng.randint(0, 2, size=1024)
bert = (rng.standard_normal((n_users + n_items, 768)) * 0.5).astype(np.float32)
set_seed(42)
model = hybrid.HybridBertGCN(adj, n_hiddens=[16, 16], n_layers=2, embedding_dim=16, dense_units=[[48, 48], [256, 64], [64, 64]], clf_units=[64, 64], feature_based=True, l2_regularizer=1e-4)
model.set_content_table(bert)
model.compile(loss="binary_crossentropy", optimizer={"learning_rate": 1e-3})
for s in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ops.LAUNCHES = 0
    loss, _ = model.train_on_batch((u, i), y)
    torch.cuda.synchronize(); print("step", s, (time.perf_counter() - t0) * 1e3, "ms", ops.LAUNCHES, "launches")
