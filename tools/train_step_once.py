"""one eager optimiser step at MovieLens-1M shape (for `ncu --metrics gpu__time_duration.sum` launch lists)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_cbrs_amar_renaissance_b200.keras_like import set_seed
from deep_cbrs_amar_renaissance_b200.models import basic
from tests.helpers import random_bipartite
name = sys.argv[1] if len(sys.argv) > 1 else "BasicGCN"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.cuda.set_device(0)
n_users, n_items = 6040, 3706
adj = random_bipartite(n_users, n_items, 572000, seed=42)
rng = np.random.RandomState(0)
u = rng.randint(0, n_users, size=1024); i = rng.randint(0, n_items, size=1024) + n_users; y = rng.randint(0, 2, size=1024)
set_seed(42)
model = getattr(basic, name)(adj, n_hiddens=[16, 16], n_layers=2, embedding_dim=16, dense_units=[48, 48], clf_units=[64, 64], l2_regularizer=1e-4)
model.compile(loss="binary_crossentropy", optimizer={"learning_rate": 1e-3})
for _ in range(steps):
    loss, _ = model.train_on_batch((u, i), y)
torch.cuda.synchronize()
print("loss", float(loss.item()))
