import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import train as ot
from tests.helpers import export_weights, random_bipartite
from tests.test_gpu_models import KINDS, _build, _oracle_graph, _randomise
from tests.test_gpu_training import _batch, _flat_views, _named_grads
from deep_cbrs_amar_renaissance_b200 import training
name = sys.argv[1] if len(sys.argv) > 1 else "BasicGraphSage"
n_users, n_items = 300, 200
adj = random_bipartite(n_users, n_items, 6000, seed=7)
model = _build(name, adj, (8, [8, 8], [24, 24], [48, 48]))
batches = [_batch(n_users, n_items, 512, 10 + s) for s in range(3)]
model((batches[0][0], batches[0][1]))
_randomise(model, seed=8)
model.compile(loss="binary_crossentropy", optimizer={"learning_rate": 1e-2})
kind = KINDS[name]
graph = _oracle_graph(kind, adj)
adam = training.Adam(learning_rate=1e-2)
state = {}
for t, (u, i, y) in enumerate(batches, start=1):
    w = export_weights(model)   # oracle restarts from the PRODUCT's weights each step: isolates per-step errors
    tape, loss, correct, _ = training.forward_backward(model, (u, i), y)
    got = _named_grads(model, tape)
    want, want_loss, _ = ot.gradients(kind, w, graph, (u, i), y, l2=0.0)
    for k in sorted(want):
        e = np.abs(got[k] - want[k]).max() / max(np.abs(want[k]).max(), 1e-30)
        if e > 1e-5:
            print("step", t, "grad", k, "rel err", e)
    l2 = training.l2_coefficients(model)
    ws = [w_ for w_ in model.trainable_weights if id(w_) in tape.wgrads]
    before = {n: w_.detach().cpu().numpy().copy() for n, w_ in model.named_weights()}
    gr = {n: tape.wgrads[id(w_)].detach().cpu().numpy().copy() for n, w_ in model.named_weights() if id(w_) in tape.wgrads}
    adam.apply(ws, [tape.wgrads[id(w_)] for w_ in ws], [l2.get(id(w_), 0.0) for w_ in ws])
    torch.cuda.synchronize()
    for n, w_ in model.named_weights():
        g = gr[n].astype(np.float64) + 2 * l2.get(id(w_), 0.0) * before[n]
        m, v = state.get(n, (np.zeros_like(g), np.zeros_like(g)))
        new, m, v = ot.adam_update(before[n].astype(np.float64), g, m, v, t, lr=1e-2)
        state[n] = (m, v)
        e = np.abs(w_.detach().cpu().numpy() - new).max()
        if e > 1e-6:
            print("step", t, "adam", n, "abs err", e)
print("done")
