"""Training step of the Two-Step / Two-Way variants (scope row (f)-4 on top of (f)-1): the tape threads the
gradient from the scorer through the second GNN, the row slices and the stacked input back into the first GNN(s).
Gradients vs the float64 autograd oracle (oracle/train.py::gradients_kg, itself checked against finite differences in
tests/test_oracle_train_cpu.py); tolerances as in tests/test_gpu_training.py.
(File name sorts last on purpose: added after the round's last GPU session.)"""
import numpy as np
import pytest
import torch

from oracle import train as ot
from tests.helpers import assert_close, kg_graphs, kg_model, weights_struct
from tests.test_gpu_models import _oracle_graph, _randomise
from tests.test_gpu_training import _batch, assert_grad_close

pytestmark = pytest.mark.gpu

N_USERS, N_ITEMS, N_PROPS = 300, 200, 120
PARTS = {"step_one_gnn_layers": "step_one", "way_one_gnn_layers": "way_one", "way_two_gnn_layers": "way_two",
         "step_two_gnn_layers": "step_two"}
KIND = {"GCN": "gcn", "GAT": "gat", "GraphSage": "sage", "LightGCN": "lightgcn", "DGCF": "dgcf"}


@pytest.fixture(scope="module", autouse=True)
def _device():
    from deep_cbrs_amar_renaissance_b200 import ops
    ops.check_device()
    torch.cuda.set_device(0)


@pytest.fixture(scope="module")
def graphs():
    from deep_cbrs_amar_renaissance_b200.data.preprocess import get_user_properties
    ui, ip = kg_graphs(N_USERS, N_ITEMS, N_PROPS, n_pos=6000, n_links=700, dup_links=30)
    return ui, ip, get_user_properties(ui, ip, N_USERS, N_ITEMS)


def export_parts(model, kind, graphs):
    """numpy weights of every SequentialGNN of the model in the oracle's structures + the scorer stacks"""
    ui, ip, up = graphs
    adj_of = {"step_one": ip, "way_one": up, "way_two": ip, "step_two": ui}
    named = {n: w.detach().cpu().numpy() for n, w in model.named_weights()}
    parts = {}
    for attr, part in PARTS.items():
        seq = getattr(model.gnn, attr, None)
        if seq is None:
            continue
        pre = "gnn/%s/" % attr
        sub = {"gnn/gnn_layers/" + n[len(pre):]: v for n, v in named.items() if n.startswith(pre)}
        emb = sub.get("gnn/gnn_layers/embeddings")
        sub.setdefault("gnn/gnn_layers/embeddings", np.zeros((1, 1), np.float32))
        layers = weights_struct(sub)["layers"]
        layers += [{} for _ in range(len(seq.seq_layers) - len(layers))]
        parts[part] = dict(embeddings=emb, layers=layers, graph=_oracle_graph(kind, adj_of[part]))
    rs = {n: v for n, v in named.items() if n.startswith("rs/")}
    rs["gnn/gnn_layers/embeddings"] = np.zeros((1, 1), np.float32)
    return parts, weights_struct(rs)


def named_grads(model, tape):
    """oracle leaf name -> product gradient"""
    out = {}
    for name, w in model.named_weights():
        g = tape.wgrads.get(id(w))
        if g is None:
            continue
        g = g.detach().cpu().numpy()
        if name.startswith("gnn/"):
            _, attr, rest = name.split("/", 2)
            if rest == "embeddings":
                out[PARTS[attr] + ".embeddings"] = g
            else:
                k, leaf = rest[len("seq_layers."):].split("/", 1)
                g = g.reshape(g.shape[0], -1) if leaf == "kernel" else g.reshape(-1) if leaf.startswith("attn") else g
                out["%s.layers.%s.%s" % (PARTS[attr], k, leaf)] = g
        else:
            stack, layer, leaf = name[3:].split("/")
            out["%s.%s.%s" % (stack, layer.split(".")[1], leaf)] = g
    return out


CASES = ["BasicTSGCN", "BasicTSGraphSage", "BasicTSGAT", "BasicTSLightGCN", "BasicTSDGCF", "BasicTSGCN-itemconcat",
         "BasicTSGraphSage-mean", "BasicTWGCN", "BasicTWGraphSage", "BasicTWGAT", "BasicTWLightGCN", "BasicTWGCN-uiconcat"]


@pytest.mark.parametrize("case", CASES)
def test_kg_gradients_match_autograd_oracle(case, graphs):
    from deep_cbrs_amar_renaissance_b200 import training
    from deep_cbrs_amar_renaissance_b200.keras_like import set_seed
    set_seed(42)
    model, kw = kg_model(case, graphs, N_USERS, N_ITEMS)
    name = case.split("-")[0]
    kind = KIND[name[len("BasicTS"):]]
    u, i, y = _batch(N_USERS, N_ITEMS, 512, 3)
    model((u, i))
    _randomise(model, seed=5)
    parts, w = export_parts(model, kind, graphs)
    tape, loss, correct, probs = training.forward_backward(model, (u, i), y)
    torch.cuda.synchronize()
    side = kw.get("item_node", kw.get("user_item_node", "mean"))
    want, want_loss, want_p = ot.gradients_kg(kind, parts, w, (u, i), y, N_USERS, N_ITEMS, final_node=kw["final_node"],
                                              side_node=side, aggregate=kw["aggregate"])
    assert_close(probs.cpu().numpy().reshape(-1), want_p, rtol=2e-5, what=case + " probabilities")
    assert abs(float(loss.item()) - want_loss) <= 1e-5 * max(1.0, abs(want_loss))
    got = named_grads(model, tape)
    assert set(got) == set(want), (sorted(set(got) ^ set(want)))
    for k in sorted(want):
        assert_grad_close(got[k], want[k], "%s grad %s" % (case, k))


@pytest.mark.parametrize("case", ["BasicTSGCN", "BasicTWGraphSage"])
def test_kg_train_steps_reduce_the_loss_and_replay_matches_eager(case, graphs):
    """a few Adam steps on one batch lower the loss; the CUDA-graph replay of the step equals the eager step"""
    from deep_cbrs_amar_renaissance_b200 import training
    from deep_cbrs_amar_renaissance_b200.keras_like import set_seed
    u, i, y = _batch(N_USERS, N_ITEMS, 512, 6)

    def run(graphed):
        set_seed(42)
        model, _ = kg_model(case, graphs, N_USERS, N_ITEMS)
        model.build_weights()
        adam = training.Adam(learning_rate=1e-2)
        step = training.GraphedTrainStep(model, adam, 512) if graphed else None
        losses = []
        for _ in range(4):
            loss, _ = step((u, i), y) if graphed else training.train_step(model, adam, (u, i), y)
            losses.append(float(loss.item()))
        return losses, [w.detach().cpu().numpy() for w in model.weights]

    eager, w_eager = run(False)
    assert eager[-1] < eager[0]
    replay, w_replay = run(True)
    assert eager == replay
    for a, b in zip(w_eager, w_replay):
        assert np.array_equal(a, b)


def test_rgcn_gradients():
    """training step of the relational extension (row R): one product with the transposed stacked operator"""
    from deep_cbrs_amar_renaissance_b200 import training
    from deep_cbrs_amar_renaissance_b200.graph import DeviceGraph
    from oracle import graph as og
    from tests.helpers import export_weights, random_bipartite, relation_blocks
    from tests.test_gpu_models import GRIDS, _build
    from tests.test_gpu_training import _named_grads
    n_users, n_items, n_props = 100, 80, 50
    adj = random_bipartite(n_users, n_items, 2000, seed=5, n_props=n_props, n_links=300, dup_links=30)
    dev = torch.device("cuda", 0)
    row, col = torch.from_numpy(adj.row).to(dev), torch.from_numpy(adj.col).to(dev)
    rel_np = ((adj.row >= n_users + n_items) | (adj.col >= n_users + n_items)).astype(np.int32)
    g2 = DeviceGraph(row, col, torch.from_numpy(adj.data).to(dev), adj.shape[0], rel=torch.from_numpy(rel_np).to(dev), n_rel=2)
    model = _build("BasicRGCN", g2, GRIDS[1])
    u, i, y = _batch(n_users, n_items, 256, 3)
    model((u, i))
    _randomise(model, seed=5)
    named = {n: w.detach().cpu().numpy() for n, w in model.named_weights()}
    w = export_weights(model)
    for l, lw in enumerate(w["layers"]):   # export_weights knows 'kernel'; the relational layer names them kernel_<r>
        for r in range(2):
            lw["kernel_%d" % r] = named["gnn/gnn_layers/seq_layers.%d/kernel_%d" % (l, r)]
    tape, loss, _, probs = training.forward_backward(model, (u, i), y)
    want, want_loss, want_p = ot.gradients("rgcn", w, relation_blocks(og.gcn_filter(adj), n_users + n_items), (u, i), y)
    assert_close(probs.cpu().numpy().reshape(-1), want_p, rtol=2e-5, what="rgcn probabilities")
    got = _named_grads(model, tape)
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    for k in sorted(want):
        assert_grad_close(got[k], want[k], "rgcn grad %s" % k)
