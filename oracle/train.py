"""Oracle (TEST INFRASTRUCTURE): one training step of the hot path, differentiated by
torch-CPU autograd in float64/float32, with the Keras Adam update written out.

Restates what Keras `fit` does per batch for the reference's models
(/root/reference/src/experiment.py:155-188): forward (SequentialGNN + scorer), loss =
binary cross-entropy on probabilities clipped to [1e-7, 1-1e-7] (Keras backend) + l2 * sum(w^2)
over the embeddings and GNN kernels/biases (src/models/gnn.py:239-246,293-294), Adam with
lr_t = lr*sqrt(1-b2^t)/(1-b1^t) and w -= lr_t*m/(sqrt(v)+eps) (Keras optimizer_v2; epsilon 1e-7).
PARITY UNPINNED for the same reason as oracle/layers.py: TensorFlow/Keras cannot run here.
The forward reuses the layer definitions of oracle/layers.py, re-expressed on torch tensors so
autograd can differentiate them; tests/test_gpu_training.py checks the product's explicit
backward kernels against these gradients.
"""
import numpy as np
import torch


def _t(a, dtype):
    return torch.tensor(np.asarray(a), dtype=dtype, requires_grad=True)


def _csr_torch(indptr, indices, data, n, dtype):
    rows = np.repeat(np.arange(n), np.diff(indptr))
    idx = torch.tensor(np.stack([rows, np.asarray(indices)]), dtype=torch.int64)
    return torch.sparse_coo_tensor(idx, torch.tensor(np.asarray(data), dtype=dtype), (n, n)).coalesce()


def _dense_stack(x, layers, last_sigmoid=False):
    for k, (w, b) in enumerate(layers):
        x = x @ w + b
        x = torch.sigmoid(x) if (last_sigmoid and k == len(layers) - 1) else torch.relu(x)
    return x


def _operator(kind, graph, n, dtype):
    """the propagation operator of one SequentialGNN as torch tensors"""
    if kind in ("gcn", "lightgcn", "dgcf"):
        return dict(a=_csr_torch(graph.indptr, graph.indices, graph.data, n, dtype))
    if kind == "rgcn":   # graph = one scipy CSR per relation (blocks of the normalised adjacency)
        return dict(a=[_csr_torch(g.indptr, g.indices, g.data, n, dtype) for g in graph])
    if kind == "gat":
        from .graph import gat_edges
        gptr, gcols = gat_edges(np.asarray(graph[0]), np.asarray(graph[1]))
        return dict(g_rows=torch.tensor(np.repeat(np.arange(n), np.diff(gptr)), dtype=torch.int64),
                    g_cols=torch.tensor(np.asarray(gcols), dtype=torch.int64))
    return dict(a=_csr_torch(graph[0], graph[1], np.ones(len(graph[1]), np.float32), n, dtype),
                deg=torch.tensor(np.diff(graph[0]).astype(np.float64), dtype=dtype).clamp(min=1.0).reshape(-1, 1))


def _propagate(kind, x, layers, graph, final_node, aggregate, leaves, reg, prefix, dtype):
    """SequentialGNN's loop + reduction (models/gnn.py:74-84) on torch tensors, starting from node features x.
    New leaves are registered as prefix + 'layers.<l>.<name>'; regularised tensors are appended to reg."""
    n = x.shape[0]
    op = _operator(kind, graph, n, dtype)
    a, deg, g_rows, g_cols = op.get("a"), op.get("deg"), op.get("g_rows"), op.get("g_cols")
    hs = [x]
    for li, lw in enumerate(layers):
        if kind == "gcn":
            k = leaves[prefix + "layers.%d.kernel" % li] = _t(lw["kernel"], dtype)
            b = leaves[prefix + "layers.%d.bias" % li] = _t(lw["bias"], dtype)
            reg += [k, b]
            x = torch.relu(torch.sparse.mm(a, x @ k) + b)
        elif kind == "rgcn":   # relational extension: relu(sum_r A_r (x W_r) + b), oracle.layers.rgcn_conv
            ks = []
            for r in range(len(a)):
                ks.append(_t(lw["kernel_%d" % r], dtype))
                leaves[prefix + "layers.%d.kernel_%d" % (li, r)] = ks[-1]
            b = leaves[prefix + "layers.%d.bias" % li] = _t(lw["bias"], dtype)
            reg += ks + [b]
            x = torch.relu(sum(torch.sparse.mm(a_r, x @ k_r) for a_r, k_r in zip(a, ks)) + b)
        elif kind == "lightgcn":
            x = torch.sparse.mm(a, x)
        elif kind == "dgcf":
            gw = leaves[prefix + "layers.%d.locality_adaptive/locality-adaptive-weights" % li] = _t(
                lw["locality_adaptive/locality-adaptive-weights"], dtype)
            reg.append(gw)
            x = torch.sparse.mm(a, x * torch.sigmoid(gw))
        elif kind == "sage":
            k = leaves[prefix + "layers.%d.kernel" % li] = _t(lw["kernel"], dtype)
            b = leaves[prefix + "layers.%d.bias" % li] = _t(lw["bias"], dtype)
            reg += [k, b]
            s = torch.sparse.mm(a, x)
            agg = s / deg if aggregate == "mean" else s
            o = torch.cat([x, agg], dim=1) @ k + b
            o = o / torch.sqrt(torch.clamp((o * o).sum(dim=1, keepdim=True), min=1e-12))
            x = torch.relu(o)
        elif kind == "gat":
            # spektral GATConv, single head, dropout 0; autograd also differentiates the max shift
            k = leaves[prefix + "layers.%d.kernel" % li] = _t(lw["kernel"], dtype)
            a_s = leaves[prefix + "layers.%d.attn_kernel_self" % li] = _t(lw["attn_self"], dtype)
            a_n = leaves[prefix + "layers.%d.attn_kernel_neigh" % li] = _t(lw["attn_neigh"], dtype)
            b = leaves[prefix + "layers.%d.bias" % li] = _t(lw["bias"], dtype)
            reg += [k, b]  # attention kernels carry no regulariser in the reference (gnn.py:321-328)
            z = x @ k
            p_, q_ = z @ a_s, z @ a_n
            e = torch.nn.functional.leaky_relu(p_[g_rows] + q_[g_cols], 0.2)
            mx = torch.full((n,), -float("inf"), dtype=dtype).scatter_reduce(0, g_rows, e, reduce="amax")
            wgt = torch.exp(e - mx[g_rows])
            ssum = torch.zeros(n, dtype=dtype).index_add(0, g_rows, wgt)
            alpha = wgt / (ssum[g_rows] + 1e-9)
            o = torch.zeros(n, z.shape[1], dtype=dtype).index_add(0, g_rows, alpha[:, None] * z[g_cols])
            x = torch.relu(o + b)
        else:
            raise ValueError(kind)
        hs.append(x)
    if kind in ("lightgcn", "dgcf"):
        final_node = "mean"
    if final_node == "concatenation":
        return torch.cat(hs, dim=1)
    if final_node == "mean":
        return sum(hs) / len(hs)
    if final_node == "sum":
        return sum(hs)
    return hs[-1]


def forward_loss(kind, w, graph, inputs, y, final_node="concatenation", aggregate="mean", l2=0.0, hybrid=False,
                 feature_based=False, dtype=torch.float64):
    """w: the export_weights() structure (numpy); graph: scipy CSR A_hat (gcn/lightgcn) or
    (indptr, indices) of the raw adjacency (sage).  Returns (loss, bce, probs, leaves) where
    leaves maps names to the torch leaf tensors whose .grad the caller reads after backward()."""
    leaves = {}
    emb = leaves["embeddings"] = _t(w["embeddings"], dtype)
    reg = [emb]
    red = _propagate(kind, emb, w["layers"], graph, final_node, aggregate, leaves, reg, "", dtype)
    return _score_and_loss(red, w, inputs, y, leaves, reg, l2, hybrid, feature_based, dtype)


def forward_loss_kg(kind, parts, w, inputs, y, n_users, n_items, final_node="concatenation", side_node="mean",
                    aggregate="mean", l2=0.0, hybrid=False, feature_based=False, dtype=torch.float64):
    """The Two-Step / Two-Way variants (models/tsgnn.py:92-94, models/twgnn.py:93-100) under autograd.
    parts: {'step_one' | 'way_one', 'way_two', 'step_two': dict(embeddings=, layers=, graph=)}; two-way when
    'way_one' is present.  side_node = item_node / user_item_node.  Leaves are named '<part>.embeddings',
    '<part>.layers.<l>.<name>'.  The embeddings of HalfInputSequentialGNN carry no regulariser
    (tsgnn.py:79-83 does not pass one); FullInputSequentialGNN has none at all."""
    leaves, reg = {}, []

    def side(part):
        emb = leaves[part + ".embeddings"] = _t(parts[part]["embeddings"], dtype)
        reg.append(emb)
        return _propagate(kind, emb, parts[part]["layers"], parts[part]["graph"], side_node, aggregate, leaves, reg,
                          part + ".", dtype)

    if "way_one" in parts:
        x = torch.cat([side("way_one")[:n_users], side("way_two")[:n_items]], dim=0)
    else:
        items = side("step_one")[:n_items]
        emb2 = leaves["step_two.embeddings"] = _t(parts["step_two"]["embeddings"], dtype)
        x = torch.cat([emb2, items], dim=0)
    red = _propagate(kind, x, parts["step_two"]["layers"], parts["step_two"]["graph"], final_node, aggregate, leaves,
                     reg, "step_two.", dtype)
    return _score_and_loss(red, w, inputs, y, leaves, reg, l2, hybrid, feature_based, dtype)


def _score_and_loss(red, w, inputs, y, leaves, reg, l2, hybrid, feature_based, dtype):
    """lookups + BasicRS / HybridCBRS + binary cross-entropy + l2 penalty on `reg`"""
    def stack(name):
        out = []
        for k, (kern, bias) in enumerate(w[name]):
            kk = leaves["%s.%d.kernel" % (name, k)] = _t(kern, dtype)
            bb = leaves["%s.%d.bias" % (name, k)] = _t(bias, dtype)
            out.append((kk, bb))
        return out

    u = torch.as_tensor(np.asarray(inputs[0]), dtype=torch.int64)
    i = torch.as_tensor(np.asarray(inputs[1]), dtype=torch.int64)
    if not hybrid:
        uu = _dense_stack(red[u], stack("unet"))
        ii = _dense_stack(red[i], stack("inet"))
        p = _dense_stack(torch.cat([uu, ii], dim=1), stack("clf"), last_sigmoid=True)
    else:
        ub = torch.tensor(np.asarray(inputs[2]), dtype=dtype)
        ib = torch.tensor(np.asarray(inputs[3]), dtype=dtype)
        ug = _dense_stack(red[u], stack("dense1a"))
        ig = _dense_stack(red[i], stack("dense1b"))
        ub = _dense_stack(ub, stack("dense2a"))
        ib = _dense_stack(ib, stack("dense2b"))
        def fuse(name, a, b):
            fw = w.get(name)
            if fw is None:
                return torch.cat([a, b], dim=1)
            if fw.get("proj_weight") is not None:
                pw = leaves["%s.proj_weight" % name] = _t(fw["proj_weight"], dtype)
                if fw["proj_first"]:
                    a = a @ pw
                else:
                    b = b @ pw
            aw = leaves["%s.att_weight" % name] = _t(fw["att_weight"], dtype)
            x = torch.stack([a, b], dim=1)
            att = torch.softmax(torch.tanh(x @ aw), dim=1)
            return (att * x).sum(dim=1)

        if feature_based:
            x1 = _dense_stack(fuse("fuse1a", ug, ig), stack("dense3a"))
            x2 = _dense_stack(fuse("fuse1b", ub, ib), stack("dense3b"))
        else:
            x1 = _dense_stack(fuse("fuse1a", ug, ub), stack("dense3a"))
            x2 = _dense_stack(fuse("fuse1b", ig, ib), stack("dense3b"))
        x = fuse("fuse2", x1, x2)
        if w.get("residual") is None:
            p = _dense_stack(x, stack("clf"), last_sigmoid=True)
        else:
            rl = stack("residual")
            r = _dense_stack(x, rl[:-1])
            r = r @ rl[-1][0] + rl[-1][1]
            p = _dense_stack(torch.relu(r + x1 + x2), stack("clf"), last_sigmoid=True)
    p = p.reshape(-1)
    yt = torch.tensor(np.asarray(y, dtype=np.float64), dtype=dtype)
    pc = torch.clamp(p, 1e-7, 1 - 1e-7)
    bce = -(yt * torch.log(pc) + (1 - yt) * torch.log(1 - pc)).mean()
    loss = bce
    if l2:
        loss = loss + l2 * sum((t * t).sum() for t in reg)
    return loss, bce, p, leaves


def gradients_kg(kind, parts, w, inputs, y, n_users, n_items, **kw):
    """gradients() for the Two-Step / Two-Way variants"""
    return _grads(*forward_loss_kg(kind, parts, w, inputs, y, n_users, n_items, **kw))


def gradients(kind, w, graph, inputs, y, **kw):
    """{leaf name: dLoss/dleaf (numpy float64)} plus the loss value and the probabilities."""
    return _grads(*forward_loss(kind, w, graph, inputs, y, **kw))


def _grads(loss, bce, p, leaves):
    loss.backward()
    grads = {k: (v.grad.detach().numpy().astype(np.float64) if v.grad is not None else np.zeros(tuple(v.shape)))
             for k, v in leaves.items()}
    return grads, float(loss.detach()), p.detach().numpy()


def adam_update(w, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """One Keras Adam step on numpy arrays (float64 math); returns (w, m, v)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return w - lr_t * m / (np.sqrt(v) + eps), m, v
