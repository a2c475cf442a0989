"""Oracle (TEST INFRASTRUCTURE): multi-threaded torch-CPU twin of the numpy oracle, used only
as the timed CPU baseline in bench.py (cpu_baseline / --impl reference).  The reference's own
CPU path is TensorFlow's multi-threaded kernels, which cannot be installed here; torch's CSR
sparse-dense product and addmm use all host threads the same way.  Checked against the numpy
oracle in tests/test_oracle_torch_cpu.py."""
import numpy as np
import torch
from scipy import sparse

from . import graph as og


def glorot(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def gcn_filter_torch(row, col, n):
    """scipy graph build (the reference's own code path for this step) -> torch CSR."""
    adj = sparse.coo_matrix((np.ones(len(row), np.float32), (row, col)), shape=(n, n))
    a = og.gcn_filter(adj)
    import warnings
    warnings.filterwarnings("ignore", message="Sparse")
    t = torch.sparse_csr_tensor(torch.from_numpy(a.indptr.astype(np.int64)), torch.from_numpy(a.indices.astype(np.int64)),
                                torch.from_numpy(a.data), size=a.shape)
    return t, a.nnz


def gcn_forward(emb, a_hat, layers):
    """SequentialGNN.call with GCNConv layers + 'concatenation' (models/gnn.py:74-84)."""
    x = emb
    hs = [x]
    for w, b in layers:
        x = torch.relu(torch.sparse.mm(a_hat, x @ w) + b)
        hs.append(x)
    return torch.cat(hs, dim=1)


def random_basic_rs(rng, d_in, dense_units, clf_units):
    def stack(d, units):
        out = []
        for u in units:
            out.append((torch.from_numpy(glorot(rng, (d, u))), torch.zeros(u)))
            d = u
        return out, d
    unet, du = stack(d_in, dense_units)
    inet, di = stack(d_in, dense_units)
    clf, _ = stack(du + di, list(clf_units) + [1])
    return dict(unet=unet, inet=inet, clf=clf)


def basic_rs(emb, u, i, w):
    def net(x, layers, last_sigmoid=False):
        for k, (kern, b) in enumerate(layers):
            x = x @ kern + b
            x = torch.sigmoid(x) if (last_sigmoid and k == len(layers) - 1) else torch.relu(x)
        return x
    uu = net(emb[u], w["unet"])
    ii = net(emb[i], w["inet"])
    return net(torch.cat([uu, ii], dim=1), w["clf"], last_sigmoid=True)
