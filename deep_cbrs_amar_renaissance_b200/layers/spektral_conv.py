"""GCNConv, GraphSageConv, GATConv with the constructor / call convention of the
Spektral layers the reference instantiates at /root/reference/src/models/gnn.py:289-295,
321-328,354-361 (`layer([x, a]) -> x'`, static `preprocess`).  Spektral itself is a
third-party dependency of the reference (unpinned, requirements.txt:8); the arithmetic
follows its published algorithm as restated in SURVEY.md Appendix A.  RGCNConv is the
relational extension (row R; no reference counterpart).
"""
import torch

from .. import _lib as L
from .. import ops
from ..graph import DeviceGraph
from ..keras_like import Layer


class _Conv(Layer):
    def __init__(self, channels, activation=None, use_bias=True, kernel_regularizer=None, bias_regularizer=None,
                 name=None, **kwargs):
        super().__init__(name)
        self.channels, self.activation, self.use_bias = int(channels), activation, use_bias
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        if activation not in (None, "linear", "relu"):
            raise NotImplementedError("fused epilogues cover relu / linear (all the reference's grids use relu)")
        # storage type of the operand the sparse kernel gathers ("fp32" = the reference's arithmetic; "bf16" halves the
        # bytes per edge: products and sums stay fp32, only the stored transform is rounded; SequentialGNN.set_feature_dtype)
        self.feature_dtype = "fp32"

    def _out(self, out, n_rows, device):
        if out is None:
            out = torch.empty(n_rows, self.channels, dtype=torch.float32, device=device)
        return out

    @staticmethod
    def preprocess(a):
        return a


class GCNConv(_Conv):
    """relu(A_hat @ (x @ kernel) + bias): transform first, then propagate."""

    def __init__(self, channels, activation=None, **kwargs):
        super().__init__(channels, activation, name="gcn_conv", **kwargs)

    def build(self, input_shape):
        f = int(input_shape[0][-1])
        self.kernel = self.add_weight("kernel", (f, self.channels), "glorot_uniform", self.kernel_regularizer)
        self.bias = self.add_weight("bias", (self.channels,), "zeros", self.bias_regularizer) if self.use_bias else None

    def call(self, inputs, out=None, csr=None, **kwargs):
        x, a = inputs
        csr = csr or a.norm
        z = ops.gcn_transform(x, self.kernel, x.shape[0],
                              out_dtype=torch.bfloat16 if self.feature_dtype == "bf16" else None)
        return ops.spmm(csr, z, self._out(out, csr.n_rows, x.device), bias=self.bias,
                        relu=self.activation == "relu")

    @staticmethod
    def preprocess(a):
        """GCNConv.preprocess == gcn_filter (gnn.py:283): served by DeviceGraph.norm."""
        return DeviceGraph.from_scipy(a)


class GraphSageConv(_Conv):
    """relu(l2_normalize([x || agg(x)] @ kernel + bias)), agg over the raw edge list."""

    def __init__(self, channels, aggregate="mean", activation=None, **kwargs):
        super().__init__(channels, activation, name="graph_sage_conv", **kwargs)
        if aggregate not in ("mean", "sum"):
            raise NotImplementedError("aggregate '{}' (only 'mean' appears in the reference's configs; "
                                      "'sum' is also built)".format(aggregate))
        self.aggregate = aggregate

    def build(self, input_shape):
        f = int(input_shape[0][-1])
        self.kernel = self.add_weight("kernel", (2 * f, self.channels), "glorot_uniform", self.kernel_regularizer)
        self.bias = self.add_weight("bias", (self.channels,), "zeros", self.bias_regularizer) if self.use_bias else None

    def call(self, inputs, out=None, csr=None, **kwargs):
        x, a = inputs
        csr = csr or a.raw
        agg = torch.empty(csr.n_rows, x.shape[1], dtype=torch.float32, device=x.device)
        ops.spmm(csr, x, agg, agg=L.AGG_MEAN if self.aggregate == "mean" else L.AGG_SUM)
        x_self = x[csr.row_offset:csr.row_offset + csr.n_rows]
        return ops.sage_dense(x_self, agg, self.kernel, self.bias, self.activation, x.shape[0],
                              out=self._out(out, csr.n_rows, x.device))


class GATConv(_Conv):
    """Single-head GAT with self loops, LeakyReLU(0.2) scores, +1e-9 softmax."""

    def __init__(self, channels, attn_heads=1, concat_heads=True, dropout_rate=0.5, add_self_loops=True,
                 activation=None, **kwargs):
        super().__init__(channels, activation, name="gat_conv", **kwargs)
        if attn_heads != 1:
            raise NotImplementedError("the reference builds GATConv with the default single head (gnn.py:321-328)")
        if dropout_rate not in (0, 0.0, None):
            raise NotImplementedError("attention dropout is 0.0 in every reference config (config.yaml:20)")
        if not add_self_loops:
            raise NotImplementedError("the reference keeps add_self_loops=True")
        self.dropout_rate = dropout_rate

    def build(self, input_shape):
        f = int(input_shape[0][-1])
        self.kernel = self.add_weight("kernel", (f, 1, self.channels), "glorot_uniform", self.kernel_regularizer)
        self.attn_kernel_self = self.add_weight("attn_kernel_self", (self.channels, 1, 1), "glorot_uniform")
        self.attn_kernel_neighs = self.add_weight("attn_kernel_neigh", (self.channels, 1, 1), "glorot_uniform")
        self.bias = self.add_weight("bias", (self.channels,), "zeros", self.bias_regularizer) if self.use_bias else None

    def call(self, inputs, out=None, csr=None, **kwargs):
        x, a = inputs
        csr = csr or a.raw
        z, p, q = ops.gat_transform(x, self.kernel.reshape(x.shape[1], self.channels), self.attn_kernel_self.reshape(-1),
                                    self.attn_kernel_neighs.reshape(-1), x.shape[0])
        return ops.gat(csr, z, p, q, self._out(out, csr.n_rows, x.device), bias=self.bias,
                       relu=self.activation == "relu", row_offset=csr.row_offset)


class RGCNConv(_Conv):
    """Relational GCN (extension): relu(sum_r A_hat_r @ (x @ W_r) + b).

    `a` is a DeviceGraph built with relation ids; its norm view carries column ids
    r*N + j, so ONE SpMM over the stacked transforms [R*N, H] is the grouped scatter.
    With n_rel == 1 this is GCNConv exactly."""

    def __init__(self, channels, n_rel, activation=None, **kwargs):
        super().__init__(channels, activation, name="rgcn_conv", **kwargs)
        self.n_rel = int(n_rel)

    def build(self, input_shape):
        f = int(input_shape[0][-1])
        self.kernels = [self.add_weight("kernel_%d" % r, (f, self.channels), "glorot_uniform", self.kernel_regularizer)
                        for r in range(self.n_rel)]
        self.bias = self.add_weight("bias", (self.channels,), "zeros", self.bias_regularizer) if self.use_bias else None

    def call(self, inputs, out=None, csr=None, **kwargs):
        x, a = inputs
        csr = csr or a.norm
        n = x.shape[0]
        z = torch.empty(self.n_rel * n, self.channels, dtype=torch.float32, device=x.device)
        ops.dense_grouped(x, self.kernels, z, n)   # all relations' transforms, one launch, straight into the stack
        return ops.spmm(csr, z, self._out(out, csr.n_rows, x.device), bias=self.bias,
                        relu=self.activation == "relu")
