"""DGCF layer (mirror of /root/reference/src/layers/dgcf_conv.py:11-102): out = M (x * sigmoid(w)) with the
DGCF operator M = gcn_filter(A) + high_pass(gcn_filter(A.A)) + I built on device (graph.DeviceGraph.dgcf) and
the per-node locality-adaptive gate w [N,1] (initialised to ones, l2-regularised)."""
import torch

from .. import ops
from ..graph import DeviceGraph
from ..keras_like import Layer


class LocalityAdaptive(Layer):
    def __init__(self, regularizer=None, **kwargs):
        super().__init__("locality_adaptive")
        self.regularizer = regularizer
        self.w = None

    def build(self, input_shape):
        self.w = self.add_weight("locality-adaptive-weights", (int(input_shape[0]), 1), "ones", self.regularizer)

    def call(self, inputs, **kwargs):
        return ops.row_gate(inputs, self.w)


class DGCFConv(Layer):
    def __init__(self, regularizer=None, **kwargs):
        super().__init__("dgcf_conv")
        self.regularizer = regularizer
        self.locality_adaptive = LocalityAdaptive(regularizer)

    def build(self, input_shape):
        if not self.locality_adaptive.built:
            self.locality_adaptive.build(input_shape[0])
            self.locality_adaptive.built = True

    def call(self, inputs, out=None, mask=None, csr=None, **kwargs):
        x, a = inputs
        csr = csr or a.dgcf
        gated = self.locality_adaptive(x)
        if out is None:
            out = torch.empty(csr.n_rows, x.shape[1], dtype=torch.float32, device=x.device)
        return ops.spmm(csr, gated, out)

    @staticmethod
    def preprocess(a):
        """The crosshop operator is built lazily on device by DeviceGraph.dgcf (dgcf_conv.py:38-49)."""
        return DeviceGraph.from_scipy(a)
