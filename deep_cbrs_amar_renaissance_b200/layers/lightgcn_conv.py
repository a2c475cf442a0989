"""Weight-free propagation X' = A_hat X (mirror of /root/reference/src/layers/lightgcn_conv.py:6-58)."""
import torch

from .. import ops
from ..graph import DeviceGraph
from ..keras_like import Layer


class LightGCNConv(Layer):
    def __init__(self, activity_regularizer=None, **kwargs):
        super().__init__("light_gcn_conv")

    def call(self, inputs, out=None, mask=None, csr=None, **kwargs):
        x, a = inputs
        csr = csr or a.norm
        if out is None:
            out = torch.empty(csr.n_rows, x.shape[1], dtype=torch.float32, device=x.device)
        return ops.spmm(csr, x, out)

    @staticmethod
    def preprocess(a):
        """gcn_filter(a): here the normalisation happens in the device graph build; the
        returned DeviceGraph serves its `norm` view (lightgcn_conv.py:56-58)."""
        return DeviceGraph.from_scipy(a)
