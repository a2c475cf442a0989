"""[3P] stand-ins for the three Spektral layers the reference builds (src/models/gnn.py:289-295,321-328,354-361):
constructor keywords and weight names as Spektral's, arithmetic = the ORACLE's restatement (oracle/layers.py).  Goldens
that pass through them pin the reference-owned wiring around the layers, not the layers' own arithmetic."""
import numpy as np
from tensorflow.keras.layers import RNG, Layer

from oracle import layers as ol

from . import ops  # noqa: F401
from ..utils import gcn_filter


class _Conv(Layer):
    def __init__(self, channels, activation=None, use_bias=True, kernel_regularizer=None, bias_regularizer=None, **kwargs):
        super().__init__()
        self.channels, self.activation, self.use_bias = int(channels), activation, use_bias

    @staticmethod
    def preprocess(a):
        return a

    def _bias(self):
        self.bias = self.add_weight(name='bias', shape=(self.channels,), initializer='zeros')
        self.bias[...] = RNG.uniform(-0.1, 0.1, size=self.channels).astype(np.float32)


class GCNConv(_Conv):
    def build(self, input_shape):
        self.kernel = self.add_weight(name='kernel', shape=(int(input_shape[0][-1]), self.channels))
        self._bias()

    def call(self, inputs, **kwargs):
        x, a = inputs
        return ol.gcn_conv(x, a.to_scipy(), self.kernel, self.bias, self.activation)

    @staticmethod
    def preprocess(a):
        return gcn_filter(a)


class GraphSageConv(_Conv):
    def __init__(self, channels, aggregate='mean', **kwargs):
        super().__init__(channels, **kwargs)
        self.aggregate = aggregate

    def build(self, input_shape):
        self.kernel = self.add_weight(name='kernel', shape=(2 * int(input_shape[0][-1]), self.channels))
        self._bias()

    def call(self, inputs, **kwargs):
        x, a = inputs
        indptr, indices, _ = a.csr_arrays()
        return ol.sage_conv(x, indptr, indices, self.kernel, self.bias, self.aggregate, self.activation)


class GATConv(_Conv):
    def __init__(self, channels, attn_heads=1, concat_heads=True, dropout_rate=0.5, add_self_loops=True, **kwargs):
        super().__init__(channels, **kwargs)
        assert attn_heads == 1 and add_self_loops and not dropout_rate

    def build(self, input_shape):
        f = int(input_shape[0][-1])
        self.kernel = self.add_weight(name='kernel', shape=(f, 1, self.channels))
        self.attn_kernel_self = self.add_weight(name='attn_kernel_self', shape=(self.channels, 1, 1))
        self.attn_kernel_neighs = self.add_weight(name='attn_kernel_neigh', shape=(self.channels, 1, 1))
        self._bias()

    def call(self, inputs, **kwargs):
        x, a = inputs
        indptr, indices, _ = a.csr_arrays()
        return ol.gat_conv(x, indptr, indices, self.kernel.reshape(self.kernel.shape[0], -1), self.attn_kernel_self.reshape(-1),
                           self.attn_kernel_neighs.reshape(-1), self.bias, self.activation)
