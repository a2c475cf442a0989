"""Reduction of the per-layer node representations (mirror of
/root/reference/src/layers/reduction.py:5-55).

'concatenation' costs nothing here: SequentialGNN lets every layer write into its
column slice of one [N, D_out] buffer and hands that buffer over, so `call`
only has to recognise the case."""
import torch

from .. import ops
from ..keras_like import Layer


class WeightedSum(Layer):
    """sum_l w_l^2 * h_l with learnable w initialised to ones (reduction.py:36-55)."""

    def __init__(self, regularizer=None, **kwargs):
        super().__init__("weighted_sum")
        self.regularizer = regularizer
        self.w = None

    def build(self, input_shape):
        self.w = self.add_weight("reduction-weights", (len(input_shape), 1, 1), "ones", self.regularizer)

    def call(self, inputs, **kwargs):
        w = self.w.reshape(-1)
        return ops.reduce_layers(list(inputs), coefs=(w * w).tolist())


class ReductionLayer(Layer):
    def __init__(self, method='concatenate', regularizer=None):
        super().__init__("reduction_layer")
        if method not in ('concatenation', 'sum', 'mean', 'w-sum', 'last'):
            raise ValueError('Reduction method not supported: ' + method)
        self.method = method
        self.layer = WeightedSum(regularizer) if method == 'w-sum' else None

    def call(self, inputs, out=None, **kwargs):
        hs = list(inputs)
        if self.method == 'concatenation':
            base = getattr(hs[0], "_cbrs_concat_buf", None)
            if base is not None and all(getattr(h, "_cbrs_concat_buf", None) is base for h in hs):
                return base  # the layers already wrote their slices
            return torch.cat(hs, dim=-1)
        if self.method == 'last':
            return hs[-1]
        if self.method == 'sum':
            return ops.reduce_layers(hs, out=out)
        if self.method == 'mean':
            return ops.reduce_layers(hs, divide_by=float(len(hs)), out=out)
        return self.layer(hs)
