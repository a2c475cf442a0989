"""Fusion of two [B, F] feature blocks (mirror of /root/reference/src/layers/fusion.py:5-68).

'concatenate' is never executed as a copy on the hot path: HybridCBRS feeds the two
blocks to the next Dense as a two-source input.  The learned 'attention' fusion is
used only by the tweaks grid ("next") and raises."""
import torch

from ..keras_like import Layer


class FusionLayer(Layer):
    def __init__(self, method='concatenate'):
        super().__init__("fusion_layer")
        if method not in ['concatenate', 'attention']:
            raise ValueError("Unknown concatenation method called {}".format(method))
        if method == 'attention':
            raise NotImplementedError("attention fusion is outside the first hot-path bar (DESIGN.md)")
        self.method = method

    def call(self, inputs, *args, **kwargs):
        a, b = inputs
        return torch.cat([a, b], dim=1)
