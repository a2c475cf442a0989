"""Fusion of two [B, F] feature blocks (mirror of /root/reference/src/layers/fusion.py:5-68).

'concatenate' is never executed as a copy on the hot path: HybridCBRS feeds the two blocks to the next Dense as a
two-source input.  'attention' (the hybrid tweaks grids): both blocks - the narrower one first projected to the wider
width by `proj_weight` - go through the shared `att_weight` and tanh, a softmax over the TWO sources per feature weighs
them, and the weighted sum is the output: two cbrs_dense calls + cbrs_attn_fuse."""
import torch

from .. import ops
from ..keras_like import Layer


class FusionLayer(Layer):
    def __init__(self, method='concatenate'):
        super().__init__("fusion_layer")
        if method not in ['concatenate', 'attention']:
            raise ValueError("Unknown concatenation method called {}".format(method))
        self.method = method
        self.proj_first = None
        self.proj_weight = self.att_weight = None

    def build_for(self, fa, fb):
        """Create the attention weights for input widths (fa, fb); returns the output width."""
        if self.method == 'concatenate':
            self.built = True
            return fa + fb
        if not self.built:
            f = max(fa, fb)
            if fa != fb:
                self.proj_first = fa < fb   # the narrower block is projected up (fusion.py:24-32)
                self.proj_weight = self.add_weight("proj_weight", (min(fa, fb), f), "glorot_uniform")
            self.att_weight = self.add_weight("att_weight", (f, f), "glorot_uniform")
            self.built = True
        return max(fa, fb)

    def build(self, input_shape):
        self.build_for(int(input_shape[0][-1]), int(input_shape[1][-1]))

    def call(self, inputs, *args, **kwargs):
        a, b = inputs
        if self.method == 'concatenate':
            return torch.cat([a, b], dim=1)
        self.build_for(a.shape[1], b.shape[1])
        if self.proj_first is not None:
            if self.proj_first:
                a = ops.dense(a, self.proj_weight)
            else:
                b = ops.dense(b, self.proj_weight)
        ta = ops.dense(a, self.att_weight, None, "tanh")
        tb = ops.dense(b, self.att_weight, None, "tanh")
        return ops.attn_fuse(a, b, ta, tb)
