"""Keras Dense on the fused gather + concat + bias + activation kernel (cbrs_dense)."""
import torch

from .. import ops
from ..keras_like import Layer


class Dense(Layer):
    """layers.Dense(units, activation): act(x @ kernel[in, units] + bias)  (SURVEY A.6)."""

    def __init__(self, units, activation=None, name=None):
        super().__init__(name or "dense")
        self.units, self.activation = int(units), activation
        self.kernel = self.bias = None
        # "fp32": the FFMA kernel (1e-5 parity path).  "bf16": operands rounded to bf16, product on the tensor cores
        # (cbrs_dense_tc) for the shapes it takes, inference only - set through set_scorer_precision().
        self.precision = "fp32"

    def build(self, input_shape):
        self.build_for(int(input_shape[-1]))

    def build_for(self, in_features):
        if not self.built:
            self.kernel = self.add_weight("kernel", (in_features, self.units), "glorot_uniform")
            self.bias = self.add_weight("bias", (self.units,), "zeros")
            self.built = True

    def call(self, x, **kwargs):
        return self.call_sources([(x, None)])

    def call_sources(self, sources):
        """sources: one or two (matrix, row_index_or_None); concatenated along features
        and (when indexed) gathered inside the kernel, never materialised."""
        (x1, i1) = sources[0]
        x2, i2 = (sources[1] if len(sources) > 1 else (None, None))
        f1, f2 = x1.shape[1], (x2.shape[1] if x2 is not None else 0)
        self.build_for(f1 + f2)
        if x1.dtype == torch.bfloat16 or (x2 is not None and x2.dtype == torch.bfloat16):
            # bf16-STORED sources (a content table kept as bf16, set_content_table(dtype="bf16")): TMA-fed tensor-core
            # kernel; there is no fp32 kernel that reads bf16 rows, so the layer must be in bf16 precision
            if self.precision != "bf16" or not ops.dense_tc_bf16_eligible(f1, f2, self.units) or \
                    (x2 is not None and x2.dtype != x1.dtype):
                raise ValueError("a source stored as bf16 needs set_scorer_precision('bf16'), source widths that are "
                                 "multiples of 64 and at most 256 units (got {} + {} -> {}, precision {})".format(
                                     f1, f2, self.units, self.precision))
            return ops.dense_tc_bf16(x1, self.kernel, self.bias, self.activation, x2=x2, idx1=i1, idx2=i2)
        if self.precision == "bf16" and ops.dense_tc_eligible(f1, f2, self.units) and _aligned(x1) and _aligned(x2):
            return ops.dense_tc(x1, self.kernel, self.bias, self.activation, x2=x2, idx1=i1, idx2=i2)
        return ops.dense(x1, self.kernel, self.bias, self.activation, x2=x2, idx1=i1, idx2=i2)


def _aligned(x):
    """rows start on 16-byte boundaries (what the tensor-core kernel's 128-bit loads need)"""
    return x is None or (x.data_ptr() % 16 == 0 and (x.shape[0] <= 1 or x.stride(0) % 4 == 0))


def set_scorer_precision(layer, precision):
    """'fp32' | 'bf16' on every Dense below `layer` (a scorer: BasicRS / HybridCBRS).  bf16 = tensor-core towers and
    classifier for inference (BASELINE config 4); layers whose shapes the kernel does not take stay on the fp32 kernel."""
    if precision not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    if isinstance(layer, Dense):
        layer.precision = precision
    for _, sub in layer._sublayers():
        set_scorer_precision(sub, precision)


class DenseStack(Layer):
    """models.Sequential([Dense(u, ...) for u in units]) (src/models/dense.py:4-17)."""

    def __init__(self, units, activation=None, last_activation="same", last_units=None, name=None):
        super().__init__(name or "sequential")
        acts = [activation] * len(units)
        units = list(units)
        if last_units is not None:
            units.append(last_units)
            acts.append(activation if last_activation == "same" else last_activation)
        elif last_activation != "same" and units:
            acts[-1] = last_activation
        self.layers = [Dense(u, a, name="dense_%d" % k) for k, (u, a) in enumerate(zip(units, acts))]

    def build_for(self, in_features):
        """Create the weights for a given input width; returns the output width."""
        for layer in self.layers:
            layer.build_for(in_features)
            in_features = layer.units
        self.built = True
        return in_features

    def call(self, x, **kwargs):
        return self.call_sources([(x, None)])

    def call_sources(self, sources):
        if not self.layers:
            return materialize(sources)
        x = self.layers[0].call_sources(sources)
        for layer in self.layers[1:]:
            x = layer(x)
        return x


def materialize(sources):
    """Explicit gather + concat, only needed when an empty stack sits between two concats."""
    parts = [ops.gather_rows(x, i) if i is not None else x for x, i in sources]
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
