"""numpy float32 stand-in for the handful of TensorFlow ops the reference's own layer code calls (see ../README.md)."""
import numpy as np

from . import keras, math, nn, sparse  # noqa: F401
from .sparse import SparseTensor  # noqa: F401

float32 = np.float32


def _f(x):
    return np.asarray(x, dtype=np.float32)


def concat(values, axis):
    return np.concatenate([_f(v) for v in values], axis=axis)


def add_n(inputs):
    acc = _f(inputs[0])
    for x in inputs[1:]:   # tf.add_n adds in list order
        acc = acc + _f(x)
    return acc


def divide(a, b):
    return (_f(a) / np.float32(b)).astype(np.float32)


def stack(values, axis=0):
    return np.stack([_f(v) for v in values], axis=axis)


def matmul(a, b):
    return np.matmul(_f(a), _f(b)).astype(np.float32)


def tanh(x):
    return np.tanh(_f(x), dtype=np.float32)


def sigmoid(x):
    return (np.float32(1) / (np.float32(1) + np.exp(-_f(x), dtype=np.float32))).astype(np.float32)


def multiply(a, b):
    return (_f(a) * _f(b)).astype(np.float32)


def reduce_sum(x, axis=None):
    return _f(x).sum(axis=axis, dtype=np.float32)


def argmin(values):
    return int(np.argmin(np.asarray(values, dtype=np.float64)))   # first minimum, like tf.argmin


def cast(x, dtype):
    if isinstance(x, SparseTensor):
        x.values = x.values.astype(np.float32)
        return x
    return np.asarray(x).astype(dtype)


def convert_to_tensor(x, dtype=None):
    return np.asarray(x, dtype=dtype)


def slice(input_, begin, size):  # noqa: A001  (tf.slice; -1 in `size` = to the end of that axis)
    x = np.asarray(input_)
    idx = tuple(np.s_[b:(None if s == -1 else b + s)] for b, s in zip(begin, size))
    return x[idx]
