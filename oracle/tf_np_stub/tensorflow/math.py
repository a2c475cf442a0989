import numpy as np


def softmax(x, axis=-1):
    x = np.asarray(x, dtype=np.float32)
    e = np.exp(x - x.max(axis=axis, keepdims=True), dtype=np.float32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=np.float32)).astype(np.float32)


def count_nonzero(x):
    return int(np.count_nonzero(x))
