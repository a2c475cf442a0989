import numpy as np


def modal_dot(a, x):
    """[3P] spektral.layers.ops.modal_dot in single mode with a sparse `a`: a @ x"""
    return np.asarray(a @ np.asarray(x, np.float32), dtype=np.float32)
