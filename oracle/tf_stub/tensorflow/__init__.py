"""Minimal stand-in for `tensorflow`, used ONLY by tests/golden/make_golden.py.

TensorFlow is not installable in this image.  The reference's numpy/scipy data
code (data.loaders, data.preprocess, data.datasets, utilities.math) only touches
`tf.float32` as a default argument and `tf.keras.utils.Sequence` as a base
class, so this stub lets those modules import and run UNMODIFIED from
/root/reference/src.  Nothing in the product imports this package.
"""
float32 = "float32"
from . import keras  # noqa: E402,F401
