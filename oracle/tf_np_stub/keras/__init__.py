# `from keras import models, layers, regularizers` (src/models/gnn.py:4, src/layers/dgcf_conv.py:1)
from tensorflow.keras import layers, models, regularizers  # noqa: F401
