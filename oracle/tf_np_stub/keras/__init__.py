from tensorflow.keras import layers  # noqa: F401  (`from keras import layers` in src/layers/dgcf_conv.py)
