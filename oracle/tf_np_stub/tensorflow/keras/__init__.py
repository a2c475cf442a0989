from . import layers, models, regularizers, utils  # noqa: F401
