from . import layers, utils  # noqa: F401
