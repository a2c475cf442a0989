from . import basic, dense, gnn, hybrid  # noqa: F401
