from .spektral_conv import GCNConv, GraphSageConv, GATConv, RGCNConv  # noqa: F401
from .lightgcn_conv import LightGCNConv  # noqa: F401
from .dgcf_conv import DGCFConv, LocalityAdaptive  # noqa: F401
from .reduction import ReductionLayer, WeightedSum  # noqa: F401
from .fusion import FusionLayer  # noqa: F401
from .dense import Dense, DenseStack  # noqa: F401
