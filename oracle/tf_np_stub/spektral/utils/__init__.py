"""[3P] spektral.utils.gcn_filter is NOT available: this is the oracle's restatement of its published algorithm
(oracle.graph.gcn_filter_scipy), so goldens that pass through it pin the reference-owned code AROUND it only."""
from oracle.graph import gcn_filter_scipy as gcn_filter  # noqa: F401
