"""B200-native hot path of Deep_CBRS_Amar_Renaissance.

Sub-packages mirror the reference's src/ tree for the path that is in scope:
data/ (loaders, batch sequences, adjacency), layers/ (GNN layers, reduction,
fusion), models/ (GNN builders, BasicGNN / HybridBertGNN factories), utilities/.
The kernels live in csrc/ behind the C ABI of include/cbrs_b200.h.
"""
__version__ = "0.1.0"
