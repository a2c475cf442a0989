"""Host-side sparse helpers (mirror of /root/reference/src/utilities/math.py:6-56)."""
import numpy as np
from scipy import sparse


def symmetrize_matrix(x):
    """Undirected version of a matrix.

    Sparse input: the transposed entries are APPENDED to the COO lists (no
    dedup; entry order = all (r,c) then all (c,r)), exactly the layout the
    reference produces at utilities/math.py:13-20 - the device graph build sums
    or keeps the duplicates later depending on the layer family.
    Dense input: elementwise max with the transpose (math.py:21).
    """
    if not sparse.issparse(x):
        return np.maximum(x, x.T)
    x = x.tocoo()
    both = lambda a, b: np.concatenate([a, b])  # noqa: E731
    return sparse.coo_matrix((both(x.data, x.data), (both(x.row, x.col), both(x.col, x.row))),
                             shape=x.shape, dtype=x.dtype)


def convert_to_tensor(x, dtype=None, device=None):
    """Adjacency -> resident device graph (replaces math.py:24-56, which built a
    tf.SparseTensor and row-major reordered it).  Returns a DeviceGraph whose
    raw edge list is the same (row, col)-sorted, duplicates-kept order."""
    from ..graph import DeviceGraph
    return DeviceGraph.from_scipy(x, device=device)
