"""tf.SparseTensor / tf.sparse.reorder as the reference uses them (src/utilities/math.py:37-56): a COO triple that
reorder() puts in row-major order, duplicates KEPT (tf.sparse.reorder does not coalesce)."""
import numpy as np


class SparseTensor:
    def __init__(self, indices, values, dense_shape):
        idx = np.asarray(indices)
        self.row, self.col = np.asarray(idx[:, 0]).ravel().astype(np.int64), np.asarray(idx[:, 1]).ravel().astype(np.int64)
        self.values = np.asarray(values)
        self.shape = tuple(int(s) for s in dense_shape)

    # what the layer stand-ins need
    def csr_arrays(self):
        n = self.shape[0]
        indptr = np.zeros(n + 1, np.int64)
        np.add.at(indptr, self.row + 1, 1)
        return np.cumsum(indptr), self.col.astype(np.int32), self.values.astype(np.float32)

    def to_scipy(self):
        from scipy import sparse
        return sparse.csr_matrix((self.values.astype(np.float32), (self.row, self.col)), shape=self.shape)

    def __matmul__(self, x):
        return np.asarray(self.to_scipy() @ np.asarray(x, np.float32), dtype=np.float32)


def reorder(sp):
    order = np.lexsort((sp.col, sp.row))  # stable: equal (row, col) keep their input order
    out = SparseTensor(np.stack([sp.row[order], sp.col[order]], axis=1), sp.values[order], sp.shape)
    return out
