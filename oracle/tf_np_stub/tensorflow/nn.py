import numpy as np


def embedding_lookup(params, ids):
    return np.asarray(params)[np.asarray(ids, dtype=np.int64)]
