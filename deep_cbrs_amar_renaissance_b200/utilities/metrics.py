"""Per-user top-k of an explicit pair list (mirror of
/root/reference/src/utilities/metrics.py:11-34).

The reference sorts a pandas frame by (users asc, scores desc) - a stable sort - and
keeps head(k) per user with an O(U) filter loop; here one stable device radix sort of
(user, ~score) keys does both.  `top_k_metrics` (the shell-out to binaries/mimir.jar)
is out of scope: no JRE in the image."""
import numpy as np
import pandas as pd
import torch

from .. import ops


def top_k_predictions(predictions, users, items, k=5):
    """predictions [n,3] = (user idx, item idx offset by len(users), score).
    Returns a DataFrame(users, items, scores) with ORIGINAL ids, users ascending,
    scores descending, ties in input order."""
    predictions = np.asarray(predictions)
    dev = torch.device("cuda", torch.cuda.current_device())
    u = torch.from_numpy(np.ascontiguousarray(predictions[:, 0].astype(np.int64))).to(dev)
    s = torch.from_numpy(np.ascontiguousarray(predictions[:, 2].astype(np.float32))).to(dev)
    keep = ops.topk_pairs(u, s, len(users), k).cpu().numpy()
    df = pd.DataFrame()
    df['users'] = np.asarray(users)[predictions[keep, 0].astype(np.int64)]
    df['items'] = np.asarray(items)[predictions[keep, 1].astype(np.int64) - len(users)]
    df['scores'] = predictions[keep, 2]
    return df


def top_k_metrics(test_filepath, predictions_path):
    raise NotImplementedError("Precision/Recall/F1@k come from binaries/mimir.jar (Java, RiVal); no JRE in this "
                              "image - scope row (f)-2 (DESIGN.md)")
