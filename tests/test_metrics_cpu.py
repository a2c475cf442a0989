"""Precision/Recall/F1@k (the RiVal definitions restated in utilities/metrics.py) on a hand-checked case
and the results.tsv walk of top_k_metrics.  Host-only."""
import numpy as np

from deep_cbrs_amar_renaissance_b200.utilities.metrics import precision_recall_f1_at_k, top_k_metrics


def test_precision_recall_f1_hand_case():
    test = np.array([[1, 10, 1], [1, 11, 1], [1, 12, 0], [2, 10, 1], [3, 13, 0]])     # user 3 has no relevant item
    preds = np.array([[1, 10, 0.9], [1, 12, 0.8], [1, 11, 0.1], [2, 11, 0.7], [2, 10, 0.6], [3, 13, 0.5]])
    r = precision_recall_f1_at_k(test, preds, 2)
    # user 1: top2 = {10, 12} -> tp 1: P 1/2, R 1/2; user 2: top2 = {11, 10} -> tp 1: P 1/2, R 1/1
    assert r["users"] == 2 and abs(r["precision"] - 0.5) < 1e-12 and abs(r["recall"] - 0.75) < 1e-12
    assert abs(r["f1"] - 2 * 0.5 * 0.75 / 1.25) < 1e-12
    r1 = precision_recall_f1_at_k(test, preds, 1)
    assert abs(r1["precision"] - 0.5) < 1e-12 and abs(r1["recall"] - 0.25) < 1e-12  # user 1 hit, user 2 miss


def test_top_k_metrics_writes_results(tmp_path):
    test = np.array([[1, 10, 1], [1, 11, 1], [2, 10, 1]])
    np.savetxt(tmp_path / "test.tsv", test, fmt="%d", delimiter="\t")
    d = tmp_path / "preds" / "top_5"
    d.mkdir(parents=True)
    np.savetxt(d / "predictions_1.tsv", np.array([[1, 10, 0.9], [1, 11, 0.8], [2, 10, 0.7]]), fmt="%g", delimiter="\t")
    out = top_k_metrics(str(tmp_path / "test.tsv"), str(tmp_path / "preds"))
    assert abs(out[5]["recall"] - 1.0) < 1e-12 and abs(out[5]["precision"] - (2 / 5 + 1 / 5) / 2) < 1e-12
    # read back exactly as the reference's Experimenter.evaluate does (experiment.py:211-213)
    import pandas as pd
    results = pd.read_csv(d / "results.tsv", sep='\t', header=None)
    results = results.drop(0, axis=1).to_numpy().squeeze()
    assert results.shape == (3,) and abs(results[0] - out[5]["precision"]) < 1e-6 and abs(results[1] - 1.0) < 1e-6
    assert abs(results[2] - out[5]["f1"]) < 1e-6
