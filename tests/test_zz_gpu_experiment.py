"""The experiment driver end to end on the GPU (deep_cbrs_amar_renaissance_b200/experiment.py): a base config in the
reference's config.yaml layout + an experiment file with a 'grid' and a 'linear' section, on a synthetic dataset in the
reference's file formats -> per-run config.yaml / log.txt / metrics.json / predictions/top_<k>/predictions_1.tsv.
Covers the one-step GNN families, a Two-Step and a Two-Way model, the hybrid, and the scorers alone over pre-computed
embedding rows (basic.BasicRS / hybrid.HybridCBRS with the KGE loaders)."""
import json
import os

import numpy as np
import pytest
import yaml

pytestmark = pytest.mark.gpu


def write_inputs(root):
    from deep_cbrs_amar_renaissance_b200.data import synthetic
    paths = synthetic.write_dataset(str(root), n_users=60, n_items=40, n_ratings=1100, seed=5, n_props=30, n_triples=160,
                                    bert_dim=12)
    train = np.loadtxt(paths["train_ratings_filepath"], dtype=np.int64, delimiter="\t")
    rows = np.random.RandomState(2).standard_normal((int(train[:, :2].max()) + 1, 10)) * 0.3
    paths["graph_filepath"] = os.path.join(str(root), "768TransH.json")
    with open(paths["graph_filepath"], "w") as fp:
        json.dump({"ent_embeddings": rows.tolist()}, fp)
    base = {
        "details": "", "n_workers": 12, "seed": 42,
        "model": {"name": "basic.BasicRS", "embedding_dim": 8, "n_hiddens": [8, 8], "l2_regularizer": 1e-4,
                  "final_node": "concatenation", "item_node": "mean", "user_item_node": "mean", "aggregate": "mean",
                  "dropout_rate": 0.0, "n_layers": 2, "dense_units": [16, 16], "clf_units": [16, 16], "activation": "relu",
                  "feature_based": True, "fusion_method": "concatenate", "residual": False},
        "dataset": {"load_function_name": "load_graph_embeddings", "type_adjacency": "unary", "sparse_adjacency": True,
                    "symmetric_adjacency": True, "graph_filepath": paths["graph_filepath"],
                    "bert_user_filepath": paths["bert_user_filepath"], "bert_item_filepath": paths["bert_item_filepath"],
                    "props_triples_filepath": None, "train_ratings_filepath": paths["train_ratings_filepath"],
                    "test_ratings_filepath": paths["test_ratings_filepath"], "train_batch_size": 128, "test_batch_size": 64,
                    "shuffle": True},
        "parameters": {"epochs": 3, "optimizer": {"name": "Adam", "learning_rate": 0.01, "beta_1": 0.9},
                       "metrics": ["accuracy"], "loss": "binary_crossentropy"},
    }
    graph = {"load_function_name": "load_user_item_graph"}
    kg = dict(graph, type_adjacency="unary-kg", props_triples_filepath=paths["props_triples_filepath"])
    hyb_units = [[16, 16], [16, 8], [16, 16]]
    experiments = {
        "linear": {
            "base": None,   # the base config as it is: BasicRS over knowledge-graph embedding rows
            "hybrid-cbrs": {"model": {"name": "hybrid.HybridCBRS", "dense_units": hyb_units},
                            "dataset": {"load_function_name": "load_hybrid_embeddings"}},
            "two-step": {"model": {"name": "basic.BasicTSGCN"}, "dataset": kg},
            "two-way": {"model": {"name": "basic.BasicTWGraphSage"}, "dataset": dict(kg, user_properties=True)},
            "hybrid-gnn": {"model": {"name": "hybrid.HybridBertGCN", "dense_units": hyb_units},
                           "dataset": {"load_function_name": "load_user_item_graph_bert_embeddings"}},
            "broken": {"model": {"name": "basic.BasicGCN", "cache_neighbours": True}, "dataset": graph},
        },
        "grid": {"grid1": {"model": {"name": ["basic.BasicGCN", "basic.BasicGAT", "basic.BasicLightGCN"], "l2_regularizer": [1e-5]},
                           "dataset": {"load_function_name": ["load_user_item_graph"]}}},
    }
    cfg, exps = os.path.join(str(root), "config.yaml"), os.path.join(str(root), "experiments.yaml")
    with open(cfg, "w") as fp:
        yaml.safe_dump(base, fp)
    with open(exps, "w") as fp:
        yaml.safe_dump(experiments, fp)
    return cfg, exps, paths


def test_driver_runs_linear_and_grid_experiments(tmp_path):
    from deep_cbrs_amar_renaissance_b200 import experiment as ex
    from deep_cbrs_amar_renaissance_b200 import ops
    ops.check_device()
    cfg, exps, paths = write_inputs(tmp_path / "data")
    out = tmp_path / "runs"
    results = ex.main(["-c", cfg, "-e", exps, "--out", str(out), "--exp_name", "driver test"])
    assert len(results) == 9
    assert results["broken"] is None            # its exception was printed, the others still ran (experiment.py:299-302)
    ok = {k: v for k, v in results.items() if k != "broken"}
    test = np.loadtxt(paths["test_ratings_filepath"], dtype=np.int64, delimiter="\t")
    runs = sorted(os.listdir(out / "driver_test"))
    assert len(runs) == 9 and any("basic.BasicTSGCN-0.0001-concatenation" in r for r in runs)
    for name, m in ok.items():
        assert m is not None, name
        hist = m["history"]["loss"]
        assert len(hist) == 3 and hist[-1] < hist[0], (name, hist)
        assert 0.0 <= m["test_accuracy"] <= 1.0 and m["trainable_params"] > 0 and m["training_time"] > 0
        for k in (5, 10):
            assert 0.0 <= m["precision_at_%d" % k] <= 1.0 and 0.0 <= m["recall_at_%d" % k] <= 1.0
    for run in runs:
        art = out / "driver_test" / run / "artifacts"
        assert (art / "config.yaml").exists() and (art / "log.txt").exists() and (art / "metrics.json").exists()
        if "cache" in (art / "config.yaml").read_text() and "cache_neighbours: true" in (art / "config.yaml").read_text():
            continue
        for k in (5, 10):
            pred = np.loadtxt(art / "predictions" / ("top_%d" % k) / "predictions_1.tsv", delimiter="\t", ndmin=2)
            users, counts = np.unique(pred[:, 0], return_counts=True)
            assert counts.max() <= k and set(users.astype(np.int64)) <= set(test[:, 0])
            assert (np.diff(pred[:, 0]) >= 0).all()                      # users ascending
            same = np.diff(pred[:, 0]) == 0
            assert (np.diff(pred[:, 2])[same] <= 0).all()                # scores descending within a user
            assert (art / "predictions" / ("top_%d" % k) / "results.tsv").exists()


def test_relational_grid_runs_from_the_packaged_econfig(tmp_path):
    """row R end to end: deep_cbrs_amar_renaissance_b200/econfigs/basic-rgcn-uip-2relconf.yaml (basic.BasicRGCN +
    `relations:` in the dataset section) through the driver - loader keeps the edge types, the model trains, evaluates
    and writes its predictions; 'predicate' relations give one kernel per distinct predicate + the rating relation."""
    import deep_cbrs_amar_renaissance_b200 as pkg
    from deep_cbrs_amar_renaissance_b200 import experiment as ex
    from deep_cbrs_amar_renaissance_b200.data import loaders
    cfg, _, paths = write_inputs(tmp_path / "data")
    grids = ex.load_yaml(os.path.join(os.path.dirname(pkg.__file__), "econfigs", "basic-rgcn-uip-2relconf.yaml"))
    g2 = grids["grid"]["grid2"]
    assert g2["model"]["name"] == ["basic.BasicRGCN"] and g2["dataset"]["relations"] == ["node-range", "predicate"]
    g2["dataset"]["props_triples_filepath"] = [paths["props_triples_filepath"]]
    g2["model"]["l2_regularizer"] = [1e-4]
    exps = os.path.join(str(tmp_path), "rgcn.yaml")
    with open(exps, "w") as fp:
        yaml.safe_dump({"grid": {"grid2": g2}}, fp)
    results = ex.main(["-c", cfg, "-e", exps, "--out", str(tmp_path / "runs"), "--exp_name", "rgcn"])
    assert len(results) == 2 and all(m is not None for m in results.values())
    for m in results.values():
        hist = m["history"]["loss"]
        assert len(hist) == 3 and hist[-1] < hist[0] and 0.0 <= m["test_accuracy"] <= 1.0
    # the loader's typed graph: both directions of an edge carry the same relation; predicates are compacted
    train, _ = loaders.load_user_item_graph(paths["train_ratings_filepath"], paths["test_ratings_filepath"],
                                            paths["props_triples_filepath"], type_adjacency="unary-uip", relations="predicate")
    adj = train.adj_matrix
    raw = np.loadtxt(paths["props_triples_filepath"], dtype=np.int64, delimiter="\t")
    assert adj.n_rel == 1 + len(np.unique(raw[:, 2])) and len(adj.rel) == adj.coo.nnz
    half = adj.coo.nnz // 2
    assert np.array_equal(adj.rel[:half], adj.rel[half:]) and (adj.rel[:half] == 0).sum() == (train.ratings[:, 2] == 1).sum()
    plain, _ = loaders.load_user_item_graph(paths["train_ratings_filepath"], paths["test_ratings_filepath"],
                                            paths["props_triples_filepath"], type_adjacency="unary-uip")
    assert np.array_equal(plain.adj_matrix.row, adj.coo.row) and np.array_equal(plain.adj_matrix.col, adj.coo.col)
    params = sorted(m["trainable_params"] for m in results.values())
    assert params[0] < params[1]    # more relations, more kernels
