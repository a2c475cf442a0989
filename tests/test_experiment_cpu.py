"""The experiment driver's host logic (no GPU): YAML forms, grid expansion, name lookup, keyword filtering - checked on
the REFERENCE's own config.yaml / econfigs/*.yaml and against the reference's own make_grid / nested_dict_update when
/root/reference is present (build container), on a written-out example otherwise."""
import copy
import glob
import inspect
import os
import sys
import types

import numpy as np
import pytest

from deep_cbrs_amar_renaissance_b200 import experiment as ex
from deep_cbrs_amar_renaissance_b200.data import loaders
from deep_cbrs_amar_renaissance_b200.utilities.utils import make_grid, nested_dict_update

REF = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists in the build container only")


def test_yaml_reads_the_float_forms_of_yaml_1_2(tmp_path):
    p = tmp_path / "c.yaml"
    p.write_text("a: 1e-4\nb: [1e-5, 1e-3, 0.5, 3]\nc: '1e-4'\nd: null\ne: True\nf: 1.5e+2\n")
    cfg = ex.load_yaml(str(p))
    assert cfg == dict(a=1e-4, b=[1e-5, 1e-3, 0.5, 3], c='1e-4', d=None, e=True, f=150.0)
    assert isinstance(cfg["a"], float) and isinstance(cfg["b"][3], int)


def test_make_grid_and_update_written_out():
    grid = {"model": {"name": ["basic.BasicGCN", "basic.BasicGAT"], "n_hiddens": [[8, 8]], "l2_regularizer": [1e-5, 1e-4]},
            "dataset": {"type_adjacency": ["unary"]}}
    got = make_grid(grid)
    assert len(got) == 4
    assert got[0] == {"model": {"name": "basic.BasicGCN", "n_hiddens": [8, 8], "l2_regularizer": 1e-5}, "dataset": {"type_adjacency": "unary"}}
    assert [g["model"]["l2_regularizer"] for g in got] == [1e-5, 1e-4, 1e-5, 1e-4]   # last listed key varies fastest
    with pytest.raises(ValueError):
        make_grid({"model": {"name": "basic.BasicGCN"}})
    base = {"model": {"name": "x", "embedding_dim": 16}, "seed": 42}
    out = nested_dict_update(copy.deepcopy(base), got[1])
    assert out["model"] == {"name": "basic.BasicGCN", "embedding_dim": 16, "n_hiddens": [8, 8], "l2_regularizer": 1e-4}
    assert out["seed"] == 42 and out["dataset"] == {"type_adjacency": "unary"}


def test_config_attribute_access():
    c = ex.Config({"model": {"name": "basic.BasicGCN", "n_hiddens": [8, 8]}, "seed": 1})
    assert c.model.name == "basic.BasicGCN" and c.seed == 1 and c.plain() == {"model": {"name": "basic.BasicGCN", "n_hiddens": [8, 8]}, "seed": 1}
    with pytest.raises(AttributeError):
        c.nope


@pytest.fixture(scope="module")
def reference_utils():
    """the reference's own utilities/utils.py, imported unmodified with an empty `mlflow` module in its way"""
    saved = {k: sys.modules.get(k) for k in ("mlflow", "utilities", "utilities.utils")}
    sys.modules["mlflow"] = types.ModuleType("mlflow")
    sys.modules.pop("utilities", None)
    sys.modules.pop("utilities.utils", None)
    sys.path.insert(0, os.path.join(REF, "src"))
    try:
        import utilities.utils as ru
        yield ru
    finally:
        sys.path.remove(os.path.join(REF, "src"))
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@needs_reference
def test_grids_of_every_reference_experiment_file(reference_utils):
    """every econfigs/*.yaml + experiments.yaml: same experiments, in the same order, as the reference's make_grid;
    every model class and loader named there exists with the keywords the driver will pass"""
    base = ex.load_yaml(os.path.join(REF, "config.yaml"))
    assert base["model"]["l2_regularizer"] == 1e-4 and base["parameters"]["optimizer"]["learning_rate"] == 0.001
    files = sorted(glob.glob(os.path.join(REF, "econfigs", "*.yaml"))) + [os.path.join(REF, "experiments.yaml")]
    assert len(files) == 11
    total = 0
    for path in files:
        cfg = ex.load_yaml(path)
        for grid in (cfg.get("grid") or {}).values():
            ours, theirs = make_grid(grid), reference_utils.make_grid(copy.deepcopy(grid))
            assert ours == theirs and [str(o) for o in ours] == [str(t) for t in theirs], path
            for exp in ours:
                merged = nested_dict_update(copy.deepcopy(base), exp)
                assert merged == reference_utils.nested_dict_update(copy.deepcopy(base), copy.deepcopy(exp))
                module, cls = merged["model"]["name"].split(".")
                model_class = getattr(__import__("deep_cbrs_amar_renaissance_b200.models." + module, fromlist=[cls]), cls)
                assert inspect.isclass(model_class), merged["model"]["name"]
                fn = getattr(loaders, merged["dataset"]["load_function_name"])
                kw = ex._by_signature(fn, merged["dataset"])
                assert {"train_ratings_filepath", "test_ratings_filepath"} <= set(kw)
                if merged["dataset"].get("props_triples_filepath") and "graph" in fn.__name__:
                    assert "props_triples_filepath" in kw and kw["type_adjacency"] in ("unary-uip", "unary-kg")
                total += 1
    assert total > 300
    opt = ex._by_signature(ex.Adam.__init__, base["parameters"]["optimizer"])
    assert opt == {"learning_rate": 0.001, "beta_1": 0.9}


@needs_reference
def test_multi_experimenter_lists_the_reference_grid(tmp_path, capsys):
    m = ex.MultiExperimenter(os.path.join(REF, "config.yaml"), os.path.join(REF, "econfigs", "basic-gnn.yaml"), str(tmp_path))
    assert len(m.experiments) == 90   # 6 grids x 5 families x 3 l2 values (econfigs/basic-gnn.yaml)
    first = next(iter(m.experiments.values()))
    assert first["model"]["name"] == "basic.BasicGCN" and first["dataset"]["load_function_name"] == "load_user_item_graph"


@needs_reference
def test_log_callback_times_the_same_batches_as_the_reference(monkeypatch):
    """utilities/keras.py:LogCallback is the reference's only timing instrumentation; its trace flags make it time every
    batch from index 500 of the first epoch on (the closing branch is unreachable once tracing started).  The
    reference's own class, imported unmodified over in-test stand-ins for mlflow / keras, and ours are driven through the
    same two epochs with a scripted clock and must record the same batch times and training time."""
    import importlib.util
    import logging
    import time

    from deep_cbrs_amar_renaissance_b200.utilities import keras as ours

    logged = {}
    fake = {"mlflow": types.ModuleType("mlflow"), "keras": types.ModuleType("keras"), "keras.utils": types.ModuleType("keras.utils"),
            "keras.utils.layer_utils": types.ModuleType("keras.utils.layer_utils"), "tensorflow": types.ModuleType("tensorflow"),
            "tensorflow.keras": types.ModuleType("tensorflow.keras")}
    fake["mlflow"].log_metric = lambda k, v, **kw: logged.__setitem__(k, v)
    fake["mlflow"].log_metrics = lambda d, **kw: None
    fake["keras.utils.layer_utils"].count_params = lambda ws: int(sum(int(w.numel()) for w in ws))
    fake["tensorflow.keras"].callbacks = types.SimpleNamespace(Callback=object)
    fake["tensorflow"].keras = fake["tensorflow.keras"]
    for name, mod in fake.items():
        monkeypatch.setitem(sys.modules, name, mod)
    spec = importlib.util.spec_from_file_location("reference_utilities_keras", os.path.join(REF, "src", "utilities", "keras.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    now = [0.0]
    monkeypatch.setattr(time, "perf_counter", lambda: now[0])
    monkeypatch.setattr(ours, "_device_sync", lambda: None)
    log = logging.getLogger("cbrs.test.logcallback")
    log.addHandler(logging.NullHandler())
    log.propagate = False
    cbs = [ref.LogCallback(log, 100), ours.LogCallback(log, 100)]
    t = 0.0
    for cb in cbs:
        now[0] = 5.0
        cb.on_train_begin({})
    for epoch in range(2):
        for cb in cbs:
            cb.on_epoch_begin(epoch, {})
        for b in range(530):
            dur = 1.0 + 0.001 * b + 10.0 * epoch
            for cb in cbs:
                now[0] = t
                cb.on_train_batch_begin(b, {})
                now[0] = t + dur
                cb.on_train_batch_end(b, {})
            t += dur + 0.25
        for cb in cbs:
            cb.on_epoch_end(epoch, {"loss": 0.5, "accuracy": 0.7})
    for cb in cbs:
        now[0] = 5.0 + t
        cb.on_train_end({})
    assert len(cbs[0].batch_times) == 30 + 530
    assert np.allclose(cbs[0].batch_times, cbs[1].batch_times, rtol=0, atol=1e-9)
    assert abs(cbs[0].get_batch_time() - cbs[1].get_batch_time()) < 1e-12
    assert abs(logged["batch_time"] - cbs[1].get_batch_time()) < 1e-12      # what the reference sends to MLflow
    assert abs(logged["training_time"] - cbs[1].training_time) < 1e-9

    class W:   # parameter counting (keras.py:10-22)
        def __init__(self, n):
            self.n = n

        def numel(self):
            return self.n

    class Model:
        trainable_weights, non_trainable_weights = [W(12), W(5)], [W(2)]

    assert ref.get_total_parameters(Model()) == ours.get_total_parameters(Model()) == (17, 2)
