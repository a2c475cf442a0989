"""Host-side logic that needs no GPU: optimiser configuration, weight shapes of the hybrid tweaks, builders."""
import math

import numpy as np
import pytest

from deep_cbrs_amar_renaissance_b200 import training
from deep_cbrs_amar_renaissance_b200.layers import FusionLayer
from deep_cbrs_amar_renaissance_b200.models.dense import (build_dense_classifier, build_dense_network,
                                                          build_residual_dense_network)
from deep_cbrs_amar_renaissance_b200.models.hybrid import HybridCBRS


def test_adam_from_config_forms():
    a = training.Adam.from_config({"learning_rate": 1e-2, "beta_1": 0.8, "name": "Adam"})
    assert (a.learning_rate, a.beta_1, a.beta_2, a.epsilon) == (1e-2, 0.8, 0.999, 1e-7)
    assert training.Adam.from_config(None).learning_rate == 1e-3 and training.Adam.from_config("adam").beta_1 == 0.9
    assert training.Adam.from_config(a) is a

    class KerasLike:  # what optimizers.Adam(**config) looks like to the caller (experiment.py:158)
        learning_rate, beta_1, beta_2, epsilon = 5e-4, 0.9, 0.99, 1e-8

    b = training.Adam.from_config(KerasLike())
    assert (b.learning_rate, b.beta_2, b.epsilon) == (5e-4, 0.99, 1e-8)
    with pytest.raises(NotImplementedError):
        training.Adam.from_config("sgd")


def test_adam_rate_is_the_keras_bias_correction():
    a = training.Adam(learning_rate=1e-3)
    for t in (1, 2, 10, 1000):
        assert math.isclose(a.rate(t), 1e-3 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t), rel_tol=1e-12)
    assert math.isclose(a.rate(1), 1e-3 * math.sqrt(0.001) / 0.1, rel_tol=1e-12)


def test_fusion_layer_weights_follow_the_reference_shapes():
    f = FusionLayer('attention')
    assert f.build_for(64, 64) == 64 and tuple(f.att_weight.shape) == (64, 64) and f.proj_weight is None
    g = FusionLayer('attention')   # narrower first input is projected up (fusion.py:24-32)
    assert g.build_for(32, 48) == 48 and g.proj_first is True and tuple(g.proj_weight.shape) == (32, 48)
    h = FusionLayer('attention')
    assert h.build_for(48, 32) == 48 and h.proj_first is False and tuple(h.proj_weight.shape) == (32, 48)
    c = FusionLayer('concatenate')
    assert c.build_for(10, 7) == 17 and c.weights == []
    with pytest.raises(ValueError):
        FusionLayer('sum')


def test_dense_builders():
    net = build_dense_network([48, 48], activation='relu')
    assert [l.activation for l in net.layers] == ['relu', 'relu']
    clf = build_dense_classifier([64, 64], n_classes=1, activation='relu')
    assert [(l.units, l.activation) for l in clf.layers] == [(64, 'relu'), (64, 'relu'), (1, 'sigmoid')]
    res = build_residual_dense_network([64, 64], activation='relu')
    assert [(l.units, l.activation) for l in res.layers] == [(64, 'relu'), (64, None)]   # dense.py:20-27
    assert [(l.units, l.activation) for l in build_dense_classifier([], n_classes=1).layers] == [(1, 'sigmoid')]


def test_hybrid_cbrs_parameter_counts_for_the_tweaks():
    units = [[48, 48], [256, 64], [64, 64]]
    plain = HybridCBRS(feature_based=True, dense_units=units, clf_units=[64, 64])
    plain.build_for(48, 768)
    att = HybridCBRS(feature_based=True, dense_units=units, clf_units=[64, 64], fusion_method='attention')
    att.build_for(48, 768)
    res = HybridCBRS(feature_based=True, dense_units=units, clf_units=[64, 64], residual=True)
    res.build_for(48, 768)
    n_plain = plain.count_params()
    # attention fuses the two 64-wide branches into ONE 64-wide vector: + 64*64 attention weights, and the
    # classifier's first kernel shrinks from 128x64 to 64x64
    assert att.count_params() == n_plain + 64 * 64 - 64 * 64
    # residual: the [64, 64] classifier body becomes the residual stack (same shapes); the head reads 64 features
    assert res.count_params() == n_plain
    with pytest.raises(ValueError):
        HybridCBRS(dense_units=[[48], [64], [32]], clf_units=[64], residual=True)
