"""Dense stacks (mirror of /root/reference/src/models/dense.py:4-27)."""
from ..layers.dense import DenseStack


def build_dense_network(units, **kwargs):
    return DenseStack(units, activation=kwargs.get('activation'))


def build_dense_classifier(units, n_classes, **kwargs):
    if n_classes != 1:
        raise NotImplementedError("the reference only builds binary classifiers (n_classes=1)")
    return DenseStack(units, activation=kwargs.get('activation'), last_units=1, last_activation='sigmoid')


def build_residual_dense_network(units, **kwargs):
    raise NotImplementedError("residual classifier: tweaks grid only, outside the first hot-path bar (DESIGN.md)")
