"""Dense stacks (mirror of /root/reference/src/models/dense.py:4-27)."""
from ..layers.dense import DenseStack


def build_dense_network(units, **kwargs):
    return DenseStack(units, activation=kwargs.get('activation'))


def build_dense_classifier(units, n_classes, **kwargs):
    if n_classes != 1:
        raise NotImplementedError("the reference only builds binary classifiers (n_classes=1)")
    return DenseStack(units, activation=kwargs.get('activation'), last_units=1, last_activation='sigmoid')


def build_residual_dense_network(units, **kwargs):
    """dense.py:20-27: like build_dense_network but the last Dense is linear (the activation follows the residual add)."""
    units = list(units)
    return DenseStack(units, activation=kwargs.get('activation'), last_activation=None if units else "same")
