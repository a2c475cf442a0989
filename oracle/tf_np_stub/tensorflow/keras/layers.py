"""Keras Layer protocol (lazy build on first call, add_weight) over numpy arrays.  Weights are drawn from a seeded
generator so that the golden script can record them: glorot_uniform like Keras, 'ones' / 'zeros' literal."""
import numpy as np

RNG = np.random.RandomState(1234)


def _shape_of(x):
    if isinstance(x, (list, tuple)):
        return [_shape_of(v) for v in x]
    return tuple(np.asarray(x).shape)


class Layer:
    def __init__(self, **kwargs):
        self.built = False
        self.weights = {}

    def add_weight(self, name=None, shape=None, initializer='glorot_uniform', regularizer=None, **kwargs):
        shape = tuple(int(s) for s in shape)
        if initializer == 'ones':
            w = np.ones(shape, np.float32)
        elif initializer == 'zeros':
            w = np.zeros(shape, np.float32)
        elif initializer == 'glorot_uniform':
            limit = np.sqrt(6.0 / (shape[-2] + shape[-1]))
            w = RNG.uniform(-limit, limit, size=shape).astype(np.float32)
        else:
            raise ValueError(initializer)
        self.weights[name] = w
        return w

    def build(self, input_shape):
        self.built = True

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            self.build(_shape_of(inputs))
            self.built = True
        return self.call(inputs, *args, **kwargs)


class Concatenate(Layer):
    def call(self, inputs, **kwargs):
        return np.concatenate([np.asarray(v, np.float32) for v in inputs], axis=-1)
