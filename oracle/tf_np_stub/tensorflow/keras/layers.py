"""Keras Layer protocol (lazy build on first call, add_weight) and the few concrete layers the reference's model code
instantiates, over numpy float32 arrays.  Weights are drawn from a seeded generator so the golden scripts can record
them: glorot_uniform like Keras, 'ones' / 'zeros' literal.  Dense is the definition of keras.layers.Dense
(activation(x @ kernel + bias)), [3P] like everything in this package."""
import numpy as np

RNG = np.random.RandomState(1234)


def _shape_of(x):
    if isinstance(x, (list, tuple)):
        return [_shape_of(v) for v in x]
    if x is None:
        return None
    return tuple(getattr(x, "shape", np.asarray(x).shape))


def _activation(name):
    if name is None or name == 'linear':
        return lambda x: x
    if name == 'relu':
        return lambda x: np.maximum(x, np.float32(0))
    if name == 'sigmoid':
        return lambda x: (np.float32(1) / (np.float32(1) + np.exp(-x, dtype=np.float32))).astype(np.float32)
    if name == 'tanh':
        return lambda x: np.tanh(x, dtype=np.float32)
    raise ValueError(name)


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
            rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            limit = np.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
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


class Dense(Layer):
    def __init__(self, units, activation=None, **kwargs):
        super().__init__()
        self.units, self.activation = int(units), activation

    def build(self, input_shape):
        self.kernel = self.add_weight(name='kernel', shape=(int(input_shape[-1]), self.units), initializer='glorot_uniform')
        # Keras initialises the bias with zeros; a non-zero bias exercises the term, and the goldens record it
        self.bias = self.add_weight(name='bias', shape=(self.units,), initializer='zeros')
        self.bias[...] = RNG.uniform(-0.1, 0.1, size=self.units).astype(np.float32)

    def call(self, x, **kwargs):
        y = (np.asarray(x, np.float32) @ self.kernel).astype(np.float32) + self.bias
        return _activation(self.activation)(y.astype(np.float32))


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__()
        self.fn = _activation(activation)

    def call(self, x, **kwargs):
        return self.fn(np.asarray(x, np.float32))


class Dropout(Layer):
    def __init__(self, rate, **kwargs):
        super().__init__()
        self.rate = rate

    def call(self, x, **kwargs):
        return x  # inference
