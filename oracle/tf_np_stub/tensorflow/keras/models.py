from .layers import Layer


class Model(Layer):
    pass


class Sequential(Model):
    def __init__(self, layers=None, **kwargs):
        super().__init__()
        self.layers = list(layers or [])

    def call(self, x, **kwargs):
        for layer in self.layers:
            x = layer(x)
        return x
