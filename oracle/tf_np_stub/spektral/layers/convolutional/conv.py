from tensorflow.keras.layers import Layer


class Conv(Layer):
    """[3P] spektral Conv base class: a Keras layer with a static preprocess hook"""

    @staticmethod
    def preprocess(a):
        return a
