class Sequence:
    """Empty base class: keras.utils.Sequence contributes no behaviour the data path uses."""

    def __init__(self, *args, **kwargs):
        pass
