class Sequence:
    pass
