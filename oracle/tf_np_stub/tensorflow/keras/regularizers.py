class l2:
    def __init__(self, l2=0.01):
        self.l2 = float(l2)
