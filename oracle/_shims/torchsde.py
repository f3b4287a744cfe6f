class BrownianTree:  # never used on the spaced/DDIM path
    def __init__(self, *a, **k):
        raise NotImplementedError
