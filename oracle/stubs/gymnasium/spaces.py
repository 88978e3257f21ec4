"""Import stub: the hot path never touches the space objects, it only constructs them."""


class _Space:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class Discrete(_Space):
    pass


class Box(_Space):
    pass
