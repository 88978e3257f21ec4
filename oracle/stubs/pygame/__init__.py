"""Import stub (test infrastructure): SCS_Renderer.__init__ calls four pygame .init() hooks."""


class _Sub:
    def init(self, *a, **k):
        return None

    def __getattr__(self, name):
        raise AttributeError("pygame stub: %s is not available (rendering is out of scope)" % name)


display = _Sub()
fastevent = _Sub()
font = _Sub()
scrap = _Sub()


def init(*a, **k):
    return None
