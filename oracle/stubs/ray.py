"""Import stub (test infrastructure): ray is transport only; make .remote()/ray.get local calls."""


def remote(*args, **kwargs):
    if len(args) == 1 and not kwargs and callable(args[0]):
        return args[0]

    def deco(cls):
        return cls

    return deco


def get(x, timeout=None):
    return x


class _Util:
    pass


util = _Util()
