"""Import stub (test infrastructure): see metrohash.py."""


class BitArray:
    def __init__(self, *a, **k):
        raise RuntimeError("bitstring stub: KeylessCache is not part of the parity protocol")
