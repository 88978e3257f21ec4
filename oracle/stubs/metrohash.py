"""Import stub (test infrastructure): metrohash is only used by Utils/Caches/KeylessCache.py, which the harness never
instantiates (cache_choice "disabled"); importing Utils.Functions.general_utils needs the name to exist."""


def hash64_int(*a, **k):
    raise RuntimeError("metrohash stub: KeylessCache is not part of the parity protocol")


hash128_int = hash64_int
