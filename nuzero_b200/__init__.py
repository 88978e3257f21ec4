"""nuzero_b200 — B200-native self-play search engine behind NuZero's Explorer/Gamer/Game surface.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all search and game work
runs in hand-written sm_100a CUDA kernels reached through the C ABI of libnz_engine.so
(include/nz_engine.h).  There is no CPU fallback.
"""
from ._ffi import NzError  # noqa: F401

__all__ = ["NzError"]
