"""CPU: the C-ABI library builds, loads, and exports every symbol include/nz_engine.h declares
(no compute calls — there is no GPU in the build container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nz_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nz_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nuzero_b200 import _ffi, build

    build.build()
    L = ctypes.CDLL(_ffi.LIB_PATH)
    names = _declared_symbols()
    assert "nz_advance" in names and "nz_env_step" in names and len(names) >= 15
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    assert set(names) == set(_ffi.EXPORTS), set(names) ^ set(_ffi.EXPORTS)
    L.nz_config_bytes.restype = ctypes.c_size_t
    assert L.nz_config_bytes() == ctypes.sizeof(_ffi.NzConfig)


def test_create_layout_and_errors_without_gpu():
    from nuzero_b200 import _ffi

    L = _ffi.lib()
    c = _ffi.NzConfig()
    h = ctypes.c_void_p()
    assert L.nz_engine_create(ctypes.byref(c), ctypes.byref(h)) != 0
    assert b"ABI" in L.nz_last_error()
    c.abi_version = _ffi.NZ_ABI_VERSION
    c.game_kind = _ffi.GAME_TTT
    c.n_games, c.pool_nodes, c.max_depth, c.max_children = 8, 128, 12, 9
    c.mcts_simulations, c.ctable_len, c.max_sims_per_launch, c.arena_words = 10, 100, 4, 1024
    assert L.nz_engine_create(ctypes.byref(c), ctypes.byref(h)) == 0, L.nz_last_error()
    total = L.nz_engine_workspace_bytes(h)
    off, n = ctypes.c_size_t(), ctypes.c_size_t()
    assert L.nz_engine_buffer(h, b"nodes", ctypes.byref(off), ctypes.byref(n)) == 0
    assert n.value == 8 * 128 * 32 and off.value % 256 == 0 and off.value + n.value <= total
    assert L.nz_engine_buffer(h, b"nope", ctypes.byref(off), ctypes.byref(n)) != 0
    assert L.nz_reset(h, None) != 0 and b"not bound" in L.nz_last_error()
    shape = (ctypes.c_int32 * 6)()
    assert L.nz_game_shape(h, shape) == 0 and list(shape) == [1, 3, 3, 2, 3, 3]
    L.nz_engine_destroy(h)
