"""ctypes binding of include/nz_engine.h.  There is deliberately no fallback: if the CUDA library is
missing the import of anything that computes raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NZ_ENGINE_LIB") or os.path.join(HERE, "libnz_engine.so")

NZ_ABI_VERSION = 3
GAME_TTT, GAME_SCS = 0, 1
F32, BF16 = 0, 1
PHASE_READY, PHASE_LEAF_PENDING, PHASE_MOVE_READY, PHASE_IDLE, PHASE_ERROR, PHASE_DESCENDING = range(6)
ERR_POOL_FULL, ERR_DEPTH, ERR_ILLEGAL, ERR_ARENA_FULL, ERR_CTABLE = 1, 2, 4, 8, 16
CTL_WORDS = 32
(CTL_PHASE, CTL_POOL_TOP, CTL_SIMS_DONE, CTL_ROOT_N0, CTL_ROOT_K, CTL_PATH_LEN, CTL_LEAF, CTL_ERROR,
 CTL_NOISED, CTL_MAP, CTL_MOVE, CTL_UID, CTL_GAMES_DONE, CTL_CHOSEN, CTL_HALF, CTL_N_PENDING,
 CTL_ROOT, CTL_N_SIMS, CTL_N_LEVELS, CTL_N_SCANNED, CTL_N_EXPAND, CTL_N_CREATED, CTL_N_MOVES, CTL_N_TERMINAL,
 CTL_LEAF_ROW, CTL_N_CACHE_HITS, CTL_N_CACHE_SHARED, CTL_LEAF_SLOT) = range(28)
REC_HDR = 12


class NzConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("game_kind", C.c_int32), ("n_games", C.c_int32),
        ("pool_nodes", C.c_int32), ("max_depth", C.c_int32), ("max_children", C.c_int32),
        ("mcts_simulations", C.c_int32), ("training", C.c_int32), ("policy_is_prob", C.c_int32),
        ("leaf_dtype", C.c_int32), ("policy_dtype", C.c_int32), ("auto_advance", C.c_int32),
        ("games_per_slot", C.c_int32), ("max_sims_per_launch", C.c_int32), ("record_detail", C.c_int32),
        ("number_of_softmax_moves", C.c_int32),
        ("pb_c_base", C.c_double), ("pb_c_init", C.c_double), ("value_factor", C.c_double),
        ("root_exploration_fraction", C.c_double), ("root_dist_alpha", C.c_double),
        ("root_dist_beta", C.c_double), ("epsilon_softmax_exploration", C.c_double),
        ("epsilon_random_exploration", C.c_double),
        ("seed", C.c_uint64),
        ("ctable_len", C.c_int32), ("tape_moves", C.c_int32), ("tape_width", C.c_int32),
        ("arena_words", C.c_int32),
        ("scs_desc", C.POINTER(C.c_int32)), ("scs_desc_len", C.c_int32), ("compact_on_reroot", C.c_int32),
        ("max_levels_per_launch", C.c_int32), ("virtual_loss_width", C.c_int32), ("node_state_cache", C.c_int32),
    ]


EXPORTS = {
    "nz_last_error": (C.c_char_p, []),
    "nz_abi_version": (C.c_int, []),
    "nz_config_bytes": (C.c_size_t, []),
    "nz_engine_create": (C.c_int, [C.POINTER(NzConfig), C.POINTER(C.c_void_p)]),
    "nz_engine_destroy": (None, [C.c_void_p]),
    "nz_engine_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "nz_engine_bind": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "nz_engine_buffer": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "nz_reset": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nz_advance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nz_commit_moves": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nz_env_state_words": (C.c_int, [C.c_void_p]),
    "nz_env_reset": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nz_env_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nz_env_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nz_env_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "nz_env_status": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nz_game_shape": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "nz_scs_static_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "nz_replay_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nz_cache_lookup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nz_cache_insert": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nz_engine_attach_cache": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "nz_cache_insert_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_void_p]),
    "nz_engine_set_lane": (C.c_int, [C.c_void_p, C.c_int]),
    "nz_engine_attach_expansions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "nz_im2col_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nz_hexconv_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nz_hexconv_set_trace": (C.c_int, [C.c_void_p]),
    "nz_noise_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_void_p]),
    "nz_stubnet_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


class NzError(Exception):
    """Raised where the reference raises a bare Exception (e.g. Games/SCS/SCS_Game.py:382)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libnz_engine.so is missing (%s): build it with `python -m nuzero_b200.build` — "
                "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.nz_abi_version() != NZ_ABI_VERSION:
            raise ImportError("libnz_engine.so ABI version mismatch")
        if L.nz_config_bytes() != C.sizeof(NzConfig):
            raise ImportError("libnz_engine.so was built from a different nz_config layout (%d bytes, binding %d)"
                              % (L.nz_config_bytes(), C.sizeof(NzConfig)))
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise NzError(lib().nz_last_error().decode())
