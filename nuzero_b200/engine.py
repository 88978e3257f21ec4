"""Host side of the search engine: owns the device memory (PyTorch tensors), builds the C-ABI
config from a NuZero search-config dict, and exposes the per-step calls.

PyTorch is plumbing here (allocation, streams, CUDA graphs); every search / game operation is a
kernel of libnz_engine.so.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _ffi
from ._ffi import NzConfig, NzError, check, lib

_TORCH_DT = {_ffi.F32: torch.float32, _ffi.BF16: torch.bfloat16}


class GameSpec:
    """What the engine needs to know about a game family."""

    def __init__(self, kind, desc=None, max_children=None, max_moves=None, name=""):
        self.kind = kind
        self.desc = None if desc is None else np.ascontiguousarray(desc, dtype=np.int32)
        self.max_children = max_children
        self.max_moves = max_moves
        self.name = name


def tic_tac_toe_spec():
    # Games/Tic_Tac_Toe/tic_tac_toe.py: 9 actions, at most 9 moves
    return GameSpec(_ffi.GAME_TTT, None, max_children=9, max_moves=9, name="tic_tac_toe")


def bias_table(search_config, n):
    """Rows (c[N], sqrt(N)) with c[N] = math.log((N + pb_c_base + 1) / pb_c_base) + pb_c_init — the
    exact expressions of Search/Explorer.py:103-112 evaluated with the host libm, so that the device
    never has to reproduce `log` bit for bit (sqrt is IEEE-exact on both sides; it shares the row so
    that one 16-byte load serves both)."""
    base, init = search_config["UCT"]["pb_c_base"], search_config["UCT"]["pb_c_init"]
    return np.array([(math.log((i + base + 1) / base) + init, math.sqrt(i)) for i in range(n)], dtype=np.float64)


class SearchEngine:
    def __init__(self, spec, search_config, n_games, training, device="cuda:0", pool_nodes=None,
                 max_depth=None, policy_is_prob=False, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.F32,
                 auto_advance=True, games_per_slot=0, max_sims_per_launch=8, record_detail=False,
                 seed=0, tape_moves=0, tape_width=0, arena_words=1 << 22, ctable_len=None, compact=True, max_levels_per_launch=0,
                 virtual_loss=1, node_state_cache=True):
        if not search_config["Simulation"].get("keep_subtree", True):
            # Gamer/MctsAgent never reset the root when keep_subtree is False
            raise NzError("only keep_subtree: True is supported")
        self.lib = lib()
        self.spec = spec
        self.cfg_dict = search_config
        self.device = torch.device(device)
        sims = int(search_config["Simulation"]["mcts_simulations"])
        ex = search_config["Exploration"]
        max_moves = spec.max_moves or 512
        if ctable_len is None:
            ctable_len = sims * (max_moves + 1) + 2
        # slot layout (csrc/common.cuh): node 0 = root, nodes 1.. = the root's children, general pool from `g0`
        self.g0 = (2 + max(int(spec.max_children), 32)) & ~1
        if pool_nodes is None:
            pool_nodes = self.g0 + min(sims * max_moves, 1 << 16) * (spec.max_children + 1)
        pool_nodes = (max(int(pool_nodes), self.g0 + 8) + 1) & ~1  # even: child runs (and their state rows) start on even nodes
        if max_depth is None:
            max_depth = min(max_moves + 2, 256)
        c = NzConfig()
        c.abi_version = _ffi.NZ_ABI_VERSION
        c.game_kind = spec.kind
        c.n_games = n_games
        c.pool_nodes = int(pool_nodes)
        c.max_depth = int(max_depth)
        c.max_children = int(spec.max_children)
        c.mcts_simulations = sims
        c.training = int(bool(training))
        c.policy_is_prob = int(bool(policy_is_prob))
        c.leaf_dtype, c.policy_dtype = leaf_dtype, policy_dtype
        c.auto_advance = int(bool(auto_advance))
        c.games_per_slot = int(games_per_slot)
        c.max_sims_per_launch = int(max_sims_per_launch)
        c.record_detail = int(bool(record_detail))
        c.number_of_softmax_moves = int(ex["number_of_softmax_moves"])
        c.pb_c_base = float(search_config["UCT"]["pb_c_base"])
        c.pb_c_init = float(search_config["UCT"]["pb_c_init"])
        c.value_factor = float(ex["value_factor"])
        c.root_exploration_fraction = float(ex["root_exploration_fraction"])
        c.root_dist_alpha = float(ex["root_dist_alpha"])
        c.root_dist_beta = float(ex["root_dist_beta"])
        c.epsilon_softmax_exploration = float(ex["epsilon_softmax_exploration"])
        c.epsilon_random_exploration = float(ex["epsilon_random_exploration"])
        c.seed = int(seed) & ((1 << 64) - 1)
        c.ctable_len = int(ctable_len)
        c.tape_moves, c.tape_width = int(tape_moves), int(tape_width)
        c.arena_words = int(arena_words)
        c.compact_on_reroot = int(bool(compact) and bool(auto_advance))
        c.max_levels_per_launch = int(max_levels_per_launch)
        c.virtual_loss_width = int(virtual_loss)
        # SCS: expanded nodes keep their game state (one game step per simulation instead of one per tree level); costs
        # pool_nodes x state_words x 4 bytes per slot.  Ignored by Tic-Tac-Toe.
        c.node_state_cache = int(bool(node_state_cache))
        if spec.desc is not None:
            self._desc = spec.desc
            c.scs_desc = self._desc.ctypes.data_as(C.POINTER(C.c_int32))
            c.scs_desc_len = int(self._desc.size)
        self.c = c
        h = C.c_void_p()
        check(self.lib.nz_engine_create(C.byref(c), C.byref(h)))
        self.h = h
        shape = (C.c_int32 * 6)()
        check(self.lib.nz_game_shape(h, shape))
        self.action_shape = tuple(shape[0:3])
        self.state_shape = tuple(shape[3:6])
        self.A = int(np.prod(self.action_shape))
        self.G = n_games
        self.V = max(1, int(virtual_loss))   # leaves one game may have waiting at the network
        self.rows = n_games * self.V         # rows of the leaf / policy / value tensors (row g * V + j)
        self.P = int(pool_nodes)
        self.sims = sims
        self.state_words = self.lib.nz_env_state_words(h)
        if self.device.type != "cuda":
            raise NzError("the search engine needs a CUDA device; there is no CPU fallback")
        nbytes = self.lib.nz_engine_workspace_bytes(h)
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        check(self.lib.nz_engine_bind(h, C.c_void_p(self.workspace.data_ptr()), nbytes))
        # node pool: 32-byte records {prior f64, W f64, N i32, flags u32 (bit 0 noised prior, bits 16-31 action),
        # first child u32, n_children u32}; per slot node 0 is the root and nodes 1..K(root) are its children
        nodes_i = self.view("nodes", torch.int32).view(n_games, self.P, 8)
        nodes_f = self.view("nodes", torch.float64).view(n_games, self.P, 4)
        self.node_prior = nodes_f[:, :, 0]
        self.node_W = nodes_f[:, :, 1]
        self.node_N = nodes_i[:, :, 4]
        self.node_flags = nodes_i[:, :, 5]
        self.node_base = nodes_i[:, :, 6]
        self.node_K = nodes_i[:, :, 7]
        self.ctl = self.view("ctl", torch.int32).view(n_games, _ffi.CTL_WORDS)
        self.gstate = self.view("gstate", torch.int32).view(n_games, 1 + self.V, self.state_words)  # root, leaf state(s)
        self.arena = self.view("arena", torch.int32)
        self.arena_top = self.view("arena_top", torch.int32)   # [words used, records dropped, records written, -]
        self.rec_index = self.view("rec_index", torch.int32)   # arena offset of every record
        self.dense_count = self.view("dense_count", torch.int32)  # [0]: leaf rows the last launch handed out (dense rows)
        self.dense_rows = self.view("dense_rows", torch.int32)    # dense row -> game slot
        self.view("ctable", torch.float64).copy_(torch.from_numpy(bias_table(search_config, ctable_len).reshape(-1)))
        if spec.kind == _ffi.GAME_SCS:
            img = np.zeros(self.buffer_bytes("scs_static"), dtype=np.uint8)
            check(self.lib.nz_scs_static_image(h, C.c_void_p(img.ctypes.data), img.size))
            self.view("scs_static", torch.uint8).copy_(torch.from_numpy(img))
        self.leaf = torch.zeros((self.rows,) + self.state_shape, dtype=_TORCH_DT[leaf_dtype], device=self.device)
        self.policy = torch.zeros((self.rows, self.A), dtype=_TORCH_DT[policy_dtype], device=self.device)
        self.value = torch.zeros((self.rows,), dtype=torch.float32, device=self.device)
        self.launches = 0
        self._pre_advance = None
        self.reset()

    # -- memory -----------------------------------------------------------------------------------
    def buffer_bytes(self, name):
        off, n = C.c_size_t(), C.c_size_t()
        check(self.lib.nz_engine_buffer(self.h, name.encode(), C.byref(off), C.byref(n)))
        return n.value

    def view(self, name, dtype):
        off, n = C.c_size_t(), C.c_size_t()
        check(self.lib.nz_engine_buffer(self.h, name.encode(), C.byref(off), C.byref(n)))
        return self.workspace[off.value: off.value + n.value].view(dtype)

    def set_tapes(self, gamma, unif):
        """Parity mode: pre-drawn gamma noise [G, tape_moves, tape_width] and uniforms [G, tape_moves, 3]."""
        g = self.view("gamma_tape", torch.float64).view(self.G, self.c.tape_moves, self.c.tape_width)
        u = self.view("unif_tape", torch.float64).view(self.G, self.c.tape_moves, 3)
        g.copy_(torch.as_tensor(gamma, dtype=torch.float64))
        u.copy_(torch.as_tensor(unif, dtype=torch.float64))

    def set_maps(self, map_ids):
        self.ctl[:, _ffi.CTL_MAP] = torch.as_tensor(map_ids, dtype=torch.int32, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- per-step calls -------------------------------------------------------------------------------
    def reset(self):
        check(self.lib.nz_reset(self.h, self._stream()))
        self.arena_top.zero_()

    def advance(self):
        """One search launch over all slots (consumes self.policy / self.value, fills self.leaf)."""
        if self._pre_advance is not None:
            self._pre_advance()  # CachedForward(pipeline=True): picks the lane's tensors and waits for that lane's last network call
        check(self.lib.nz_advance(self.h, C.c_void_p(self.leaf.data_ptr()), C.c_void_p(self.policy.data_ptr()),
                                  C.c_void_p(self.value.data_ptr()), self._stream()))
        self.launches += 1

    def commit_moves(self, actions=None):
        ptr = None
        if actions is not None:
            self._actions = torch.as_tensor(actions, dtype=torch.int32, device=self.device)
            ptr = C.c_void_p(self._actions.data_ptr())
        check(self.lib.nz_commit_moves(self.h, ptr, self._stream()))

    # -- results --------------------------------------------------------------------------------------
    def node_action(self, g, idx):
        """Action that leads to node(s) `idx` of slot(s) `g` (bits 16-31 of the record's flags word)."""
        return (self.node_flags[g, idx] >> 16) & 0xFFFF

    def research_root(self, g=0):
        """Manual mode: let slot g run another `mcts_simulations` on its current root (MctsAgent.update_subtree,
        MctsAgent.py:35-39).  The kernel takes the root's visit count as ROOT_N0 + SIMS_DONE."""
        self.ctl[g, _ffi.CTL_ROOT_N0] = self.node_N[g, 0]
        self.ctl[g, _ffi.CTL_SIMS_DONE] = 0
        self.ctl[g, _ffi.CTL_PHASE] = _ffi.PHASE_READY

    def phases(self):
        return self.ctl[:, _ffi.CTL_PHASE]

    def errors(self):
        return self.ctl[:, _ffi.CTL_ERROR]

    def counters(self):
        c = self.ctl[:, _ffi.CTL_N_SIMS:_ffi.CTL_N_TERMINAL + 1].to(torch.int64) & 0xFFFFFFFF
        s = c.sum(0).tolist()
        keys = ["sims", "levels", "scanned", "expansions", "created", "moves", "terminal_leaves"]
        out = dict(zip(keys, s))
        out["games"] = int((self.ctl[:, _ffi.CTL_GAMES_DONE].to(torch.int64) & 0xFFFFFFFF).sum())
        out["cache_hits"] = int((self.ctl[:, _ffi.CTL_N_CACHE_HITS].to(torch.int64) & 0xFFFFFFFF).sum())
        out["cache_shared"] = int((self.ctl[:, _ffi.CTL_N_CACHE_SHARED].to(torch.int64) & 0xFFFFFFFF).sum())
        return out

    def raise_on_error(self):
        e = self.errors()
        fatal = e & ~(_ffi.ERR_CTABLE | _ffi.ERR_ARENA_FULL)
        if bool((fatal != 0).any()):
            g = int(torch.nonzero(fatal)[0])
            bits = int(e[g])
            names = [n for n, b in (("node pool full", 1), ("path deeper than max_depth", 2),
                                    ("illegal action", 4), ("record arena full", 8)) if bits & b]
            raise NzError("search engine fault in slot %d: %s" % (g, ", ".join(names)))

    def drain_records(self):
        """Copy the move-record arena to the host, parse it, and reset the arena."""
        top = self.arena_top.cpu()
        used, dropped = int(top[0]), int(top[1])
        used = min(used, self.c.arena_words)
        words = self.arena[:used].cpu().numpy().view(np.uint32)
        self.arena_top.zero_()
        recs = parse_records(words, self.state_words)
        return recs, dropped

    def close(self):
        if self.h:
            self.lib.nz_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _f64(lo, hi):
    return np.array([(int(hi) << 32) | int(lo)], dtype=np.uint64).view(np.float64)[0]


def parse_records(words, state_words):
    """Arena layout (csrc/mcts.cuh write_record): 12 header words, the compact root state before the
    move, (action, N) per root child, and optionally (W, prior) f64 pairs per child."""
    out = []
    pos, n = 0, len(words)
    H = _ffi.REC_HDR
    while pos < n:
        ln = int(words[pos])
        if ln < H or pos + ln > n:
            break
        r = words[pos:pos + ln]
        K = int(r[2] >> 16)
        flags = int(r[3] >> 24)
        body = r[H + state_words:]
        rec = {
            "uid": int(r[1]), "move": int(r[2] & 0xFFFF), "n_children": K,
            "action": int(r[3] & 0xFFFF), "player": int((r[3] >> 16) & 0xFF),
            "game_end": bool(flags & 2), "terminal_value": ((flags >> 2) & 3) - 1,
            "root_N": int(r[4]), "root_W": float(_f64(r[5], r[6])), "slot": int(r[7]) & 0xFFFFF, "map": int(r[7]) >> 20,
            "bias": float(_f64(r[8], r[9])), "length": int(r[10]), "child": int(r[11]),
            "state": r[H:H + state_words].copy(),
            "child_actions": body[0:2 * K:2].astype(np.int32),
            "child_N": body[1:2 * K:2].astype(np.int64),
        }
        if flags & 1:
            d = body[2 * K:2 * K + 4 * K].copy().view(np.float64).reshape(K, 2)
            rec["child_W"], rec["child_prior"] = d[:, 0].copy(), d[:, 1].copy()
        out.append(rec)
        pos += ln
    return out


class EnvOps:
    """Batched Game-interface calls over compact device states (Games/Game.py:3-106):
    possible_actions / step / generate_network_input / is_terminal ... for n states at once."""

    def __init__(self, engine):
        self.e = engine

    def _maps(self, map_ids, n):
        if map_ids is None:
            return None, None
        t = torch.as_tensor(map_ids, dtype=torch.int32, device=self.e.device).contiguous()
        return t, C.c_void_p(t.data_ptr())

    def _states(self, states):
        return torch.as_tensor(np.asarray(states).astype(np.int64) if not torch.is_tensor(states) else states,
                               device=self.e.device).to(torch.int32).contiguous().view(-1, self.e.state_words)

    def reset(self, n, map_ids=None):
        e = self.e
        st = torch.zeros((n, e.state_words), dtype=torch.int32, device=e.device)
        keep, mp = self._maps(map_ids, n)
        check(e.lib.nz_env_reset(e.h, C.c_void_p(st.data_ptr()), mp, n, e._stream()))
        return st

    def step(self, states, actions, map_ids=None):
        """In place; raises like the reference on an illegal action (SCS_Game.py:382)."""
        e = self.e
        n = states.shape[0]
        act = torch.as_tensor(actions, dtype=torch.int32, device=e.device).contiguous()
        err = torch.zeros(n, dtype=torch.int32, device=e.device)
        keep, mp = self._maps(map_ids, n)
        check(e.lib.nz_env_step(e.h, C.c_void_p(states.data_ptr()), mp, C.c_void_p(act.data_ptr()),
                                C.c_void_p(err.data_ptr()), n, e._stream()))
        if bool(err.any()):
            raise NzError("Tried to play an illegal action!")
        return states

    def mask(self, states, map_ids=None):
        e = self.e
        n = states.shape[0]
        out = torch.zeros((n, e.A), dtype=torch.uint8, device=e.device)
        keep, mp = self._maps(map_ids, n)
        check(e.lib.nz_env_mask(e.h, C.c_void_p(states.data_ptr()), mp, C.c_void_p(out.data_ptr()), n, e._stream()))
        return out

    def encode(self, states, map_ids=None, dtype=_ffi.F32):
        e = self.e
        n = states.shape[0]
        out = torch.zeros((n,) + e.state_shape, dtype=_TORCH_DT[dtype], device=e.device)
        keep, mp = self._maps(map_ids, n)
        check(e.lib.nz_env_encode(e.h, C.c_void_p(states.data_ptr()), mp, C.c_void_p(out.data_ptr()), dtype, n,
                                  e._stream()))
        return out

    def status(self, states, map_ids=None):
        """-> int32 [n, 4]: terminal, terminal_value, current player, length."""
        e = self.e
        n = states.shape[0]
        out = torch.zeros((n, 4), dtype=torch.int32, device=e.device)
        keep, mp = self._maps(map_ids, n)
        check(e.lib.nz_env_status(e.h, C.c_void_p(states.data_ptr()), mp, C.c_void_p(out.data_ptr()), n, e._stream()))
        return out
