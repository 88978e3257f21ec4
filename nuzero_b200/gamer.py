"""Drop-in for Training/Gamer.py: `Gamer(...).play_game(cache=None) -> (stats, cache)` plus the batched
entry `play_games(n_games)` that runs thousands of games concurrently on the device.

Differences a caller can see: there is no Ray (the class is a plain object; `.remote`-style handles
for `buffer` / `shared_storage` are still accepted), and the inference cache argument is accepted and
returned untouched — leaves are batched across games, which is what the cache approximated.
"""
import numpy as np
import torch

from . import _ffi
from .engine import EnvOps, SearchEngine
from .network import GraphedForward


def batched_forward(engine, network, iters, use_graph=True):
    """Fastest available batched forward for `network` on the engine's leaf tensor: the fused tcgen05
    convolution kernel for RecurrentNet with recall and a filter count that is a multiple of 64, the nn.Module
    under a CUDA graph otherwise."""
    from .fastnet import FusedRecurrentForward
    from .nets import RecurrentNet

    model = network.get_model() if hasattr(network, "get_model") else network
    if isinstance(model, RecurrentNet) and model.recall and model.num_filters % 64 == 0:
        return FusedRecurrentForward(engine, network, iters, use_graph=use_graph)
    return GraphedForward(engine, network, iters, use_graph=use_graph)
from .selfplay import game_record, group_games


def _call(method, *args):
    if hasattr(method, "remote"):
        out = method.remote(*args)
        try:
            import ray  # pragma: no cover - only where the reference's actors are real

            return ray.get(out)
        except Exception:
            return out
    return method(*args)


class FinishedGame:
    """What Training/ReplayBuffer.save_game reads from a played game (ReplayBuffer.py:31-33)."""

    def __init__(self, states, policies, terminal_value, length, actions):
        self.state_history = states          # list of [1, C, R, Cc] float32 tensors (Gamer.py:65-66)
        self.child_policy = policies         # list of length-A visit-fraction lists (store_search_statistics)
        self.terminal_value = terminal_value
        self.length = length
        self.action_history = actions

    def get_state_from_history(self, i):
        return self.state_history[i]

    def make_target(self, i):
        return (self.terminal_value, self.child_policy[i])

    def get_terminal_value(self):
        return self.terminal_value

    def get_length(self):
        return self.length

    def get_winner(self):
        return 2 if self.terminal_value < 0 else (1 if self.terminal_value > 0 else 0)


def stats_of(rec):
    """The six statistics of Training/Gamer.py:42-50,81-92, accumulated move by move like the reference (the builtin
    sum() is a compensated summation since Python 3.12 and can differ from `+=` in the last bit)."""
    L = rec["length"]
    children = tree = 0
    bias = 0
    for k, n, b in zip(rec["n_children"], rec["root_N"], rec["bias"]):
        children += k
        tree += n
        bias += b
    return {
        "number_of_moves": L,
        "average_children": children / L,
        "average_tree_size": tree / L,
        "final_tree_size": rec["root_N"][-1],
        "average_bias_value": bias / L,
        "final_bias_value": rec["bias"][-1],
    }


def finished_game_of(rec, num_actions):
    pol = []
    for acts, n in zip(rec["child_actions"], rec["child_N"]):
        row = [0] * num_actions
        total = int(n.sum())
        for a, c in zip(acts.tolist(), n.tolist()):
            row[a] = c / total
        pol.append(row)
    states = [torch.from_numpy(s).unsqueeze(0) for s in rec["states"]]
    return FinishedGame(states, pol, rec["terminal_value"], rec["length"], list(rec["actions"]))


class Gamer:
    def __init__(self, buffer, shared_storage, game_class, game_args, game_index, search_config, recurrent_iterations,
                 cache_choice, size_estimate=10000, device="cuda:0", max_concurrent=4096, pool_nodes=None,
                 use_graph=True, seed=0, rng_tape=None):
        self.buffer, self.shared_storage = buffer, shared_storage
        self.game_class, self.game_args, self.game_index = game_class, game_args, game_index
        self.search_config, self.recurrent_iterations = search_config, recurrent_iterations
        self.cache_choice, self.size_estimate = cache_choice, size_estimate
        self.device, self.max_concurrent, self.pool_nodes = device, max_concurrent, pool_nodes
        self.use_graph, self.seed = use_graph, seed
        # parity runs: (gamma [M, K], uniforms [M, 3]) pre-drawn in the reference's call order (oracle/ref_harness.TapeRandom);
        # every game of a call replays the same tape.  None -> the device Philox generator.
        self.rng_tape = rng_tape
        self.time_to_stop = False
        self._template = None
        self.games_played = 0  # game ids handed out so far: every play_games() call draws from fresh random streams

    def _spec(self):
        if self._template is None:
            self._template = self.game_class(*self.game_args)
        return self._template.spec()

    def play_game(self, cache=None):  # Training/Gamer.py:39-97
        stats, _ = self.play_games(1)
        return stats[0], cache

    def play_games(self, n_games, concurrent=None):
        """Plays `n_games` self-play games, `concurrent` at a time, and ships each to the buffer.
        Returns (list of stats dicts, list of FinishedGame)."""
        network = _call(self.shared_storage.get)
        if hasattr(network, "check_devices"):
            network.check_devices()
        spec = self._spec()
        G = min(n_games, concurrent or self.max_concurrent)
        per_slot = -(-n_games // G)
        is_prob = bool(getattr(network, "outputs_probabilities", False))
        sims = int(self.search_config["Simulation"]["mcts_simulations"])
        pool = self.pool_nodes or (2 + sims * (spec.max_moves + 1)) * min(spec.max_children, 16)
        eng = SearchEngine(spec, self.search_config, G, True, device=self.device, pool_nodes=pool,
                           policy_is_prob=is_prob, leaf_dtype=_ffi.F32 if is_prob else _ffi.BF16,
                           policy_dtype=_ffi.F32, auto_advance=True, games_per_slot=per_slot,
                           max_sims_per_launch=4, seed=self.seed + self.game_index, arena_words=1 << 24,
                           tape_moves=0 if self.rng_tape is None else int(self.rng_tape[0].shape[0]),
                           tape_width=0 if self.rng_tape is None else int(self.rng_tape[0].shape[1]))
        if self.rng_tape is not None:
            gm, un = (np.asarray(a, dtype=np.float64) for a in self.rng_tape)
            eng.set_tapes(np.broadcast_to(gm, (G,) + gm.shape).copy(), np.broadcast_to(un, (G,) + un.shape).copy())
        # The device generator is keyed by (seed, game id, move): a fresh engine would hand out the ids 0, 1, ... again and
        # replay the same root noise / move-selection uniforms (the reference draws from numpy's global stream, so its
        # games differ from call to call).  Ids continue where the previous call stopped.
        uid0 = self.games_played
        if uid0:
            eng.ctl[:, _ffi.CTL_UID] += uid0
        self.games_played += G * per_slot
        if hasattr(network, "bind_engine"):
            net = network.bind_engine(eng)  # e.g. the CUDA stub network
        elif self.cache_choice not in (None, "disabled"):
            # cache_choice "dict" / "keyless" (Utils/Functions/general_utils.py:14-26): the device inference cache, exact keys,
            # sized from size_estimate like the reference's caches (Gamer.py:33-36)
            from .cache import CachedForward

            log2 = max(12, int(2 * max(1, self.size_estimate) - 1).bit_length())
            net = CachedForward(eng, lambda view: batched_forward(view, network, self.recurrent_iterations, use_graph=self.use_graph),
                                capacity_log2=min(log2, 26), min_rows=min(256, eng.rows))
        else:
            net = batched_forward(eng, network, self.recurrent_iterations, use_graph=self.use_graph)
        env = EnvOps(eng)
        recs = []
        it = 0
        while True:
            eng.advance()
            net()
            it += 1
            if it % 64 == 0:
                ph = eng.phases()
                if int(eng.arena_top[0]) > eng.c.arena_words // 2:
                    part, dropped = eng.drain_records()
                    if dropped:
                        raise _ffi.NzError("%d move records were dropped: the record arena is too small" % dropped)
                    recs += part
                if bool(((ph == _ffi.PHASE_IDLE) | (ph == _ffi.PHASE_ERROR)).all()):
                    break
        eng.raise_on_error()
        last, dropped = eng.drain_records()
        if dropped or int((eng.errors() & _ffi.ERR_ARENA_FULL).any()):
            raise _ffi.NzError("%d move records were dropped: the record arena is too small" % dropped)
        recs += last
        games = group_games(recs)
        if len(games) < n_games:
            raise _ffi.NzError("self-play produced %d complete games, %d were asked for" % (len(games), n_games))
        stats, finished = [], []
        for uid in sorted(games)[:n_games]:
            rec = game_record(games[uid], env, None if spec.kind == _ffi.GAME_TTT else 0)
            fg = finished_game_of(rec, eng.A)
            _call(self.buffer.save_game, fg, self.game_index)  # Gamer.py:95
            stats.append(stats_of(rec))
            finished.append(fg)
        eng.close()
        return stats, finished

    def play_forever(self):
        while not self.time_to_stop:
            self.play_game()

    def stop(self):
        self.time_to_stop = True
