"""Batched evaluation: what Testing/Tester.Test_using_agents (Tester.py:46-121) does for one game and
TestManager.run_test_batch (TestManager.py:85-175) repeats over an actor pool, for G games at once on the device.

Pairing: MctsAgent (keep_subtree, training=False -> arg-max of the visit counts, no noise) against RandomAgent.  The
MCTS side searches on every ply — `choose_action` on its turn, `update_subtree` on the opponent's (MctsAgent.py:28-39) — so
every ply is: all games search until their simulations are done, then each game commits either the search's choice or the
random agent's action and re-roots on the child that was played (`nz_commit_moves`).  The random agent draws on the host
from the legal masks (`nz_env_mask`) exactly like `np.random.choice(num_actions, p=mask / mask.sum())`
(RandomAgent.py:10-15); a tape of uniforms makes it reproducible for the parity tests.
"""
import numpy as np
import torch

from . import _ffi
from .engine import EnvOps, SearchEngine


def _choice_from_uniform(mask_row, u):
    m = mask_row.astype(np.float64)
    p = m / m.sum()
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


class BatchedTester:
    def __init__(self, spec, search_config, n_games, net_factory, device="cuda:0", pool_nodes=None, map_ids=None,
                 policy_is_prob=False, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.F32, max_sims_per_launch=4, max_depth=None,
                 virtual_loss=1):
        """net_factory(engine) -> callable running the network on engine.leaf into engine.policy / engine.value
        (GraphedForward / FusedRecurrentForward / DyadicStubNet)."""
        self.e = SearchEngine(spec, search_config, n_games, False, device=device, pool_nodes=pool_nodes,
                              policy_is_prob=policy_is_prob, leaf_dtype=leaf_dtype, policy_dtype=policy_dtype,
                              auto_advance=False, max_sims_per_launch=max_sims_per_launch, max_depth=max_depth,
                              virtual_loss=virtual_loss)
        self.map_ids = None if map_ids is None else list(map_ids)
        if self.map_ids is not None:
            self.e.set_maps(self.map_ids)
            self.e.reset()
        self.net = net_factory(self.e)
        self.env = EnvOps(self.e)
        self.G = n_games

    def _search_all(self, max_launches=1_000_000):
        e = self.e
        for it in range(max_launches):
            e.advance()
            self.net()
            if (it & 7) == 7 or e.sims <= 8:
                ph = e.phases()
                if bool(((ph == _ffi.PHASE_MOVE_READY) | (ph == _ffi.PHASE_IDLE) | (ph == _ffi.PHASE_ERROR)).all()):
                    break
        e.raise_on_error()

    def play(self, mcts_player, unif_tape=None, rng=None, max_plies=100000):
        """Plays the G games to the end.  mcts_player: the value of get_current_player() on the MCTS agent's turns (1 or 2
        for Tic-Tac-Toe, 0 or 1 for SCS).  unif_tape [G, n]: uniforms of the random agent, one per random move (else `rng` / np.random).
        Returns dict(winner [G] (0 draw / 1 / 2 as Game.get_winner), terminal_value [G], length [G], actions: list per game,
        root_N: list per game)."""
        e, G = self.e, self.G
        rng = rng or np.random
        actions_hist = [[] for _ in range(G)]
        rootn_hist = [[] for _ in range(G)]
        alive = np.ones(G, dtype=bool)
        draws = np.zeros(G, dtype=np.int64)  # uniforms the random agent of each game has consumed
        for ply in range(max_plies):
            if not alive.any():
                break
            self._search_all()
            roots = e.gstate[:, 0].contiguous()
            st = self.env.status(roots, self.map_ids).cpu().numpy()
            masks = self.env.mask(roots, self.map_ids).cpu().numpy()
            root_idx = e.ctl[:, _ffi.CTL_ROOT].to(torch.int64)
            root_n = e.node_N[torch.arange(G, device=e.device), root_idx].cpu().numpy()
            chosen = e.ctl[:, _ffi.CTL_CHOSEN].to(torch.int64)
            base = e.node_link[torch.arange(G, device=e.device), root_idx, 0].to(torch.int64) & 0xFFFFFFFF
            chosen_action = ((e.node_link[torch.arange(G, device=e.device), base + chosen, 1].to(torch.int64) >> 16) & 0xFFFF).cpu().numpy()
            forced = np.full(G, -1, dtype=np.int32)
            for g in np.nonzero(alive)[0]:
                if st[g, 2] == mcts_player:
                    a = int(chosen_action[g])
                else:
                    u = float(unif_tape[g][draws[g]]) if unif_tape is not None else float(rng.random())
                    draws[g] += 1
                    a = _choice_from_uniform(masks[g], u)
                    forced[g] = a
                actions_hist[g].append(a)
                rootn_hist[g].append(int(root_n[g]))
            e.commit_moves(forced)
            e.raise_on_error()
            ph = e.phases().cpu().numpy()
            alive &= ph != _ffi.PHASE_IDLE
        roots = e.gstate[:, 0].contiguous()
        st = self.env.status(roots, self.map_ids).cpu().numpy()
        tv = st[:, 1]
        return dict(winner=np.where(tv > 0, 1, np.where(tv < 0, 2, 0)), terminal_value=tv, length=st[:, 3],
                    actions=actions_hist, root_N=rootn_hist)

    def win_rates(self, result, mcts_is_first):
        """(mcts wins, random wins, draws) as fractions, like TestManager's win-rate summary (TestManager.py:140-175)."""
        w = result["winner"]
        first, second = float((w == 1).mean()), float((w == 2).mean())
        return (first, second, float((w == 0).mean())) if mcts_is_first else (second, first, float((w == 0).mean()))
