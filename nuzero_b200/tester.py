"""Batched evaluation: what Testing/Tester.Test_using_agents (Tester.py:46-121) does for one game and
TestManager.run_test_batch (TestManager.py:85-175) repeats over an actor pool, for G games at once on the device.

Pairings: MctsAgent (keep_subtree, training=False -> arg-max of the visit counts, no noise), PolicyAgent (arg-max of the
network's policy, PolicyAgent.py:21-68) and RandomAgent in any combination with at most one MctsAgent.  The
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
                 virtual_loss=1, policy_net_factory=None, second_net_factory=None, second_search_config=None):
        """net_factory(engine) -> callable running the network on engine.leaf into engine.policy / engine.value
        (GraphedForward / FusedRecurrentForward / DyadicStubNet); policy_net_factory: the same for the PolicyAgent's
        network when it is not the MCTS agent's; second_net_factory (and optionally second_search_config): a second
        MctsAgent with its own trees — a second engine — for MCTS-against-MCTS pairings (it plays p2_agent)."""
        self.e = SearchEngine(spec, search_config, n_games, False, device=device, pool_nodes=pool_nodes,
                              policy_is_prob=policy_is_prob, leaf_dtype=leaf_dtype, policy_dtype=policy_dtype,
                              auto_advance=False, max_sims_per_launch=max_sims_per_launch, max_depth=max_depth,
                              virtual_loss=virtual_loss)
        self.map_ids = None if map_ids is None else list(map_ids)
        if self.map_ids is not None:
            self.e.set_maps(self.map_ids)
            self.e.reset()
        self.net = net_factory(self.e)
        self.policy_net = policy_net_factory(self.e) if policy_net_factory is not None else self.net
        self.env = EnvOps(self.e)
        self.G = n_games
        self.leaf_dtype, self.policy_is_prob = leaf_dtype, policy_is_prob
        self.e2 = self.net2 = None
        if second_net_factory is not None:
            self.e2 = SearchEngine(spec, second_search_config or search_config, n_games, False, device=device,
                                   pool_nodes=pool_nodes, policy_is_prob=policy_is_prob, leaf_dtype=leaf_dtype,
                                   policy_dtype=policy_dtype, auto_advance=False, max_sims_per_launch=max_sims_per_launch,
                                   max_depth=max_depth, virtual_loss=virtual_loss)
            if self.map_ids is not None:
                self.e2.set_maps(self.map_ids)
                self.e2.reset()
            self.net2 = second_net_factory(self.e2)

    def _search_all(self, max_launches=1_000_000, second=False):
        e, net = (self.e2, self.net2) if second else (self.e, self.net)
        for it in range(max_launches):
            e.advance()
            net()
            if (it & 7) == 7 or e.sims <= 8:
                ph = e.phases()
                if bool(((ph == _ffi.PHASE_MOVE_READY) | (ph == _ffi.PHASE_IDLE) | (ph == _ffi.PHASE_ERROR)).all()):
                    break
        e.raise_on_error()

    def play(self, mcts_player, unif_tape=None, rng=None, max_plies=100000):
        """MctsAgent against RandomAgent.  mcts_player: the value of get_current_player() on the MCTS agent's turns (1 or 2
        for Tic-Tac-Toe, 0 or 1 for SCS).  Returns the dict of `play_agents`, root_N as one list per game."""
        kinds = ("mcts", "random") if mcts_player == 1 else ("random", "mcts")
        res = self.play_agents(kinds, unif_tape=unif_tape, rng=rng, max_plies=max_plies)
        return res

    def _policy_actions(self, roots, masks_dev):
        """PolicyAgent.choose_action (PolicyAgent.py:21-68) for all games: the network on the root states, arg-max of its raw
        output when that action is legal, else arg-max over the legal ones.  -> (raw ok [G], raw [G], masked arg-max [G],
        masked mass is zero [G]) as numpy."""
        e, G = self.e, self.G
        enc = self.env.encode(roots, self.map_ids, dtype=self.leaf_dtype)
        e.leaf[:G].copy_(enc.view((G,) + tuple(e.leaf.shape[1:])))
        self.policy_net()
        pol = e.policy[:G].float()
        probs = pol if self.policy_is_prob else torch.softmax(pol, 1)
        raw = probs.argmax(1)
        m = masks_dev.to(probs.dtype)
        ok = masks_dev.gather(1, raw[:, None]).squeeze(1) != 0
        masked = probs * m
        return (ok.cpu().numpy(), raw.cpu().numpy(), masked.argmax(1).cpu().numpy(), (masked.sum(1) == 0).cpu().numpy())

    def play_agents(self, kinds, unif_tape=None, rng=None, max_plies=100000):
        """Plays the G games to the end between kinds[0] = Tester.Test_using_agents' p1_agent and kinds[1] = its p2_agent,
        each one of "mcts" (MctsAgent: one engine = one tree per game), "policy" (PolicyAgent) or "random" (RandomAgent);
        ("mcts", "mcts") uses the second engine.  As in the reference (Tester.py:73-78) p1_agent moves whenever
        get_current_player() == 1: the first mover of Tic-Tac-Toe (players 1 / 2) but the SECOND player of SCS (players
        0 / 1, SCS_Game.py:93).
        unif_tape [G, n]: the uniforms np.random.choice would consume, in order, per game (else `rng` / np.random).
        Returns dict(winner [G] (0 draw / 1 / 2 as Game.get_winner), terminal_value [G], length [G], actions: list per
        game, root_N: list per game (MCTS root visits after each ply's search; pairs for two MCTS agents), draws [G]
        uniforms consumed)."""
        e, G = self.e, self.G
        kinds = tuple(kinds)
        if any(k not in ("mcts", "policy", "random") for k in kinds):
            raise Exception("play_agents: kinds are 'mcts', 'policy' or 'random'")
        both = kinds == ("mcts", "mcts")
        if both and self.e2 is None:
            raise Exception("play_agents: two MCTS agents need second_net_factory (a second engine)")
        searching = "mcts" in kinds
        rng = rng or np.random
        actions_hist = [[] for _ in range(G)]
        rootn_hist = [[] for _ in range(G)]
        alive = np.ones(G, dtype=bool)
        draws = np.zeros(G, dtype=np.int64)  # uniforms each game's agents have consumed
        ar = torch.arange(G, device=e.device)
        if searching:
            e.reset()
            if both:
                self.e2.reset()
        else:
            states = self.env.reset(G, self.map_ids)

        def uniform(g):
            u = float(unif_tape[g][draws[g]]) if unif_tape is not None else float(rng.random())
            draws[g] += 1
            return u

        for ply in range(max_plies):
            if not alive.any():
                break
            if searching:
                self._search_all()
                if both:
                    self._search_all(second=True)
                roots = e.gstate[:, 0].contiguous()
            else:
                roots = states
            st = self.env.status(roots, self.map_ids).cpu().numpy()
            if not searching:
                alive &= st[:, 0] == 0
                if not alive.any():
                    break
            masks_dev = self.env.mask(roots, self.map_ids)
            masks = masks_dev.cpu().numpy()
            def search_result(eng):
                chosen = eng.ctl[:, _ffi.CTL_CHOSEN].to(torch.int64)  # the root is node 0, its children nodes 1..K
                return (eng.node_N[:, 0].cpu().numpy(), eng.node_action(ar, 1 + chosen).cpu().numpy())

            if searching:
                root_n, chosen_action = search_result(e)
                if both:
                    root_n2, chosen_action2 = search_result(self.e2)
            side = (st[:, 2] != 1).astype(np.int64)  # 0: p1_agent's turn (Tester.py:73)
            pol = self._policy_actions(roots, masks_dev) if "policy" in kinds else None
            forced = np.full(G, -1, dtype=np.int32)
            forced2 = np.full(G, -1, dtype=np.int32)
            played = np.zeros(G, dtype=np.int32)
            for g in np.nonzero(alive)[0]:
                kind = kinds[side[g]]
                if kind == "mcts" and both:
                    # the engine of the agent to move plays its choice; the other one follows it (update_subtree)
                    if side[g] == 0:
                        a = int(chosen_action[g])
                        forced2[g] = a
                    else:
                        a = int(chosen_action2[g])
                        forced[g] = a
                elif kind == "mcts":
                    a = int(chosen_action[g])
                elif kind == "policy":
                    ok, raw, alt, empty = pol
                    if ok[g]:
                        a = int(raw[g])
                    elif not empty[g]:
                        uniform(g)  # the discarded np.random.choice of PolicyAgent.py:52
                        a = int(alt[g])
                    else:
                        a = _choice_from_uniform(masks[g], uniform(g))
                    forced[g] = a
                else:
                    a = _choice_from_uniform(masks[g], uniform(g))
                    forced[g] = a
                played[g] = a
                actions_hist[g].append(a)
                if searching:
                    rootn_hist[g].append((int(root_n[g]), int(root_n2[g])) if both else int(root_n[g]))
            if searching:
                e.commit_moves(forced)
                e.raise_on_error()
                if both:
                    self.e2.commit_moves(forced2)
                    self.e2.raise_on_error()
                ph = e.phases().cpu().numpy()
                alive &= ph != _ffi.PHASE_IDLE
            else:
                idx = torch.as_tensor(np.nonzero(alive)[0], device=e.device)
                sub = states[idx].contiguous()
                maps = None if self.map_ids is None else [self.map_ids[g] for g in np.nonzero(alive)[0]]
                self.env.step(sub, played[alive], maps)
                states[idx] = sub
        roots = e.gstate[:, 0].contiguous() if searching else states
        st = self.env.status(roots, self.map_ids).cpu().numpy()
        tv = st[:, 1]
        return dict(winner=np.where(tv > 0, 1, np.where(tv < 0, 2, 0)), terminal_value=tv, length=st[:, 3],
                    actions=actions_hist, root_N=rootn_hist, draws=draws)

    def run_test_batch(self, kinds, num_runs=1, **kw):
        """TestManager.run_test_batch averaged over runs (TestManager.py:85-175, :262-275): (p1 win rate, p2 win rate, draws)."""
        acc = np.zeros(3)
        for _ in range(num_runs):
            w = self.play_agents(kinds, **kw)["winner"]
            acc += np.array([(w == 1).mean(), (w == 2).mean(), (w == 0).mean()]) / num_runs
        return tuple(float(x) for x in acc)

    def sweep(self, kinds, values, apply, num_runs=1, **kw):
        """The "data" test of TestManager.test_from_config (TestManager.py:222-262): for each value of the changing
        parameter `apply(tester, value)` re-configures the changing agent — recurrent iterations (set_network with a
        forward built for that many iterations) or a checkpoint — then `num_runs` batches are played.
        -> [(value, (p1 win rate, p2 win rate, draws))]."""
        out = []
        for v in values:
            apply(self, v)
            out.append((v, self.run_test_batch(kinds, num_runs, **kw)))
        return out

    def test_from_config(self, test_config, make_net, load_checkpoint=None, **kw):
        """The "data" test of TestManager.test_from_config (TestManager.py:177-280) from the same YAML dict
        (Configs/Testing/*.yaml): the agent types give the pairing, Test.Data.Variable the parameter that changes
        ("iterations": `make_net(engine, iterations)` builds the forward; "checkpoints": `load_checkpoint(number)` returns
        the network manager that `make_net(engine, iterations, network)` then wraps) and Test.Data.Runs the batches per value.
        Games per run = the number of game slots of this tester.  -> [(value, (p1 win rate, p2 win rate, draws))]."""
        agents, data = test_config["Agents"], test_config["Test"]["Data"]
        kinds = tuple(agents[k]["agent_type"] for k in ("p1_agent", "p2_agent"))
        for k in kinds:
            if k not in ("mcts", "policy", "random"):
                raise Exception("Bad agent type: %s (mcts | policy | random on the batched path)" % k)
        if test_config["Test"]["test_type"] != "data":
            raise Exception("the batched tester runs the 'data' tests only")
        var, runs = data["Variable"], int(data["Runs"]["num_runs"])
        who = int(var["changing_agent"])
        name = var["changing_parameter"]["name"]
        iters = {k: int(agents[k].get("Network", {}).get("recurrent_iterations", 2)) for k in ("p1_agent", "p2_agent")}

        def factory(agent_key, iterations, network=None):
            if network is None:
                return lambda e: make_net(e, iterations)
            return lambda e: make_net(e, iterations, network)

        def configure(value=None):
            net = {"p1_agent": None, "p2_agent": None}
            it = dict(iters)
            if value is not None and who in (1, 2):
                key = "p1_agent" if who == 1 else "p2_agent"
                if name == "iterations":
                    it[key] = int(value)
                elif name == "checkpoints":
                    if load_checkpoint is None:
                        raise Exception("changing checkpoints needs load_checkpoint(number)")
                    net[key] = load_checkpoint(int(value))
            need = [k for k, kind in zip(("p1_agent", "p2_agent"), kinds) if kind in ("mcts", "policy")]
            mcts_key = next((k for k, kind in zip(("p1_agent", "p2_agent"), kinds) if kind == "mcts"), None)
            pol_key = next((k for k, kind in zip(("p1_agent", "p2_agent"), kinds) if kind == "policy"), None)
            main = mcts_key or pol_key
            if main is not None:
                second = pol_key if (pol_key is not None and pol_key != main) else None
                self.set_network(factory(main, it[main], net[main]),
                                 None if second is None else factory(second, it[second], net[second]))
            return need

        if who == 0 or name == "none":
            configure()
            return [(None, self.run_test_batch(kinds, runs, **kw))]
        rng_cfg = var["changing_parameter"]["Range"]
        values = range(int(rng_cfg["first"]), int(rng_cfg["last"]) + 1, int(rng_cfg["step"]))
        return self.sweep(kinds, values, lambda tt, v: configure(v), num_runs=runs, **kw)

    def set_network(self, net_factory, policy_net_factory=None):
        """MctsAgent.set_network / set_recurrent_iterations (MctsAgent.py:57-64): swaps the evaluator of the engine."""
        self.net = net_factory(self.e)
        self.policy_net = policy_net_factory(self.e) if policy_net_factory is not None else self.net

    def win_rates(self, result, mcts_is_first):
        """(mcts wins, random wins, draws) as fractions, like TestManager's win-rate summary (TestManager.py:140-175)."""
        w = result["winner"]
        first, second = float((w == 1).mean()), float((w == 2).mean())
        return (first, second, float((w == 0).mean())) if mcts_is_first else (second, first, float((w == 0).mean()))
