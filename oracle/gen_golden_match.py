"""TEST INFRASTRUCTURE — golden fixtures for the evaluation path: the UNMODIFIED reference MctsAgent
(Testing/Agents/Generic/MctsAgent.py) against the UNMODIFIED RandomAgent (RandomAgent.py), driven by the loop of
Testing/Tester.py:46-121 (current agent chooses; an MctsAgent that is NOT moving runs update_subtree; game.step), in the
build container.  Writes tests/golden/match_*.npz.

    python -m oracle.gen_golden_match

Tester.py itself is not imported (it sleeps, prints, calls exit() and reads pettingzoo fields that Tic-Tac-Toe does not
have); the packages MctsAgent's star-imports pull in without using them (metrohash, bitstring, ruamel.yaml, matplotlib,
ray.runtime_env, progress) are replaced by empty stand-ins for the duration of this script.  The random agent's
np.random.choice consumes one uniform per call: np.random is seeded per game and the same stream is stored as the tape.
"""
import contextlib
import importlib.abc
import importlib.machinery
import io
import os
import sys
import types

import numpy as np

from . import ref_harness as rh
from .gen_golden import GOLDEN, search_config
from .stubnet_np import StubNetwork


class _Dummy(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None})


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = {"metrohash", "bitstring", "ruamel", "more_itertools", "matplotlib", "progress"}

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _Dummy(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def load_agents():
    rh.load()
    sys.meta_path.append(_Finder())
    sys.modules.setdefault("ray.runtime_env", _Dummy("ray.runtime_env"))
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from Testing.Agents.Generic.MctsAgent import MctsAgent
        from Testing.Agents.Generic.RandomAgent import RandomAgent
    return MctsAgent, RandomAgent


def load_policy_agent():
    load_agents()
    import Testing.Agents.Generic.PolicyAgent as mod

    return mod


def play(game, cfg, salt, mcts_player, seed):
    MctsAgent, RandomAgent = load_agents()
    net = StubNetwork(game.get_action_space_shape(), salt)
    mcts, rnd = MctsAgent(cfg, net, 2, None), RandomAgent()
    mcts.new_game(game)
    np.random.seed(seed)
    tape = np.random.random(512)
    np.random.seed(seed)
    actions, root_n, players = [], [], []
    with rh.parity_patches(None, identity_softmax=True):
        while not game.is_terminal():
            player = game.get_current_player()
            root = mcts.root_node
            if player == mcts_player:
                coords = mcts.choose_action(game)
                a = int(game.get_action_index(coords))
            else:
                coords = rnd.choose_action(game)
                a = int(game.get_action_index(coords))
                if mcts.keep_subtree:  # Tester.py:95-97
                    mcts.update_subtree(game, a)
            actions.append(a)
            root_n.append(int(root.visit_count))
            players.append(int(player))
            with contextlib.redirect_stdout(io.StringIO()):
                game.step(coords)
    return dict(actions=np.array(actions), root_N=np.array(root_n), players=np.array(players),
                terminal_value=int(game.get_terminal_value()), length=int(game.get_length()), unif_tape=tape,
                salt=salt, mcts_player=mcts_player, seed=seed)


def play_agents(game, cfg, kinds, salts, seed):
    """kinds: ("mcts" | "policy" | "random") for the player that moves first and for the other one; salts: the stub network
    of each.  The loop of Tester.py:46-121: the agent to move chooses; an MctsAgent that is not moving follows with
    update_subtree.  PolicyAgent.choose_action is the reference's (its scipy softmax rebound to a pass-through inside the
    imported module object, as for the Explorer: the stub emits probabilities)."""
    MctsAgent, RandomAgent = load_agents()
    pmod = load_policy_agent()
    shape = game.get_action_space_shape()

    def make(kind, salt):
        if kind == "mcts":
            return MctsAgent(cfg, StubNetwork(shape, salt), 2, None)
        if kind == "policy":
            return pmod.PolicyAgent(StubNetwork(shape, salt), 2, None)
        return RandomAgent()

    agents = [make(k, s) for k, s in zip(kinds, salts)]
    for a in agents:
        a.new_game(game)
    first = game.get_current_player()
    np.random.seed(seed)
    tape = np.random.random(512)
    np.random.seed(seed)
    actions, players, root_n = [], [], []
    saved = pmod.softmax
    pmod.softmax = lambda x, *a, **k: np.asarray(x)
    try:
        with rh.parity_patches(None, identity_softmax=True):
            while not game.is_terminal():
                player = game.get_current_player()
                idx = 0 if player == first else 1
                roots = [a.root_node if isinstance(a, MctsAgent) else None for a in agents]
                coords = agents[idx].choose_action(game)
                a_i = int(game.get_action_index(coords))
                other = agents[1 - idx]
                if isinstance(other, MctsAgent) and other.keep_subtree:
                    other.update_subtree(game, a_i)
                actions.append(a_i)
                players.append(int(player))
                root_n.append([-1 if r is None else int(r.visit_count) for r in roots])
                with contextlib.redirect_stdout(io.StringIO()):
                    game.step(coords)
    finally:
        pmod.softmax = saved
    # uniforms the global generator handed out (every np.random.choice call consumes exactly one)
    state_after = np.random.random()
    draws = int(np.nonzero(tape == state_after)[0][0])
    return dict(actions=np.array(actions), players=np.array(players), root_N=np.array(root_n).reshape(-1, 2),
                terminal_value=int(game.get_terminal_value()), length=int(game.get_length()), unif_tape=tape,
                kinds=np.array(kinds), salts=np.array(salts), seed=seed, draws=draws)


def main():
    ns = rh.load()
    out = {}
    cfg = search_config(25)
    for i, (mp, salt, seed) in enumerate([(1, 100, 1), (2, 101, 2), (1, 102, 3), (2, 103, 4)]):
        out["match_ttt_%d" % i] = (play(ns.tic_tac_toe(), cfg, salt, mp, seed), cfg, "ttt", None)
    cfg10 = search_config(10)
    for i, (name, map_seed, mp, salt, seed) in enumerate([("solo_soldier_config_5.yml", 1, 0, 0, 5), ("solo_soldier_config_5.yml", 2, 1, 1, 6),
                                                           ("mirrored_config_5.yml", None, 0, 2, 7)]):
        out["match_scs_%d" % i] = (play(rh.make_scs(name, map_seed), cfg10, salt, mp, seed), cfg10, name, map_seed)
    cases = [("ttt", None, ("policy", "random"), (5, 0), 11, cfg), ("ttt", None, ("random", "policy"), (0, 6), 12, cfg),
             ("ttt", None, ("mcts", "policy"), (7, 8), 13, cfg), ("ttt", None, ("policy", "mcts"), (9, 10), 14, cfg),
             ("solo_soldier_config_5.yml", 1, ("policy", "random"), (1, 0), 15, cfg10),
             ("mirrored_config_5.yml", None, ("random", "policy"), (0, 2), 16, cfg10),
             ("mirrored_config_5.yml", None, ("mcts", "policy"), (3, 4), 17, cfg10),
             ("unbalanced_config_5.yml", 3, ("policy", "mcts"), (5, 6), 18, cfg10),
             ("ttt", None, ("mcts", "mcts"), (11, 12), 19, cfg), ("mirrored_config_5.yml", None, ("mcts", "mcts"), (7, 8), 20, cfg10)]
    for i, (name, map_seed, kinds, salts, seed, c) in enumerate(cases):
        game = ns.tic_tac_toe() if name == "ttt" else rh.make_scs(name, map_seed)
        out["agents_%d" % i] = (play_agents(game, c, kinds, salts, seed), c, name, map_seed)
    os.makedirs(GOLDEN, exist_ok=True)
    for k, (rec, c, game, map_seed) in out.items():
        np.savez_compressed(os.path.join(GOLDEN, k + ".npz"), sims=c["Simulation"]["mcts_simulations"], game=game,
                            map_seed=-1 if map_seed is None else map_seed, **rec)
        print(k, "plies", len(rec["actions"]), "tv", rec["terminal_value"])


if __name__ == "__main__":
    main()
