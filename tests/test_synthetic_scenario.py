"""The synthetic full-size scenario (nuzero_b200/configs/scs/synthetic_config_30.yml: 30 x 30, stacking limit 2, randomized map;
SURVEY.md 8d config 5 — the reference ships nothing of that size with stacking).
CPU: the reference's own SCS_Game loads it and agrees with the oracle on a random playout (masks, planes, players, outcome).
GPU: the environment kernels agree with the oracle on the same playouts, and a short self-play search is bit-identical."""
import contextlib
import io
import os

import numpy as np
import pytest

import golden_io

CFG = "synthetic_config_30.yml"
PATH = os.path.join(golden_io.SCS_CONFIGS, CFG)


def _playout(seed, steps=400):
    from oracle import scs as oscs

    g = oscs.SCS(oscs.load_scenario(PATH, seed))
    rng = np.random.default_rng(seed)
    trace = []
    while not g.is_terminal() and len(trace) < steps:
        mask = g.legal_mask().reshape(-1) != 0
        a = int(rng.choice(np.flatnonzero(mask)))
        trace.append((a, np.packbits(mask), np.asarray(g.encode()[0], dtype=np.float32), g.get_current_player()))
        g.step(a)
    return g, trace


@pytest.mark.parametrize("seed", [1, 2])
def test_reference_loads_the_synthetic_scenario_and_agrees_with_the_oracle(seed):
    from oracle import ref_harness as rh

    if not rh.available():
        pytest.skip("no reference tree / oracle/_ref here")
    ns = rh.load()
    path = rh.scs_config_path(CFG)
    if not os.path.isfile(path):  # source tree: the file is not the reference's; hand it this repo's copy with the key its loader insists on
        import tempfile

        import yaml

        data = yaml.safe_load(open(PATH))
        for props in data["Terrain"].values():
            props.setdefault("image_path", "")
        fd, path = tempfile.mkstemp(suffix=".yml")
        with os.fdopen(fd, "w") as fh:
            yaml.safe_dump(data, fh, sort_keys=False, default_flow_style=None)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ns.SCS_Game(path, seed)
    o, trace = _playout(seed, steps=60)
    for i, (a, mask, planes, player) in enumerate(trace):
        assert ref.get_current_player() == player, i
        np.testing.assert_array_equal(np.packbits(np.asarray(ref.possible_actions()).reshape(-1) != 0), mask, err_msg="mask @%d" % i)
        np.testing.assert_array_equal(np.asarray(ref.generate_network_input(), dtype=np.float32).reshape(planes.shape), planes, err_msg="planes @%d" % i)
        ref.step(ref.get_action_coords(a))
    assert ref.get_length() == len(trace)


@pytest.mark.gpu
def test_env_kernels_and_search_on_the_synthetic_scenario_match_the_oracle():
    import torch

    from nuzero_b200 import _ffi
    from nuzero_b200.engine import EnvOps, SearchEngine
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.selfplay import group_games
    from nuzero_b200.stubnet import DyadicStubNet
    from oracle import scs as oscs
    from oracle import selfplay as oselfplay
    from oracle.stubnet_np import stub_forward

    seeds = [1, 2]
    scn = ScsScenario(PATH, seeds)
    assert (scn.rows, scn.cols, scn.S) == (30, 30, 2)
    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 6
    e = SearchEngine(scn.spec(), cfg, 2, False, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32, auto_advance=True,
                     games_per_slot=1, record_detail=True, max_sims_per_launch=6, pool_nodes=60000, max_depth=128)
    env = EnvOps(e)
    for m, seed in enumerate(seeds):  # rules: masks, planes, player, outcome along a random playout of each map
        o, trace = _playout(seed)
        st = env.reset(1, [m])
        for i, (a, mask, planes, player) in enumerate(trace):
            np.testing.assert_array_equal(np.packbits(env.mask(st, [m])[0].cpu().numpy() != 0), mask, err_msg="mask @%d" % i)
            if i % 7 == 0:
                np.testing.assert_array_equal(env.encode(st, [m])[0].cpu().numpy(), planes, err_msg="planes @%d" % i)
            assert int(env.status(st, [m])[0, 2]) == player
            env.step(st, [a], [m])
        t, tv, _, ln = env.status(st, [m])[0].tolist()
        assert bool(t) == o.is_terminal() and ln == o.get_length() and (not t or tv == o.get_terminal_value())
    # search: the first moves of one game per map against the oracle's Explorer
    e.set_maps([0, 1])
    e.reset()
    net = DyadicStubNet(e, salt=[5, 6])
    for _ in range(36):
        e.advance()
        net()
    e.raise_on_error()
    recs, dropped = e.drain_records()
    assert dropped == 0
    checked = 0
    for slot, seed in enumerate(seeds):
        sc = oscs.load_scenario(PATH, seed)
        ref = oselfplay.play_game(_FirstMoves(oscs.SCS(sc), 5), lambda s, sl=5 + slot: stub_forward(s, sc.A, sl), cfg, False, True, keep_states=False)
        mine = sorted((r for r in recs if r["slot"] == slot), key=lambda r: r["move"])[:5]
        assert len(mine) >= 3
        for r in mine:
            k = r["move"]
            assert r["action"] == ref["actions"][k] and r["root_N"] == ref["root_N"][k]
            np.testing.assert_array_equal(r["child_N"], ref["child_N"][k])
            np.testing.assert_array_equal(r["child_prior"], ref["child_prior"][k])
            checked += 1
    assert checked >= 6


class _FirstMoves:
    """An oracle game that reports itself finished after `n` moves (the search of a 30 x 30 board is slow in Python)."""

    def __init__(self, game, n):
        self._g, self._n = game, n

    def is_terminal(self):
        return self._g.is_terminal() or self._g.get_length() >= self._n

    def __getattr__(self, name):
        return getattr(self._g, name)
