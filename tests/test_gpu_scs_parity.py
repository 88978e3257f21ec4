"""GPU: SCS game kernels (step / legal mask / plane encoding / terminal scoring) and the full search
against fixtures generated from the real reference and against the CPU oracle.  Bit-exact."""
import os

import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu
SCS_CFG = golden_io.SCS_CONFIGS


def _parse_game(g):
    s = str(g["game"])
    if s.count(":") == 2:
        _, cfg, seed = s.split(":")
        return cfg, int(seed)
    return s[4:], int(g["seed"])


def _scenario(cfg, seeds):
    from nuzero_b200.games.scs_config import ScsScenario

    return ScsScenario(os.path.join(SCS_CFG, cfg), [s or None for s in seeds])


def _engine(scn, cfg, training, G, tapes=None, sims_budget=4, **kw):
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine

    tm, tw = (0, 0) if tapes is None else tapes[0].shape[1:3]
    kw.setdefault("pool_nodes", 300000)
    e = SearchEngine(scn.spec(), cfg, G, training, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                     auto_advance=True, games_per_slot=1, record_detail=True, tape_moves=tm, tape_width=tw,
                     max_sims_per_launch=sims_budget, **kw)
    if tapes is not None:
        e.set_tapes(*tapes)
    return e


@pytest.mark.parametrize("name", golden_io.names("scsenv_"))
def test_scs_env_kernels_match_reference_playout(name):
    from nuzero_b200.engine import EnvOps

    z = np.load(os.path.join(golden_io.GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    cfg_name, seed = _parse_game(g)
    scn = _scenario(cfg_name, [seed])
    assert scn.action_shape == tuple(g["action_shape"]) and scn.state_shape == tuple(g["state_shape"])
    np.testing.assert_array_equal(np.array(scn.maps[0][0]).reshape(scn.rows, scn.cols), g["sc_terrain"])
    e = _engine(scn, golden_io.load("ttt_p0_s25_salt0")["cfg"], False, 1, pool_nodes=64)
    env = EnvOps(e)
    n = 5  # the same game in 5 warps
    st = env.reset(n)
    for i, a in enumerate(g["actions"]):
        mask = env.mask(st).cpu().numpy()
        enc = env.encode(st).cpu().numpy()
        status = env.status(st).cpu().numpy()
        for k in (0, n - 1):
            np.testing.assert_array_equal(np.packbits(mask[k] != 0), g["masks"][i], err_msg="mask @%d" % i)
            np.testing.assert_array_equal(enc[k], g["states"][i], err_msg="state @%d" % i)
            assert status[k].tolist() == [0, 0, int(g["players"][i]), int(g["lengths"][i])], "status @%d" % i
        env.step(st, [int(a)] * n)
    status = env.status(st).cpu().numpy()
    assert status[0].tolist() == [int(g["terminal"]), int(g["terminal_value"]), int(g["final_player"]), int(g["final_length"])]
    np.testing.assert_array_equal(env.encode(st).cpu().numpy()[n - 1], g["final_state"])
    if g["terminal"]:
        with pytest.raises(Exception):
            env.step(st, [0] * n)
    # bf16 leaf rows are the rounded f32 rows
    from nuzero_b200 import _ffi
    st2 = env.reset(2)
    env.step(st2, [int(g["actions"][0])] * 2)
    f32 = env.encode(st2)
    bf = env.encode(st2, dtype=_ffi.BF16)
    assert torch.equal(f32.to(torch.bfloat16), bf)


def _play(e, salts, maps=None):
    from nuzero_b200.engine import EnvOps
    from nuzero_b200.selfplay import game_record, group_games, run_until_idle
    from nuzero_b200.stubnet import DyadicStubNet

    run_until_idle(e, DyadicStubNet(e, salt=salts), max_launches=400000)
    recs, dropped = e.drain_records()
    assert dropped == 0
    games = group_games(recs)
    env = EnvOps(e)
    return {uid: game_record(m, env, None if maps is None else maps[m[0]["slot"]]) for uid, m in games.items()}


@pytest.mark.parametrize("name", golden_io.names("scs_p"))
def test_scs_engine_matches_reference_golden(name):
    g = golden_io.load(name)
    cfg_name, seed = _parse_game(g)
    scn = _scenario(cfg_name, [seed])
    G = 2
    tapes = None
    if g["training"]:
        tm, tw = g["gamma_tape"].shape[0] + 1, max(8, g["gamma_tape"].shape[1])
        gm, un = np.zeros((G, tm, tw)), np.zeros((G, tm, 3))
        gm[:, : g["gamma_tape"].shape[0], : g["gamma_tape"].shape[1]] = g["gamma_tape"]
        un[:, : g["unif_tape"].shape[0]] = g["unif_tape"]
        tapes = (gm, un)
    # a pool far smaller than the whole game's allocations: the breadth-first compaction on re-root
    # has to kick in (several times for the long games) without changing a single bit
    small = {"scs_p0_test_s3": 13000, "scs_p0_randomized5": 7000, "scs_p1_randomized5": 7000}.get(name)  # +padding of the 64-byte aligned child runs
    e = _engine(scn, g["cfg"], g["training"], G, tapes, **({"pool_nodes": small} if small else {}))
    out = _play(e, [g["salt"]] * G)
    assert sorted(out) == list(range(G))
    for uid in range(G):
        golden_io.assert_record_matches(out[uid], g, check_trees=False)


@pytest.mark.parametrize("node_states", [True, False])
@pytest.mark.parametrize("cfg_name,seeds,training", [
    ("randomized_config_5.yml", [3, 4, 7, 9, 11, 12], True),
    ("mirrored_config_5.yml", [None], False),
    ("test_config.yml", [None], True),
])
def test_scs_engine_matches_oracle_many_games(cfg_name, seeds, training, node_states):
    """Fresh seeded games (several maps in one engine) against the CPU oracle, in both descent forms: expanded nodes keep
    their game state and a simulation steps the game once from the leaf's parent (node_state_cache, the default), or the
    scratch game is stepped at every tree level like Explorer.py:54-58."""
    from oracle import mcts, scs as oscs, selfplay
    from oracle.stubnet_np import stub_forward

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 12 if "test" in cfg_name else 24
    cfg["Exploration"]["epsilon_softmax_exploration"] = 0.15
    cfg["Exploration"]["epsilon_random_exploration"] = 0.15
    scn = _scenario(cfg_name, seeds)
    G = 2 * len(seeds) if len(seeds) > 1 else 4
    maps = [i % len(seeds) for i in range(G)]
    rng = np.random.Generator(np.random.Philox(5))
    TM, TW = 400, 64
    gm, un = rng.gamma(0.15, 1.0, size=(G, TM, TW)), rng.random(size=(G, TM, 3))
    # a level budget of 2 forces most descents to pause and resume across launches
    e = _engine(scn, cfg, training, G, (gm, un) if training else None, max_levels_per_launch=2 if training else 0,
                node_state_cache=node_states, **({"pool_nodes": 40000} if cfg_name.startswith("randomized") else {}))
    e.set_maps(maps)
    e.reset()
    salts = list(range(50, 50 + G))
    out = _play(e, salts, maps)
    assert len(out) == G
    for gi in range(G):
        sc = oscs.load_scenario(os.path.join(SCS_CFG, cfg_name), seeds[maps[gi]])
        tape = mcts.ReplayTape(gm[gi], un[gi]) if training else None
        ref = selfplay.play_game(oscs.SCS(sc), lambda s, sl=salts[gi]: stub_forward(s, sc.A, sl), cfg, training, True, tape)
        got = out[gi]
        assert got["actions"] == ref["actions"], "game %d" % gi
        assert got["root_N"] == ref["root_N"] and got["terminal_value"] == ref["terminal_value"]
        assert got["players"] == ref["players"]
        for m in range(ref["length"]):
            np.testing.assert_array_equal(got["child_actions"][m], ref["child_actions"][m])
            np.testing.assert_array_equal(got["child_N"][m], ref["child_N"][m])
            np.testing.assert_array_equal(got["child_W"][m], ref["child_W"][m])
            np.testing.assert_array_equal(got["child_prior"][m], ref["child_prior"][m])
            np.testing.assert_array_equal(got["states"][m], ref["states"][m])
            np.testing.assert_array_equal(got["masks"][m], ref["masks"][m])
        np.testing.assert_array_equal(np.array(got["root_W"]), np.array(ref["root_W"]))
