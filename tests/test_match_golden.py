"""Evaluation path (MctsAgent with keep_subtree against RandomAgent, Testing/Tester.py:46-121):
* CPU: oracle/match.py replays the fixtures generated from the real reference agents (oracle/gen_golden_match.py) bit for bit;
* GPU: the batched harness (nuzero_b200.tester.BatchedTester) replays the same fixtures through the C ABI."""
import os

import numpy as np
import pytest
import yaml

import golden_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(sims):
    cfg = yaml.safe_load(open(os.path.join(ROOT, "nuzero_b200", "configs", "a1_search_config.yaml")))
    cfg["Simulation"]["mcts_simulations"] = int(sims)
    return cfg


def _load(name):
    z = np.load(os.path.join(golden_io.GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _oracle_game(g):
    from oracle.scs import SCS, load_scenario
    from oracle.ttt import TicTacToe

    if str(g["game"]) == "ttt":
        return TicTacToe()
    seed = int(g["map_seed"])
    return SCS(load_scenario(os.path.join(golden_io.SCS_CONFIGS, str(g["game"])), seed=None if seed < 0 else seed))


@pytest.mark.parametrize("name", golden_io.names("match_"))
def test_oracle_match_replays_reference_agents(name):
    from oracle import match
    from oracle.stubnet_np import stub_forward

    g = _load(name)
    game = _oracle_game(g)
    A, salt = game.get_num_actions(), int(g["salt"])
    got = match.play_match(game, lambda s: stub_forward(s, A, salt), _cfg(g["sims"]), int(g["mcts_player"]), g["unif_tape"])
    assert got["actions"] == g["actions"].tolist()
    assert got["root_N"] == g["root_N"].tolist()
    assert got["players"] == g["players"].tolist()
    assert got["terminal_value"] == int(g["terminal_value"]) and got["length"] == int(g["length"])


@pytest.mark.parametrize("name", golden_io.names("agents_"))
def test_oracle_agents_replay_reference_agents(name):
    """PolicyAgent / MctsAgent / RandomAgent pairings recorded from the reference's own agent classes."""
    from oracle import match
    from oracle.stubnet_np import stub_forward

    g = _load(name)
    game = _oracle_game(g)
    A = game.get_num_actions()
    nets = [lambda s, sl=int(sl): stub_forward(s, A, sl) for sl in g["salts"]]
    got = match.play_agents(game, nets, _cfg(g["sims"]), [str(k) for k in g["kinds"]], g["unif_tape"])
    assert got["actions"] == g["actions"].tolist()
    assert got["players"] == g["players"].tolist()
    assert got["root_N"] == g["root_N"].tolist()
    assert got["draws"] == int(g["draws"])
    assert got["terminal_value"] == int(g["terminal_value"]) and got["length"] == int(g["length"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_io.names("match_"))
def test_batched_tester_replays_reference_agents(name):
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import tic_tac_toe_spec
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.stubnet import DyadicStubNet
    from nuzero_b200.tester import BatchedTester

    g = _load(name)
    G, salt = 3, int(g["salt"])  # the same game in three slots
    if str(g["game"]) == "ttt":
        spec, maps, kw = tic_tac_toe_spec(), None, dict(pool_nodes=4000)
    else:
        seed = int(g["map_seed"])
        scn = ScsScenario(os.path.join(golden_io.SCS_CONFIGS, str(g["game"])), [None if seed < 0 else seed])
        spec, maps, kw = scn.spec(), [0] * G, dict(pool_nodes=30000, max_depth=128)
    t = BatchedTester(spec, _cfg(g["sims"]), G, lambda e: DyadicStubNet(e, salt=[salt] * G), policy_is_prob=True,
                      leaf_dtype=_ffi.F32, map_ids=maps, **kw)
    res = t.play(int(g["mcts_player"]), unif_tape=np.tile(g["unif_tape"], (G, 1)))
    for s in range(G):
        assert res["actions"][s] == g["actions"].tolist()
        assert res["root_N"][s] == g["root_N"].tolist()
        assert int(res["terminal_value"][s]) == int(g["terminal_value"]) and int(res["length"][s]) == int(g["length"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_io.names("agents_"))
def test_batched_tester_replays_reference_agent_pairings(name):
    """PolicyAgent / MctsAgent / RandomAgent pairings recorded from the reference's agent classes, through the C ABI."""
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import tic_tac_toe_spec
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.stubnet import DyadicStubNet
    from nuzero_b200.tester import BatchedTester

    g = _load(name)
    G = 3
    kinds = [str(k) for k in g["kinds"]]
    if str(g["game"]) == "ttt":
        spec, maps, kw = tic_tac_toe_spec(), None, dict(pool_nodes=4000)
    else:
        seed = int(g["map_seed"])
        scn = ScsScenario(os.path.join(golden_io.SCS_CONFIGS, str(g["game"])), [None if seed < 0 else seed])
        spec, maps, kw = scn.spec(), [0] * G, dict(pool_nodes=30000, max_depth=128)
    salts = [int(s) for s in g["salts"]]
    # the fixture lists the first mover's agent first; the tester takes (p1_agent, p2_agent) of Tester.py:73-78
    first_is_p1 = int(g["players"][0]) == 1
    pair = kinds if first_is_p1 else kinds[::-1]
    psalts = salts if first_is_p1 else salts[::-1]
    both = kinds == ["mcts", "mcts"]
    salt_of = dict(zip(pair, psalts))
    t = BatchedTester(spec, _cfg(g["sims"]), G, lambda e: DyadicStubNet(e, salt=[psalts[0] if both else salt_of.get("mcts", 0)] * G),
                      policy_net_factory=lambda e: DyadicStubNet(e, salt=[salt_of.get("policy", 0)] * G),
                      second_net_factory=(lambda e: DyadicStubNet(e, salt=[psalts[1]] * G)) if both else None,
                      policy_is_prob=True, leaf_dtype=_ffi.F32, map_ids=maps, **kw)
    res = t.play_agents(pair, unif_tape=np.tile(g["unif_tape"], (G, 1)))
    col = kinds.index("mcts") if "mcts" in kinds else None
    for s in range(G):
        assert res["actions"][s] == g["actions"].tolist()
        if both:  # (p1 agent's root visits, p2 agent's) per ply against the fixture's (first mover's, other's)
            want = g["root_N"] if first_is_p1 else g["root_N"][:, ::-1]
            assert [list(x) for x in res["root_N"][s]] == want.tolist()
        elif col is not None:
            assert res["root_N"][s] == g["root_N"][:, col].tolist()
        assert int(res["draws"][s]) == int(g["draws"])
        assert int(res["terminal_value"][s]) == int(g["terminal_value"]) and int(res["length"][s]) == int(g["length"])
    p1, p2, d = t.run_test_batch(pair, unif_tape=np.tile(g["unif_tape"], (G, 1)))  # a second batch on the same engine
    tv = int(g["terminal_value"])
    assert (p1, p2, d) == (float(tv > 0), float(tv < 0), float(tv == 0))
