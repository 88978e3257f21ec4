"""GPU: the batched evaluation harness (MctsAgent with keep_subtree against RandomAgent, Testing/Tester.py:46-121)
against the CPU oracle's restatement of the same match, game by game: plies, actions, root visit counts, results."""
import os

import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mcts_player", [1, 2])
def test_ttt_mcts_vs_random_matches_oracle(mcts_player):
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import tic_tac_toe_spec
    from nuzero_b200.stubnet import DyadicStubNet
    from nuzero_b200.tester import BatchedTester
    from oracle import match
    from oracle.stubnet_np import stub_forward
    from oracle.ttt import TicTacToe

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    G = 24
    rng = np.random.default_rng(3 + mcts_player)
    tape = rng.random((G, 12))
    salts = list(range(100, 100 + G))
    t = BatchedTester(tic_tac_toe_spec(), cfg, G, lambda e: DyadicStubNet(e, salt=salts), policy_is_prob=True,
                      leaf_dtype=_ffi.F32, pool_nodes=4000)
    res = t.play(mcts_player, unif_tape=tape)
    for g in range(G):
        want = match.play_match(TicTacToe(), lambda s, sl=salts[g]: stub_forward(s, 9, sl), cfg, mcts_player, tape[g])
        assert res["actions"][g] == want["actions"]
        assert res["root_N"][g] == want["root_N"]
        assert int(res["terminal_value"][g]) == want["terminal_value"] and int(res["length"][g]) == want["length"]
    m, r, d = t.win_rates(res, mcts_is_first=(mcts_player == 1))
    assert abs(m + r + d - 1.0) < 1e-9


def test_scs_mcts_vs_random_matches_oracle():
    from nuzero_b200 import _ffi
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.stubnet import DyadicStubNet
    from nuzero_b200.tester import BatchedTester
    from oracle import match
    from oracle.scs import SCS, load_scenario
    from oracle.stubnet_np import stub_forward

    cfg = {k: dict(v) if isinstance(v, dict) else v for k, v in golden_io.load("ttt_p0_s25_salt0")["cfg"].items()}
    cfg["Simulation"]["mcts_simulations"] = 10
    path = os.path.join(golden_io.SCS_CONFIGS, "solo_soldier_config_5.yml")
    scn = ScsScenario(path, [1, 2])
    G = 6
    rng = np.random.default_rng(9)
    tape = rng.random((G, 64))
    salts = list(range(G))
    maps = [g % 2 for g in range(G)]
    t = BatchedTester(scn.spec(), cfg, G, lambda e: DyadicStubNet(e, salt=salts), policy_is_prob=True, leaf_dtype=_ffi.F32,
                      pool_nodes=20000, map_ids=maps, max_depth=64)
    res = t.play(0, unif_tape=tape)
    for g in range(G):
        game = SCS(load_scenario(path, seed=[1, 2][maps[g]]))
        want = match.play_match(game, lambda s, sl=salts[g]: stub_forward(s, game.get_num_actions(), sl), cfg, 0, tape[g])
        assert res["actions"][g] == want["actions"]
        assert res["root_N"][g] == want["root_N"]
        assert int(res["terminal_value"][g]) == want["terminal_value"]


def test_sweep_and_policy_vs_random_batches():
    """TestManager's "data" test: one batch per value of the changing parameter (here: the network), every pairing without
    a search played straight on the environment kernels."""
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import tic_tac_toe_spec
    from nuzero_b200.stubnet import DyadicStubNet
    from nuzero_b200.tester import BatchedTester
    from oracle import match
    from oracle.stubnet_np import stub_forward
    from oracle.ttt import TicTacToe

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    G = 16
    tape = np.random.default_rng(21).random((G, 12))
    t = BatchedTester(tic_tac_toe_spec(), cfg, G, lambda e: DyadicStubNet(e, salt=[0] * G), policy_is_prob=True,
                      leaf_dtype=_ffi.F32, pool_nodes=4000)
    data = t.sweep(("policy", "random"), [3, 4], lambda tt, v: tt.set_network(lambda e: DyadicStubNet(e, salt=[v] * G)),
                   num_runs=2, unif_tape=tape)
    assert [v for v, _ in data] == [3, 4]
    for v, (p1, p2, d) in data:
        tvs = []
        for g in range(G):
            nets = [lambda s, sl=v: stub_forward(s, 9, sl), None]
            tvs.append(match.play_agents(TicTacToe(), nets, cfg, ("policy", "random"), tape[g])["terminal_value"])
        tvs = np.array(tvs)
        assert (p1, p2, d) == (float((tvs > 0).mean()), float((tvs < 0).mean()), float((tvs == 0).mean()))


def test_from_config_sweeps_recurrent_iterations():
    """Configs/Testing/test_config.yaml of the reference (policy agent against random, iterations 4..12 — here 1..3): one
    batch per value, the forward rebuilt for each number of iterations; with a recurrent network the answers change with it."""
    import torch

    from nuzero_b200.engine import tic_tac_toe_spec
    from nuzero_b200.gamer import batched_forward
    from nuzero_b200.nets import RecurrentNet
    from nuzero_b200.network import Network_Manager
    from nuzero_b200.tester import BatchedTester

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    torch.manual_seed(0)
    model = RecurrentNet(2, 1, 32, 1, recall=True, policy_head="conv", value_head="reduce", value_activation="tanh", hex=False)
    nm = Network_Manager(model.to("cuda:0"))
    test_config = {
        "Test": {"test_type": "data", "Data": {"Variable": {"changing_agent": 1, "changing_parameter": {
            "name": "iterations", "Range": {"first": 1, "last": 3, "step": 1}}}, "Runs": {"num_runs": 1, "num_games_per_run": 32}}},
        "Agents": {"p1_agent": {"agent_type": "policy", "Network": {"recurrent_iterations": 2}}, "p2_agent": {"agent_type": "random"}},
    }
    G = 32
    seen = []
    t = BatchedTester(tic_tac_toe_spec(), cfg, G, lambda e: batched_forward(e, nm, 2), pool_nodes=4000)

    def make_net(e, iterations):
        seen.append(iterations)
        return batched_forward(e, nm, iterations)

    tape = np.random.default_rng(5).random((G, 12))
    data = t.test_from_config(test_config, make_net, unif_tape=tape)
    assert seen == [1, 2, 3] and [v for v, _ in data] == [1, 2, 3]
    for _, (p1, p2, d) in data:
        assert abs(p1 + p2 + d - 1.0) < 1e-9
    # the same batch again gives the same result (fixed tape, deterministic network)
    again = t.run_test_batch(("policy", "random"), unif_tape=tape)
    assert again == data[-1][1]
