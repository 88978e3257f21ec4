"""GPU: the reference-shaped API (Explorer.run_mcts, Node views, Game classes, Gamer.play_game /
play_games, Network_Manager) driven exactly like Training/Gamer.py:64-79 drives the reference, and
checked against the fixtures generated from the real reference."""
import os

import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu
SCS_CFG = golden_io.SCS_CONFIGS


def _drive(game, net, cfg, training, tape=None):
    """The reference's own loop (Training/Gamer.py:64-79), verbatim in structure."""
    from nuzero_b200.search import Explorer, Node

    ex = Explorer(cfg, training, rng_tape=tape, pool_nodes=200000)
    root = Node(0)
    rec = dict(actions=[], root_N=[], n_children=[], child=[], bias=[])
    while not game.is_terminal():
        state = game.generate_network_input()
        game.store_state(state)
        action_i, chosen_child, root_bias = ex.run_mcts(game, net, root, 2, None)
        rec["actions"].append(action_i)
        rec["root_N"].append(root.get_visit_count())
        rec["n_children"].append(root.num_children())
        rec["bias"].append(root_bias)
        rec["child"].append({a: c.visit_count for a, c in root.children.items()})
        game.step(game.get_action_coords(action_i))
        game.store_search_statistics(root)
        root = chosen_child
    return rec


@pytest.mark.parametrize("name", ["ttt_p0_s100_salt1", "ttt_p1_s100_seed3_eps", "scs_p0_solo5", "scs_p1_unbalanced5_eps"])
def test_explorer_run_mcts_drop_in_matches_reference(name):
    from nuzero_b200.games.device_game import SCS_Game, tic_tac_toe
    from nuzero_b200.stubnet import StubNetworkManager

    g = golden_io.load(name)
    if str(g["game"]) == "ttt":
        game = tic_tac_toe()
    else:
        _, cfg_name, seed = str(g["game"]).split(":")
        game = SCS_Game(os.path.join(SCS_CFG, cfg_name), int(seed) or None)
    net = StubNetworkManager(game.get_action_space_shape(), g["salt"])
    tape = None
    if g["training"]:
        gm = np.zeros((g["gamma_tape"].shape[0] + 1, max(8, g["gamma_tape"].shape[1])))
        gm[:-1, : g["gamma_tape"].shape[1]] = g["gamma_tape"]
        un = np.zeros((gm.shape[0], 3))
        un[:-1] = g["unif_tape"]
        tape = (gm, un)
    rec = _drive(game, net, g["cfg"], g["training"], tape)
    assert rec["actions"] == g["actions"].tolist()
    assert rec["root_N"] == g["root_N"].tolist()
    np.testing.assert_array_equal(np.array(rec["bias"]), g["bias"])
    off = g["child_off"]
    for m in range(int(g["length"])):
        s = slice(off[m], off[m + 1])
        assert rec["child"][m] == dict(zip(g["child_actions"][s].tolist(), g["child_N"][s].tolist()))
        np.testing.assert_array_equal(game.state_history[m][0].numpy(), g["states"][m])
    assert game.get_terminal_value() == int(g["terminal_value"]) and game.get_winner() == int(g["winner"])
    np.testing.assert_allclose(np.array(game.child_policy, dtype=np.float64), g["child_policy"], rtol=1e-15)
    assert net.calls == int(g["net_calls"])  # one inference per expanded leaf, like the reference


def test_game_classes_expose_the_reference_duck_type():
    from nuzero_b200.games.device_game import SCS_Game, tic_tac_toe

    t = tic_tac_toe()
    assert t.get_action_space_shape() == (1, 3, 3) and t.get_num_actions() == 9 and t.get_state_shape() == (2, 3, 3)
    assert t.possible_actions().dtype == np.float64 and t.possible_actions().shape == (3, 3)
    assert t.get_current_player() == 1
    t.step((0, 1, 1))
    assert t.board[1][1] == 1 and t.get_current_player() == 2 and t.get_length() == 1
    c = t.shallow_clone()
    c.step((0, 0, 0))
    assert t.get_length() == 1 and c.get_length() == 2
    with pytest.raises(Exception):
        t.step((0, 1, 1))
    s = SCS_Game(os.path.join(SCS_CFG, "mirrored_config_5.yml"))
    assert s.get_action_space_shape() == (21, 5, 5) and s.get_state_shape() == (86, 5, 5)
    m = s.possible_actions()
    assert m.dtype == np.int8 and m.shape == (21, 5, 5) and m.sum() == 10
    assert s.generate_network_input().shape == (1, 86, 5, 5) and s.get_current_player() == 0
    a = int(np.flatnonzero(m.reshape(-1))[0])
    assert int(s.get_action_index(s.get_action_coords(a))) == a
    s.step(s.get_action_coords(a))
    assert s.get_length() == 1


def test_gamer_play_games_fills_the_replay_buffer():
    from nuzero_b200.gamer import Gamer
    from nuzero_b200.games.device_game import tic_tac_toe
    from nuzero_b200.replay import ReplayBuffer
    from nuzero_b200.stubnet import StubNetworkManager

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 40

    class Storage:
        def get(self):
            return StubNetworkManager((1, 3, 3), salt=5, uid_mul=1)

    buf = ReplayBuffer(window_size=1000, batch_size=8)
    gamer = Gamer(buf, Storage(), tic_tac_toe, [], 3, cfg, 2, "disabled", pool_nodes=20000)
    stats, cache = gamer.play_game()
    assert set(stats) == {"number_of_moves", "average_children", "average_tree_size", "final_tree_size",
                          "average_bias_value", "final_bias_value"}
    assert 5 <= stats["number_of_moves"] <= 9 and buf.played_games() == 1 and buf.len() == stats["number_of_moves"]
    all_stats, games = gamer.play_games(37, concurrent=16)
    assert len(all_stats) == 37 and buf.played_games() == 38
    assert buf.len() == stats["number_of_moves"] + sum(s["number_of_moves"] for s in all_stats)
    state, (value, policy), idx = buf.get_buffer()[-1]
    assert state.shape == (1, 2, 3, 3) and state.dtype == torch.float32 and idx == 3 and value in (-1, 0, 1)
    assert len(policy) == 9 and abs(sum(policy) - 1.0) < 1e-12
    batch = buf.get_sample(8, True, [])
    assert len(batch) == 8


def test_gamer_with_a_real_recurrent_net_in_a_cuda_graph():
    from nuzero_b200.gamer import Gamer
    from nuzero_b200.games.device_game import SCS_Game
    from nuzero_b200.nets import RecurrentNet, initialize_parameters
    from nuzero_b200.network import Network_Manager
    from nuzero_b200.replay import ReplayBuffer

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 10
    torch.manual_seed(0)
    model = RecurrentNet(67, 12, 64, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=True)
    initialize_parameters(model)
    nm = Network_Manager(model)  # 64 filters -> Gamer picks the fused tcgen05 forward

    class Storage:
        def get(self):
            return nm

    buf = ReplayBuffer(100, 4)
    gamer = Gamer(buf, Storage(), SCS_Game, [os.path.join(SCS_CFG, "solo_soldier_config_5.yml"), 1], 0, cfg, 3,
                  "disabled", pool_nodes=20000)
    stats, games = gamer.play_games(6, concurrent=6)
    assert len(stats) == 6 and all(4 <= s["number_of_moves"] <= 16 for s in stats)
    state, (value, policy), _ = buf.get_buffer()[0]
    assert state.shape == (1, 67, 5, 5) and len(policy) == 300
    p, v = nm.inference(state, False, 3)
    assert p.shape == (1, 12, 5, 5) and v.shape == (1, 1)
    # cache_choice "dict" (the reference's per-actor inference cache) -> the device inference cache: same games, same buffer
    buf2 = ReplayBuffer(100, 4)
    gamer2 = Gamer(buf2, Storage(), SCS_Game, [os.path.join(SCS_CFG, "solo_soldier_config_5.yml"), 1], 0, cfg, 3,
                   "dict", size_estimate=4000, pool_nodes=20000)
    stats2, games2 = gamer2.play_games(6, concurrent=6)
    assert [g.action_history for g in games2] == [g.action_history for g in games]
    assert stats2 == stats and len(buf2.get_buffer()) == len(buf.get_buffer())
    for (s0, (v0, p0), _), (s1, (v1, p1), _) in zip(buf.get_buffer(), buf2.get_buffer()):
        assert torch.equal(s0, s1) and v0 == v1 and p0 == p1


def test_mcts_agent_beats_random_agent_at_tic_tac_toe_and_keeps_its_subtree():
    """MctsAgent drives Explorer.run_mcts with training=False, re-rooting only on its own moves — the
    opponent's replies invalidate the kept sub-tree, which the drop-in detects and rebuilds."""
    from nuzero_b200.agents import MctsAgent, RandomAgent, play_match
    from nuzero_b200.games.device_game import tic_tac_toe
    from nuzero_b200.stubnet import StubNetworkManager

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 120
    cfg["Simulation"]["keep_subtree"] = True
    np.random.seed(0)

    class Flat:  # uniform policy, zero value: plain MCTS with terminal-value backups
        outputs_probabilities = True

        def inference(self, state, training, iters):
            return torch.full((1, 1, 3, 3), 1.0 / 9), torch.zeros(1, 1)

    results = []
    for g in range(6):
        class FreshRoot(MctsAgent):
            def choose_action(self, game):  # the reference's Tester calls new_game between games only;
                self.root_node = __import__("nuzero_b200.search", fromlist=["Node"]).Node(0)
                return super().choose_action(game)
        agent = FreshRoot(cfg, Flat(), pool_nodes=20000)
        winner, moves = play_match(tic_tac_toe(), [agent, RandomAgent()] if g % 2 == 0 else [RandomAgent(), agent])
        results.append((winner, g % 2))
    losses = sum(1 for w, side in results if w != 0 and w != (1 if side == 0 else 2))
    assert losses == 0, results
