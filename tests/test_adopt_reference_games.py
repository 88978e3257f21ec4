"""Reference game objects on the drop-in surface (nuzero_b200/games/adopt.py).

CPU (needs the reference tree or its byte code, oracle/_ref): the conversion rests on replaying a reference object's
`action_history`; the test plays random games on the reference's own tic_tac_toe / SCS_Game, converts the history to flat
action indices with `action_index` and replays it through the oracle's rules: same planes, same player, same outcome.
GPU: the reference's own MctsAgent class (Testing/Agents/Generic/MctsAgent.py:14-39), holding THIS package's Explorer, plays
reference game objects; the choices equal the committed fixtures of the all-reference run (oracle/gen_golden_match.py)."""
import os

import numpy as np
import pytest

import golden_io


def _reference():
    from oracle import ref_harness as rh

    if not rh.available():
        pytest.skip("no reference tree / oracle/_ref here")
    return rh, rh.load()


def _quiet(fn, *a):
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)


def test_conversion_rules_ttt_word_and_scs_history_replay():
    from nuzero_b200.games.adopt import _family, action_index, ttt_state_word
    from oracle import scs as oscs
    from oracle.ttt import TicTacToe

    rh, ns = _reference()
    rng = np.random.default_rng(3)
    for trial in range(8):  # Tic-Tac-Toe: the word built from the reference object's fields = the oracle's state after the same moves
        g, o = ns.tic_tac_toe(), TicTacToe()
        assert _family(g) == "ttt"
        while not g.is_terminal() and g.get_length() < 3 + trial:
            legal = np.flatnonzero(np.asarray(g.possible_actions()).reshape(-1))
            a = int(rng.choice(legal))
            g.step(g.get_action_coords(a))
            o.step(a)
        term = bool(g.is_terminal())
        w = ttt_state_word(g.board, g.get_length(), term, g.get_terminal_value() if term else 0)
        planes = np.array([(w >> i) & 1 for i in range(18)], dtype=np.float32)
        np.testing.assert_array_equal(planes, np.asarray(g.generate_state_image(), dtype=np.float32).reshape(-1))
        np.testing.assert_array_equal(planes, np.asarray(o.encode()[0], dtype=np.float32).reshape(-1))
        assert (w >> 18) & 15 == o.get_length() and bool((w >> 22) & 1) == o.is_terminal()
        assert ((w >> 23) & 3) - 1 == (o.get_terminal_value() if o.is_terminal() else 0)
    cfg = "mirrored_config_5.yml"  # SCS: the action history replayed through the oracle's rules reaches the same position
    g = _quiet(ns.SCS_Game, rh.scs_config_path(cfg))
    assert _family(g) == "scs"
    for _ in range(40):
        if g.is_terminal():
            break
        legal = np.flatnonzero(np.asarray(g.possible_actions()).reshape(-1))
        g.step(g.get_action_coords(int(rng.choice(legal))))
    o = oscs.SCS(oscs.load_scenario(os.path.join(golden_io.SCS_CONFIGS, cfg), None))
    assert len(g.action_history) == g.get_length()
    for a in g.action_history:
        o.step(action_index(g, a))
    assert o.get_length() == g.get_length() and o.is_terminal() == g.is_terminal() and o.get_current_player() == g.get_current_player()
    np.testing.assert_array_equal(np.asarray(o.encode()[0], dtype=np.float32).reshape(-1),
                                  np.asarray(g.generate_network_input(), dtype=np.float32).reshape(-1))
    np.testing.assert_array_equal(np.asarray(o.legal_mask()).reshape(-1) != 0, np.asarray(g.possible_actions()).reshape(-1) != 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["match_ttt_0", "match_ttt_1", "match_scs_0"])
def test_reference_mcts_agent_with_this_explorer_on_reference_game_objects(name):
    """The unmodified reference MctsAgent (Testing/Agents/Generic/MctsAgent.py:14-39) playing the reference's own game
    objects, starting from the reference's own Node(0) — only the Explorer inside the agent is this package's.  Moves and root
    visit counts equal the fixture recorded from the all-reference run (oracle/gen_golden_match.py), every ply."""
    import yaml

    from nuzero_b200.search import Explorer
    from nuzero_b200.stubnet import StubNetworkManager
    from oracle.gen_golden_match import load_agents

    rh, ns = _reference()
    z = np.load(os.path.join(golden_io.GOLDEN, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    cfg = yaml.safe_load(open(os.path.join(os.path.dirname(golden_io.GOLDEN), "..", "nuzero_b200", "configs", "a1_search_config.yaml")))
    cfg["Simulation"]["mcts_simulations"] = int(g["sims"])
    MctsAgent, _ = load_agents()
    if str(g["game"]) == "ttt":
        game, shape, game_args = ns.tic_tac_toe(), (1, 3, 3), None
    else:
        seed = int(g["map_seed"])
        seed = None if seed < 0 else seed
        game = _quiet(ns.SCS_Game, rh.scs_config_path(str(g["game"])), seed)
        shape, game_args = tuple(game.get_action_space_shape()), (os.path.join(golden_io.SCS_CONFIGS, str(g["game"])), seed)
    agent = MctsAgent(cfg, StubNetworkManager(shape, salt=int(g["salt"])), 2, None)
    agent.explorer = Explorer(cfg, False, pool_nodes=40000, game_args=game_args)
    mcts_player = int(g["mcts_player"])
    searched = 0
    for ply, a_ref in enumerate(int(a) for a in g["actions"]):
        assert game.get_current_player() == int(g["players"][ply])
        root = agent.root_node
        if game.get_current_player() == mcts_player:
            coords = agent.choose_action(game)
            assert int(np.ravel_multi_index(tuple(int(x) for x in coords), shape)) == a_ref, "ply %d" % ply
        else:
            agent.update_subtree(game, a_ref)  # Tester.py:92-97: an MctsAgent that is not moving follows the move played
        assert int(root.visit_count) == int(g["root_N"][ply]), "root visits at ply %d" % ply
        assert root.to_play == game.get_current_player()
        searched += 1
        game.step(game.get_action_coords(a_ref))
    assert game.is_terminal() and searched == int(g["length"])
