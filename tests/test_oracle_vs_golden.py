"""CPU: the oracle restatement replays every golden fixture produced by the real reference
(oracle/gen_golden.py) and must reproduce it bit for bit."""
import numpy as np
import pytest

from oracle import mcts, selfplay
from oracle.stubnet_np import stub_forward
from oracle.ttt import TicTacToe

import golden_io


def _net(n_actions, salt):
    return lambda state: stub_forward(state, n_actions, salt)


@pytest.mark.parametrize("name", golden_io.names("ttt_"))
def test_ttt_oracle_matches_reference(name):
    g = golden_io.load(name)
    tape = mcts.ReplayTape(g["gamma_tape"], g["unif_tape"]) if g["training"] else None
    rec = selfplay.play_game(TicTacToe(), _net(9, g["salt"]), g["cfg"], g["training"], True, tape,
                             tree_dump_moves=tuple(g["tree_moves"].tolist()))
    golden_io.assert_record_matches(rec, g)
    pol = selfplay.policy_targets(rec, 9)
    np.testing.assert_allclose(pol, g["child_policy"], rtol=1e-15, atol=0)


# ---- SCS -----------------------------------------------------------------------------------------
import os

from oracle import scs as oscs

SCS_CFG = golden_io.SCS_CONFIGS


def _scenario_for(g):
    _, cfg_name, seed = str(g["game"]).split(":") if str(g["game"]).count(":") == 2 else (None, str(g["game"])[4:], int(g["seed"]))
    sc = oscs.load_scenario(os.path.join(SCS_CFG, cfg_name), int(seed) or None)
    return sc


@pytest.mark.parametrize("name", golden_io.names("scsenv_"))
def test_scs_env_oracle_matches_reference_playout(name):
    z = np.load(os.path.join(golden_io.GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    sc = _scenario_for(g)
    # the scenario loader reproduces the reference's seeded map / victory points
    np.testing.assert_array_equal(np.array(sc.tile_terrain).reshape(sc.rows, sc.cols), g["sc_terrain"])
    np.testing.assert_array_equal(np.array(sc.terrain_types, dtype=np.float64), g["sc_terrain_types"])
    np.testing.assert_array_equal(np.array(sc.vp[0]).reshape(-1, 2), g["sc_vp0"])
    np.testing.assert_array_equal(np.array(sc.vp[1]).reshape(-1, 2), g["sc_vp1"])
    game = oscs.SCS(sc)
    assert game.action_space_shape == tuple(g["action_shape"]) and game.state_shape == tuple(g["state_shape"])
    for i, a in enumerate(g["actions"]):
        assert not game.is_terminal()
        np.testing.assert_array_equal(np.packbits(game.legal_mask() != 0), g["masks"][i], err_msg="mask @%d" % i)
        np.testing.assert_array_equal(game.encode()[0], g["states"][i], err_msg="state @%d" % i)
        assert (game.get_current_player(), game.stage, game.turn, game.length) == \
            (g["players"][i], g["stages"][i], g["turns"][i], g["lengths"][i]), "step %d" % i
        game.step(int(a), check=True)
    assert game.is_terminal() == bool(g["terminal"])
    assert game.get_terminal_value() == int(g["terminal_value"]) and game.get_winner() == int(g["winner"])
    assert game.length == int(g["final_length"]) and game.stage == int(g["final_stage"])
    assert game.get_current_player() == int(g["final_player"])
    np.testing.assert_array_equal(game.encode()[0], g["final_state"])


@pytest.mark.parametrize("name", golden_io.names("scs_p"))
def test_scs_oracle_matches_reference_selfplay(name):
    g = golden_io.load(name)
    sc = _scenario_for(g)
    tape = mcts.ReplayTape(g["gamma_tape"], g["unif_tape"]) if g["training"] else None
    rec = selfplay.play_game(oscs.SCS(sc), _net(sc.A, g["salt"]), g["cfg"], g["training"], True, tape,
                             tree_dump_moves=tuple(g["tree_moves"].tolist()))
    golden_io.assert_record_matches(rec, g)
    pol = selfplay.policy_targets(rec, sc.A)
    np.testing.assert_allclose(pol, g["child_policy"], rtol=1e-15, atol=0)
