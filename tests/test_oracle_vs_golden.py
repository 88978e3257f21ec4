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
