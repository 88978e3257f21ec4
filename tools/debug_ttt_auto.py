"""Debug aid: one TTT game in auto mode against the oracle, move by move, for a few launch budgets."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import golden_io
from nuzero_b200 import _ffi
from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
from nuzero_b200.selfplay import game_record, group_games, run_until_idle
from nuzero_b200.stubnet import DyadicStubNet
from oracle import selfplay
from oracle.stubnet_np import stub_forward
from oracle.ttt import TicTacToe

cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
cfg["Simulation"]["mcts_simulations"] = int(sys.argv[1]) if len(sys.argv) > 1 else 100
G = 4
salts = [1, 2, 3, 4]
refs = [selfplay.play_game(TicTacToe(), lambda s, sl=sl: stub_forward(s, 9, sl), cfg, False, True) for sl in salts]
for budget in (1, 2, 8):
    for compact in (True, False):
        e = SearchEngine(tic_tac_toe_spec(), cfg, G, False, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                         auto_advance=True, games_per_slot=1, record_detail=True, pool_nodes=40000, max_sims_per_launch=budget,
                         compact=compact)
        run_until_idle(e, DyadicStubNet(e, salt=salts))
        recs, _ = e.drain_records()
        games = group_games(recs)
        for g in range(G):
            got, ref = game_record(games[g]), refs[g]
            msg = "ok"
            for m in range(min(len(got["actions"]), ref["length"])):
                if not (np.array_equal(got["child_N"][m], ref["child_N"][m]) and np.array_equal(got["child_actions"][m], ref["child_actions"][m])
                        and got["root_N"][m] == ref["root_N"][m]):
                    msg = "move %d: rootN %d/%d acts %s/%s N %s/%s" % (m, got["root_N"][m], ref["root_N"][m], got["child_actions"][m].tolist(),
                                                                      ref["child_actions"][m].tolist(), got["child_N"][m].tolist(), ref["child_N"][m].tolist())
                    break
            print("budget", budget, "compact", compact, "game", g, "actions", got["actions"], "ref", ref["actions"], msg)
