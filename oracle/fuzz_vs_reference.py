"""TEST INFRASTRUCTURE — live differential run of the oracle's SCS rules against the UNMODIFIED reference SCS_Game over
EVERY scenario the reference ships (Games/SCS/Game_configs/*.yml), beyond the eleven the committed fixtures cover: random
playouts in the reference (oracle/gen_golden_scs.random_playout), replayed step by step in oracle/scs.py — seeded map and
victory points, legal mask, encoded planes, player / stage / turn / length before every action, result at the end.

    python -m oracle.fuzz_vs_reference [--seeds 2] [--out oracle/fuzz_report.json]

Needs /root/reference (build container only); writes a small report that DESIGN.md quotes.  Nothing is imported from here
by the product or by the tests.
"""
import argparse
import contextlib
import glob
import io
import json
import os
import time
import traceback

import numpy as np

from . import ref_harness as rh
from . import scs as oscs
from .gen_golden_scs import random_playout


def compare(path, seed, g):
    sc = oscs.load_scenario(path, seed=seed)
    np.testing.assert_array_equal(np.array(sc.tile_terrain).reshape(sc.rows, sc.cols), g["sc_terrain"])
    np.testing.assert_array_equal(np.array(sc.vp[0]).reshape(-1, 2), g["sc_vp0"])
    np.testing.assert_array_equal(np.array(sc.vp[1]).reshape(-1, 2), g["sc_vp1"])
    game = oscs.SCS(sc)
    assert game.action_space_shape == tuple(g["action_shape"]) and game.state_shape == tuple(g["state_shape"])
    for i, a in enumerate(g["actions"]):
        assert not game.is_terminal(), "oracle terminal early @%d" % i
        np.testing.assert_array_equal(np.packbits(game.legal_mask() != 0), g["masks"][i], err_msg="mask @%d" % i)
        np.testing.assert_array_equal(game.encode()[0], g["states"][i], err_msg="state @%d" % i)
        assert (game.get_current_player(), game.stage, game.turn, game.length) == \
            (g["players"][i], g["stages"][i], g["turns"][i], g["lengths"][i]), "bookkeeping @%d" % i
        game.step(int(a), check=True)
    assert game.is_terminal() == bool(g["terminal"])
    assert game.get_terminal_value() == int(g["terminal_value"]) and game.get_winner() == int(g["winner"])
    np.testing.assert_array_equal(game.encode()[0], g["final_state"])
    return len(g["actions"])


def selfplay_fuzz(n_ttt, n_scs):
    """Whole self-play games through the reference's Explorer / Gamer loop (oracle/gen_golden.play_reference: dyadic stub
    network, RNG tape when training) against oracle/mcts.py + oracle/selfplay.py: trajectories, root and child visit counts,
    value sums, priors (noised ones included), root bias, states and masks, bit for bit; random search-config overrides."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import golden_io

    from . import mcts, selfplay
    from .gen_golden import make_tape, pack, play_reference, search_config
    from .stubnet_np import stub_forward
    from .ttt import TicTacToe

    ns = rh.load()
    rng = np.random.default_rng(2026)
    scs_names = ["solo_soldier_config_5.yml", "mirrored_config_5.yml", "unbalanced_config_5.yml", "randomized_config_5.yml",
                 "r_unbalanced_config_5.yml", "solo_soldier_config_8.yml"]
    cfg_dir = os.path.dirname(rh.scs_config_path("mirrored_config_5.yml"))
    out = {"games": 0, "moves": 0, "mismatch": []}
    cases = [("ttt", None)] * n_ttt + [(scs_names[i % len(scs_names)], 50 + i) for i in range(n_scs)]
    for i, (name, seed) in enumerate(cases):
        training = bool(rng.integers(0, 2))
        over = {}
        if rng.random() < 0.5:
            over.update(pb_c_init=float(rng.choice([1.0, 1.15, 1.25, 2.0])), pb_c_base=int(rng.choice([500, 10000, 19652])))
        if rng.random() < 0.3:
            over.update(value_factor=float(rng.choice([0.5, 0.75, 1.0])))
        if training and rng.random() < 0.5:
            over.update(number_of_softmax_moves=int(rng.integers(0, 6)), epsilon_softmax_exploration=float(rng.choice([0.0, 0.2, 0.5])),
                        epsilon_random_exploration=float(rng.choice([0.0, 0.1, 0.4])))
        sims = int(rng.integers(20, 160)) if name == "ttt" else int(rng.integers(8, 24))
        salt = int(rng.integers(0, 1000))
        cfg = search_config(sims, **over)
        try:
            if name == "ttt":
                tape = make_tape(3000 + i, cfg, 16, 16)
                ref_game, ora_game, A = ns.tic_tac_toe(), TicTacToe(), 9
            else:
                tape = make_tape(3000 + i, cfg, 400, 128)
                ref_game = rh.make_scs(name, seed)
                sc = oscs.load_scenario(os.path.join(cfg_dir, name), seed=seed)
                ora_game, A = oscs.SCS(sc), sc.A
            with contextlib.redirect_stdout(io.StringIO()):
                rec_ref = play_reference(ref_game, cfg, training, salt, tape, tree_dump_moves=(0, 3))
            g = pack(rec_ref, {})
            ora_tape = mcts.ReplayTape(tape[0], tape[1]) if training else None
            rec = selfplay.play_game(ora_game, lambda s_, sl=salt, A_=A: stub_forward(s_, A_, sl), cfg, training, True, ora_tape,
                                     tree_dump_moves=tuple(g["tree_moves"].tolist()))
            golden_io.assert_record_matches(rec, g)
            out["games"] += 1
            out["moves"] += int(g["length"])
        except Exception as ex:
            out["mismatch"].append({"case": i, "game": name, "seed": seed, "sims": sims, "training": training, "over": over,
                                    "error": (str(ex).strip().splitlines() or [type(ex).__name__])[0][:200]})
        print("self-play %3d %-28s sims %3d training %d %s" % (i, name, sims, training, "ok" if not out["mismatch"] or out["mismatch"][-1]["case"] != i else "MISMATCH"), flush=True)
    return out


def agents_fuzz(n):
    """Evaluation matches between the reference's own MctsAgent / PolicyAgent / RandomAgent (oracle/gen_golden_match.play_agents)
    against oracle/match.play_agents: every pairing, both seats, Tic-Tac-Toe and SCS."""
    from . import match
    from .gen_golden import search_config
    from .gen_golden_match import play_agents
    from .stubnet_np import stub_forward
    from .ttt import TicTacToe

    ns = rh.load()
    rng = np.random.default_rng(7)
    kinds_all = [(a, b) for a in ("mcts", "policy", "random") for b in ("mcts", "policy", "random")]
    scs_names = ["solo_soldier_config_5.yml", "mirrored_config_5.yml", "unbalanced_config_5.yml", "randomized_config_5.yml"]
    cfg_dir = os.path.dirname(rh.scs_config_path("mirrored_config_5.yml"))
    out = {"matches": 0, "plies": 0, "mismatch": []}
    for i in range(n):
        kinds = kinds_all[i % len(kinds_all)]
        salts = (int(rng.integers(0, 500)), int(rng.integers(0, 500)))
        seed = 500 + i
        scs = i % 3 == 2
        cfg = search_config(int(rng.integers(6, 14)) if scs else int(rng.integers(10, 60)))
        try:
            if scs:
                name, map_seed = scs_names[(i // 3) % len(scs_names)], 70 + i
                ref_game = rh.make_scs(name, map_seed)
                ora_game = oscs.SCS(oscs.load_scenario(os.path.join(cfg_dir, name), seed=map_seed))
            else:
                name, ref_game, ora_game = "ttt", ns.tic_tac_toe(), TicTacToe()
            A = ora_game.get_num_actions()
            g = play_agents(ref_game, cfg, kinds, salts, seed)
            nets = [lambda s_, sl=sl, A_=A: stub_forward(s_, A_, sl) for sl in salts]
            got = match.play_agents(ora_game, nets, cfg, kinds, g["unif_tape"])
            assert got["actions"] == g["actions"].tolist() and got["players"] == g["players"].tolist()
            assert got["root_N"] == g["root_N"].tolist() and got["draws"] == int(g["draws"])
            assert got["terminal_value"] == int(g["terminal_value"]) and got["length"] == int(g["length"])
            out["matches"] += 1
            out["plies"] += len(g["actions"])
        except Exception as ex:
            out["mismatch"].append({"case": i, "game": name, "kinds": list(kinds),
                                    "error": (str(ex).strip().splitlines() or [type(ex).__name__])[0][:200]})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=27)
    ap.add_argument("--selfplay-ttt", type=int, default=24)
    ap.add_argument("--selfplay-scs", type=int, default=12)
    ap.add_argument("--seeds", type=int, default=2)
    ap.add_argument("--max-steps", type=int, default=400)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz_report.json"))
    a = ap.parse_args()
    rh.load()
    cfg_dir = os.path.dirname(rh.scs_config_path("mirrored_config_5.yml"))
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(cfg_dir, "*.yml")))
    report = {"scenarios": {}, "seeds_per_scenario": a.seeds, "max_steps": a.max_steps}
    t0 = time.time()
    for name in names:
        entry = {"playouts": 0, "steps": 0, "status": "ok"}
        for k in range(a.seeds):
            seed, play_seed = 100 + k, 1000 + 17 * k
            try:
                g = random_playout(name, seed, play_seed, max_steps=a.max_steps)
            except Exception as ex:  # the reference itself cannot load / play this file
                entry["status"] = "reference failed: %s" % (str(ex).splitlines()[0][:120] if str(ex) else type(ex).__name__)
                break
            try:
                entry["steps"] += compare(os.path.join(cfg_dir, name), seed, g)
                entry["playouts"] += 1
            except Exception as ex:
                entry["status"] = "MISMATCH: %s" % (str(ex).strip().splitlines()[0][:200] if str(ex) else type(ex).__name__)
                entry["trace"] = traceback.format_exc().splitlines()[-3:]
                break
        report["scenarios"][name] = entry
        print("%-44s %s  (%d playouts, %d steps)" % (name, entry["status"], entry["playouts"], entry["steps"]), flush=True)
    ok = [n for n, e in report["scenarios"].items() if e["status"] == "ok"]
    bad = [n for n, e in report["scenarios"].items() if e["status"].startswith("MISMATCH")]
    report["summary"] = {"scenarios": len(names), "ok": len(ok), "mismatch": len(bad),
                         "reference_failed": len(names) - len(ok) - len(bad),
                         "steps_compared": sum(e["steps"] for e in report["scenarios"].values()),
                         "seconds": round(time.time() - t0, 1)}
    if a.selfplay_ttt + a.selfplay_scs > 0:
        report["selfplay"] = selfplay_fuzz(a.selfplay_ttt, a.selfplay_scs)
        report["summary"]["selfplay_games"] = report["selfplay"]["games"]
        report["summary"]["selfplay_moves"] = report["selfplay"]["moves"]
        report["summary"]["selfplay_mismatch"] = len(report["selfplay"]["mismatch"])
        report["summary"]["seconds"] = round(time.time() - t0, 1)
    if a.agents > 0:
        report["agents"] = agents_fuzz(a.agents)
        report["summary"]["agent_matches"] = report["agents"]["matches"]
        report["summary"]["agent_plies"] = report["agents"]["plies"]
        report["summary"]["agent_mismatch"] = len(report["agents"]["mismatch"])
        report["summary"]["seconds"] = round(time.time() - t0, 1)
    with open(a.out, "w") as f:
        json.dump(report, f, indent=1)
    print(report["summary"])


if __name__ == "__main__":
    main()
