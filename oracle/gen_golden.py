"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(imported read-only from /root/reference through oracle/ref_harness.py) in the build container.

    python -m oracle.gen_golden            # regenerate everything
    python -m oracle.gen_golden ttt        # only fixtures whose name starts with "ttt"

The loop mirrors Training/Gamer.py:64-79 call for call (store state -> Explorer.run_mcts ->
game.step -> re-root) but records the root statistics after every move.  The reference `Gamer.play_game` itself
(unmodified class, ray stub, the reference ReplayBuffer as sink) is run by oracle/gen_golden_gamer.py, which pins `stats`
and the replay tuples (tests/golden/gamer_*.npz).
"""
import copy
import os
import sys

import numpy as np
import yaml

from . import ref_harness as rh
from .stubnet_np import StubNetwork

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
KMAX = 512  # widest gamma row on a tape


def search_config(sims, **over):
    with open(os.path.join(rh.REFERENCE_ROOT, "Configs", "Search", "a1_search_config.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["Simulation"]["mcts_simulations"] = sims
    for k, v in over.items():
        for sec in cfg.values():
            if k in sec:
                sec[k] = v
    return cfg


def make_tape(seed, cfg, moves, kmax=KMAX):
    rng = np.random.Generator(np.random.Philox(seed))
    ex = cfg["Exploration"]
    gamma = rng.gamma(ex["root_dist_alpha"], ex["root_dist_beta"], size=(moves, kmax))
    unif = rng.random(size=(moves, 3))
    return gamma, unif


def dump_ref_tree(root):
    ints, flts = [], []
    stack = [(root, 0, -1)]
    while stack:
        node, depth, action = stack.pop()
        ints.append((depth, action, node.visit_count, len(node.children)))
        flts.append((float(node.value_sum), float(node.prior)))
        for a in sorted(node.children, reverse=True):
            stack.append((node.children[a], depth + 1, a))
    return np.array(ints, dtype=np.int64), np.array(flts, dtype=np.float64)


def play_reference(game, cfg, training, salt, tape_arrays, tree_dump_moves=(), max_moves=400):
    ns = rh.load()
    net = StubNetwork(game.get_action_space_shape(), salt)
    ex = ns.Explorer(cfg, training)
    tape = rh.TapeRandom(*tape_arrays) if training else None
    root = ns.Node(0)
    rec = dict(actions=[], root_N=[], root_W=[], bias=[], n_children=[], child_actions=[],
               child_N=[], child_W=[], child_prior=[], states=[], masks=[], players=[], trees={})
    with rh.parity_patches(tape=tape, identity_softmax=True):
        move = 0
        while not game.is_terminal():
            state = game.generate_network_input()
            game.store_state(state)
            rec["states"].append(state[0].numpy().copy())
            rec["masks"].append(np.packbits(np.asarray(game.possible_actions()).reshape(-1) != 0))
            rec["players"].append(int(game.get_current_player()))
            action, child, bias = ex.run_mcts(game, net, root, 2, None)
            acts = list(root.children.keys())
            assert acts == sorted(acts)
            kids = [root.children[a] for a in acts]
            rec["actions"].append(int(action))
            rec["root_N"].append(int(root.visit_count))
            rec["root_W"].append(float(root.value_sum))
            rec["bias"].append(float(bias))
            rec["n_children"].append(len(acts))
            rec["child_actions"].append(np.array(acts, dtype=np.int32))
            rec["child_N"].append(np.array([k.visit_count for k in kids], dtype=np.int64))
            rec["child_W"].append(np.array([float(k.value_sum) for k in kids], dtype=np.float64))
            rec["child_prior"].append(np.array([float(k.prior) for k in kids], dtype=np.float64))
            if move in tree_dump_moves:
                rec["trees"][move] = dump_ref_tree(root)
            game.step(game.get_action_coords(action))
            game.store_search_statistics(root)
            root = child
            move += 1
            assert move < max_moves
    rec["terminal_value"] = int(game.get_terminal_value())
    rec["length"] = int(game.get_length())
    rec["winner"] = int(game.get_winner())
    rec["child_policy"] = np.array(game.child_policy, dtype=np.float64)
    rec["net_calls"] = net.calls
    return rec


def pack(rec, extra):
    """Ragged per-move lists -> flat arrays + offsets, ready for np.savez_compressed."""
    out = dict(extra)
    off = np.concatenate([[0], np.cumsum(rec["n_children"])]).astype(np.int64)
    out.update(
        actions=np.array(rec["actions"], dtype=np.int32),
        root_N=np.array(rec["root_N"], dtype=np.int64),
        root_W=np.array(rec["root_W"], dtype=np.float64),
        bias=np.array(rec["bias"], dtype=np.float64),
        players=np.array(rec["players"], dtype=np.int8),
        child_off=off,
        child_actions=np.concatenate(rec["child_actions"]),
        child_N=np.concatenate(rec["child_N"]),
        child_W=np.concatenate(rec["child_W"]),
        child_prior=np.concatenate(rec["child_prior"]),
        states=np.stack(rec["states"]).astype(np.float32),
        masks=np.stack(rec["masks"]),
        terminal_value=np.int64(rec["terminal_value"]),
        length=np.int64(rec["length"]),
        winner=np.int64(rec.get("winner", 0)),
        child_policy=rec["child_policy"],
        net_calls=np.int64(rec.get("net_calls", 0)),
    )
    for m, (ti, tf) in rec["trees"].items():
        out["tree%d_i" % m] = ti
        out["tree%d_f" % m] = tf
    out["tree_moves"] = np.array(sorted(rec["trees"]), dtype=np.int64)
    return out


def save(name, rec, cfg, training, salt, tape_arrays, game_desc, extra=None):
    extra = dict(extra or {})
    extra.update(
        cfg_yaml=np.array(yaml.safe_dump(cfg)),
        training=np.int64(training),
        salt=np.int64(salt),
        game=np.array(game_desc),
    )
    if training:
        L = len(rec["actions"])
        kmax = max(1, max(rec["n_children"]))
        extra["gamma_tape"] = tape_arrays[0][:L, :kmax]
        extra["unif_tape"] = tape_arrays[1][:L]
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **pack(rec, extra))
    print("%-40s moves=%3d tv=%+d bytes=%d" % (name, rec["length"], rec["terminal_value"],
                                              os.path.getsize(path)))


# ------------------------------------------------------------------------------------------------
def gen_ttt(only):
    ns = rh.load()
    cases = [
        ("ttt_p0_s25_salt0", 25, False, 0, {}),
        ("ttt_p0_s100_salt1", 100, False, 1, {}),
        ("ttt_p0_s100_salt2_vf", 100, False, 2, dict(value_factor=0.75, pb_c_init=1.25, pb_c_base=500)),
        ("ttt_p0_s800_salt3", 800, False, 3, {}),
        ("ttt_p1_s50_seed1", 50, True, 4, {}),
        ("ttt_p1_s100_seed2_soft3", 100, True, 5, dict(number_of_softmax_moves=3)),
        ("ttt_p1_s100_seed3_eps", 100, True, 6, dict(epsilon_softmax_exploration=0.35,
                                                   epsilon_random_exploration=0.5)),
        ("ttt_p1_s800_seed4", 800, True, 7, dict(root_exploration_fraction=0.25, root_dist_alpha=0.3)),
    ]
    for name, sims, training, salt, over in cases:
        if only and not name.startswith(only):
            continue
        cfg = search_config(sims, **over)
        tape = make_tape(1000 + salt, cfg, 16, 16)
        rec = play_reference(ns.tic_tac_toe(), cfg, training, salt, tape, tree_dump_moves=(0, 2, 5))
        save(name, rec, cfg, training, salt, tape, "ttt")


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if not rh.available():
        raise SystemExit("reference tree not found; goldens can only be generated in the build container")
    gen_ttt(only)
    from . import gen_golden_scs

    gen_golden_scs.main(only)


if __name__ == "__main__":
    main()
