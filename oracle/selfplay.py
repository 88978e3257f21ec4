"""TEST INFRASTRUCTURE — the self-play game loop of Training/Gamer.py:39-97 restated, emitting one
comparable record per move instead of shipping a pickled game to a Ray actor."""
import numpy as np

from . import mcts


def play_game(game, net, cfg, training, policy_is_prob=True, tape=None, keep_states=True,
              tree_dump_moves=()):
    """Returns a dict of per-move arrays (same schema as the golden fixtures)."""
    assert cfg["Simulation"]["keep_subtree"], "only keep_subtree=True is in scope (SURVEY I9)"
    root = mcts.Node(0)
    rec = dict(actions=[], root_N=[], root_W=[], bias=[], n_children=[], child_actions=[],
               child_N=[], child_W=[], child_prior=[], states=[], masks=[], players=[], trees={})
    move = 0
    while not game.is_terminal():
        if keep_states:
            rec["states"].append(game.encode()[0])  # Gamer.py:65-66
            rec["masks"].append(np.packbits(game.legal_mask().reshape(-1) != 0))
        rec["players"].append(game.get_current_player())
        action, child, bias = mcts.run_mcts(cfg, game, net, root, training, policy_is_prob, tape, move)
        rec["actions"].append(action)
        rec["root_N"].append(root.N)  # Gamer.py:71 tree_size
        rec["root_W"].append(float(root.W))
        rec["bias"].append(bias)
        rec["n_children"].append(len(root.kids))  # Gamer.py:72
        rec["child_actions"].append(np.array(root.actions, dtype=np.int32))
        rec["child_N"].append(np.array([k.N for k in root.kids], dtype=np.int64))
        rec["child_W"].append(np.array([float(k.W) for k in root.kids], dtype=np.float64))
        rec["child_prior"].append(np.array([float(k.prior) for k in root.kids], dtype=np.float64))
        if move in tree_dump_moves:
            rec["trees"][move] = mcts.dump_tree(root)
        game.step(action)  # Gamer.py:74-75
        root = child  # Gamer.py:78-79
        move += 1
    rec["terminal_value"] = game.get_terminal_value()
    rec["length"] = game.get_length()
    return rec


def policy_targets(rec, num_actions):
    """store_search_statistics (tic_tac_toe.py:177-182 / SCS_Game.py:1517-1521): visit fractions."""
    out = np.zeros((len(rec["actions"]), num_actions), dtype=np.float64)
    for m, (acts, n) in enumerate(zip(rec["child_actions"], rec["child_N"])):
        out[m, acts] = n / n.sum()
    return out


def stats(rec):
    """Gamer.py:42-50,81-92."""
    L = rec["length"]
    children = tree = 0
    bias = 0
    for k, n, b in zip(rec["n_children"], rec["root_N"], rec["bias"]):  # `+=` per move, as the reference accumulates
        children += k
        tree += n
        bias += b
    return {
        "number_of_moves": L,
        "average_children": children / L,
        "average_tree_size": tree / L,
        "final_tree_size": rec["root_N"][-1],
        "average_bias_value": bias / L,
        "final_bias_value": rec["bias"][-1],
    }
