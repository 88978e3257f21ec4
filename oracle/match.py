"""TEST INFRASTRUCTURE — one evaluation game of Testing/Tester.py:46-121 restated for the pairing the reference's
TestManager runs most (MctsAgent against RandomAgent): the MCTS agent keeps its sub-tree, so it searches on EVERY ply —
`choose_action` on its own turn (MctsAgent.py:28-33), `update_subtree` on the opponent's (MctsAgent.py:35-39, called from
Tester.py:95-97) — and then follows the action that was actually played.  The random agent's `np.random.choice(num_actions,
p=mask/sum(mask))` (RandomAgent.py:10-15) is replayed from a tape of uniforms with numpy's own inverse-CDF rule."""
import numpy as np

from . import mcts


def random_choice(mask, u):
    """np.random.choice(len(mask), p=mask/sum(mask)) for the uniform draw u (numpy: cdf = p.cumsum(); cdf /= cdf[-1];
    cdf.searchsorted(u, side='right'))."""
    m = np.asarray(mask, dtype=np.float64).reshape(-1)
    p = m / m.sum()
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


def play_match(game, net, cfg, mcts_player, unif_row, policy_is_prob=True):
    """Returns dict(actions, root_N (after each ply's search), winner, length)."""
    root = mcts.Node(0)
    out = dict(actions=[], root_N=[], players=[])
    ply = draws = 0
    while not game.is_terminal():
        player = game.get_current_player()
        action_m, child, _ = mcts.run_mcts(cfg, game, net, root, False, policy_is_prob, None, ply)
        if player == mcts_player:
            action, nxt = action_m, child
        else:
            action = random_choice(game.legal_mask(), unif_row[draws])  # one uniform per RandomAgent.choose_action call
            draws += 1
            nxt = root.child(action)
        out["actions"].append(int(action))
        out["root_N"].append(int(root.N))
        out["players"].append(int(player))
        game.step(action)
        root = nxt
        ply += 1
    out["terminal_value"] = game.get_terminal_value()
    out["length"] = game.get_length()
    return out
