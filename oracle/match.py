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


def policy_choice(probs, mask, u):
    """PolicyAgent.choose_action (PolicyAgent.py:21-68) after the network call, for the uniform u the next np.random.choice
    would consume -> (action, uniforms consumed).  The arg-max of the raw output is played when it is legal; otherwise the
    arg-max of the masked output — after a np.random.choice whose result the reference discards but whose draw advances the
    generator (:52) — or a uniformly random legal action when the network gave the legal actions no mass at all (:58-62)."""
    p = np.asarray(probs).reshape(-1)
    m = np.asarray(mask).reshape(-1)
    raw = int(np.argmax(p))
    if m[raw]:
        return raw, 0
    masked = p * m
    if np.sum(masked) != 0:
        return int(np.argmax(masked)), 1
    return random_choice(m, u), 1


def play_agents(game, nets, cfg, kinds, unif_row, policy_is_prob=True):
    """Tester.py:46-121 for any pairing of "mcts" / "policy" / "random": kinds[0] / nets[0] belong to the player that moves
    first.  Returns dict(actions, players, root_N [plies, 2] (-1 for agents without a tree), terminal_value, length, draws)."""
    assert policy_is_prob, "the restatement works on the stub's probabilities"
    roots = [mcts.Node(0) if k == "mcts" else None for k in kinds]
    first = game.get_current_player()
    out = dict(actions=[], players=[], root_N=[])
    ply = draws = 0
    while not game.is_terminal():
        player = game.get_current_player()
        idx = 0 if player == first else 1
        picks = [None, None]
        for i, k in enumerate(kinds):  # every MCTS agent searches on every ply (choose_action / update_subtree)
            if k == "mcts":
                picks[i] = mcts.run_mcts(cfg, game, nets[i], roots[i], False, policy_is_prob, None, ply)
        if kinds[idx] == "mcts":
            action = picks[idx][0]
        elif kinds[idx] == "policy":
            probs, _ = nets[idx](game.encode())
            action, used = policy_choice(probs, game.legal_mask(), unif_row[draws])
            draws += used
        else:
            action = random_choice(game.legal_mask(), unif_row[draws])
            draws += 1
        out["actions"].append(int(action))
        out["players"].append(int(player))
        out["root_N"].append([-1 if r is None else int(r.N) for r in roots])
        for i, k in enumerate(kinds):
            if k == "mcts":
                roots[i] = picks[i][1] if i == idx else roots[i].child(action)
        game.step(action)
        ply += 1
    out["terminal_value"] = game.get_terminal_value()
    out["length"] = game.get_length()
    out["draws"] = draws
    return out
