"""TEST INFRASTRUCTURE — PUCT search restated from Search/Node.py and Search/Explorer.py.

The arithmetic chains are spelled out with explicit numpy scalar types instead of relying on
NumPy-2 promotion: with float64 priors (TTT, or any noised root) every operation is
binary64; with float32 priors (SCS) the exploration term is rounded to binary32 after each
operation and the binary64 value term is rounded to binary32 before the final add.
"""
import math

import numpy as np

f32 = np.float32
f64 = np.float64


class Node:
    """Search/Node.py:3-32.  Children are kept as parallel lists in ascending action order, which
    is the insertion order of the reference's dict (Explorer.py:177-179)."""

    __slots__ = ("N", "W", "prior", "actions", "kids", "to_play", "terminal_value")

    def __init__(self, prior):
        self.N = 0
        self.W = 0.0
        self.prior = prior
        self.actions = []
        self.kids = []
        self.to_play = -1
        self.terminal_value = None

    def expanded(self):
        return len(self.kids) > 0

    def value(self):
        return 0.0 if self.N == 0 else self.W / self.N

    def child(self, action):
        return self.kids[self.actions.index(action)]


class ReplayTape:
    """Pre-drawn random numbers, one row per move (see oracle/ref_harness.TapeRandom)."""

    def __init__(self, gamma_tape, unif_tape):
        self.gamma_tape, self.unif_tape = gamma_tape, unif_tape

    def gamma(self, cfg, move, n):
        return self.gamma_tape[move, :n]

    def unif(self, move):
        return self.unif_tape[move]


class LiveTape:
    """Draws from numpy's global generator like the reference does (throughput baseline runs)."""

    def gamma(self, cfg, move, n):
        ex = cfg["Exploration"]
        return np.random.gamma(ex["root_dist_alpha"], ex["root_dist_beta"], n)

    def unif(self, move):
        return np.random.random(3)


def exploration_bias(cfg, n_parent):
    """Explorer.py:103-108."""
    base, init = cfg["UCT"]["pb_c_base"], cfg["UCT"]["pb_c_init"]
    return math.log((n_parent + base + 1) / base) + init


def score(cfg, parent, child):
    """Explorer.py:114-130."""
    c = exploration_bias(cfg, parent.N)
    u = math.sqrt(parent.N) / (child.N + 1)
    q = child.value()
    if parent.to_play == 2:  # literal: SCS players are 0/1, so SCS never flips
        q = -q
    q = q * cfg["Exploration"]["value_factor"]
    p = child.prior
    if isinstance(p, np.float32):
        return f32(f32(p * f32(u)) * f32(c)) + f32(q)
    return (float(p) * u) * c + q


GAP_LOG = None  # tests set this to a list: every selection appends (best score - runner-up score), see tests/test_gpu_logits_parity.py


def select_child(cfg, node):
    """Explorer.py:99-101 — python max over (score, action, child): exact ties go to the HIGHEST
    action index."""
    best_i, best_s, second = 0, None, None
    for i, kid in enumerate(node.kids):
        s = score(cfg, node, kid)
        if best_s is None or s >= best_s:  # later (higher action) wins ties
            best_i, best_s, second = i, s, best_s
        elif second is None or s > second:
            second = s
    if GAP_LOG is not None and second is not None:
        GAP_LOG.append(float(best_s) - float(second))
    return node.actions[best_i], node.kids[best_i]


def evaluate(node, game, net, policy_is_prob):
    """Explorer.py:137-181."""
    node.to_play = game.get_current_player()
    if game.is_terminal():
        node.terminal_value = game.get_terminal_value()
        return node.terminal_value
    p, v = net(game.encode())
    p = np.asarray(p, dtype=np.float32).reshape(-1)
    if not policy_is_prob:
        from scipy.special import softmax

        p = softmax(p)
    value = float(v)
    mask = game.legal_mask()
    probs = p * mask  # f32*f64 -> f64 (TTT) ; f32*int8 -> f32 (SCS)
    total = np.sum(probs)
    if total == 0:
        probs = probs + mask
        total = np.sum(probs)
    for a in np.flatnonzero(mask):
        node.actions.append(int(a))
        node.kids.append(Node(probs[a] / total))
    return value


def add_noise(cfg, node, noise):
    """Explorer.py:201-210 — un-normalised gamma draws blended into the priors."""
    frac = cfg["Exploration"]["root_exploration_fraction"]
    for kid, n in zip(node.kids, noise):
        p = kid.prior
        if isinstance(p, np.float32):
            kid.prior = f64(f32(p * f32(1 - frac))) + f64(n) * frac
        else:
            kid.prior = f64(p) * (1 - frac) + f64(n) * frac


def _choice(cdf_probs, u):
    cdf = np.asarray(cdf_probs, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


def softmax_action(node, u):
    """Explorer.py:187-199."""
    from scipy.special import softmax

    probs = np.asarray(softmax([k.N for k in node.kids]), dtype=np.float64)
    probs /= np.sum(probs)
    return node.actions[_choice(probs, u)]


def max_action(node):
    """Explorer.py:183-185 — python max with key: first maximum = LOWEST action on ties."""
    best = 0
    for i, kid in enumerate(node.kids):
        if kid.N > node.kids[best].N:
            best = i
    return node.actions[best]


def select_action(cfg, game, node, training, unif_row):
    """Explorer.py:70-97.  unif_row = (eps_softmax, eps_random, choice_uniform)."""
    if not training:
        return max_action(node)
    ex = cfg["Exploration"]
    if game.get_length() < ex["number_of_softmax_moves"]:
        return softmax_action(node, unif_row[2])
    if unif_row[0] < ex["epsilon_softmax_exploration"]:
        return softmax_action(node, unif_row[2])
    if unif_row[1] < ex["epsilon_random_exploration"]:
        mask = game.legal_mask().reshape(-1)
        return _choice(mask / np.sum(mask), unif_row[2])
    return max_action(node)


def run_mcts(cfg, game, net, root, training, policy_is_prob=True, tape=None, move=0):
    """Explorer.py:40-67 -> (action, chosen child, final root bias)."""
    if training:
        tape = tape if tape is not None else LiveTape()
        add_noise(cfg, root, tape.gamma(cfg, move, len(root.kids)))
    for _ in range(cfg["Simulation"]["mcts_simulations"]):
        node, scratch, path = root, game.clone(), [root]
        while node.expanded():
            action, node = select_child(cfg, node)
            scratch.step(action)
            path.append(node)
        value = evaluate(node, scratch, net, policy_is_prob)
        for n in path:  # Explorer.py:132-135 — no sign flip on the way up
            n.N += 1
            n.W += value
    bias = exploration_bias(cfg, root.N)
    unif_row = tape.unif(move) if training else None
    action = select_action(cfg, game, root, training, unif_row)
    return action, root.child(action), bias


def dump_tree(root, limit=None):
    """Canonical pre-order dump (children ascending by action) for tree-shape comparisons:
    rows of (depth, action, N, n_children) int64 and (W, prior) float64."""
    ints, flts = [], []
    stack = [(root, 0, -1)]
    while stack:
        node, depth, action = stack.pop()
        ints.append((depth, action, node.N, len(node.kids)))
        flts.append((float(node.W), float(node.prior)))
        for a, k in zip(reversed(node.actions), reversed(node.kids)):
            stack.append((k, depth + 1, a))
        if limit is not None and len(ints) >= limit:
            break
    return np.array(ints, dtype=np.int64), np.array(flts, dtype=np.float64)
