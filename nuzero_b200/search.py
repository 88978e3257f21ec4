"""Drop-in for the reference's Search/Node.py and Search/Explorer.py: the same constructor and
`run_mcts(game, network, root_node, recurrent_iterations=2, cache=None)` contract, with the tree in
device memory and every simulation executed by the CUDA engine (manual mode, one slot).

This is the one-game compatibility surface (the reference runs one simulation of one game at a
time, and so does this — one kernel launch + one `network.inference` call per leaf).  Throughput
comes from nuzero_b200.gamer.Gamer.play_games, which drives thousands of slots with the same kernels.
"""
import math

import torch

from . import _ffi
from .engine import SearchEngine


class Node:
    """Search/Node.py:3-32.  `Node(prior)` is an unbound root; the nodes handed back by
    `Explorer.run_mcts` are views of device nodes."""

    def __init__(self, prior):
        self._prior = prior
        self._eng = None
        self._idx = -1
        self._epoch = -1
        self._to_play = -1
        self._parent = None
        self.terminal_value = None

    def _bind(self, eng, idx, epoch, parent=None):
        self._eng, self._idx, self._epoch, self._parent = eng, int(idx), epoch, parent
        return self

    # Search/Node.py:12 + Explorer.py:138: `to_play` is -1 until the node is evaluated, then the player to move there.  The
    # device record does not hold it (the kernel takes it from the game state); a view derives it on demand by stepping the
    # root state along its path with the environment kernels.
    @property
    def to_play(self):
        if self._to_play != -1 or not self._live() or self._idx == 0 or self._parent is None or self.visit_count == 0:
            return self._to_play
        st = self._state_words()
        from .engine import EnvOps

        return int(EnvOps(self._eng).status(st[None], self._maps())[0, 2])

    @to_play.setter
    def to_play(self, v):
        self._to_play = v

    def _maps(self):
        return [int(self._eng.ctl[0, _ffi.CTL_MAP])]

    def _state_words(self):
        """Compact game state at this node: the root's state stepped along the path of actions."""
        if self._idx == 0 or self._parent is None:
            return self._eng.gstate[0, 0].clone()
        from .engine import EnvOps

        st = self._parent._state_words()
        EnvOps(self._eng).step(st[None], [self._link()[2]], self._maps())
        return st

    def _live(self):
        return self._eng is not None and self._epoch == getattr(self._eng, "_epoch", None)

    @property
    def visit_count(self):
        return int(self._eng.node_N[0, self._idx]) if self._live() else 0

    @property
    def value_sum(self):
        return float(self._eng.node_W[0, self._idx]) if self._live() else 0

    @property
    def prior(self):
        if not self._live() or self._idx == 0:
            return self._prior
        return float(self._eng.node_prior[0, self._idx])

    def _link(self):
        """(first child, number of children, action that leads here)"""
        e, i = self._eng, self._idx
        return int(e.node_base[0, i]) & 0xFFFFFFFF, int(e.node_K[0, i]) & 0xFFFF, (int(e.node_flags[0, i]) >> 16) & 0xFFFF

    @property
    def children(self):
        if not self._live():
            return {}
        base, k, _ = self._link()
        if k == 0:
            return {}
        acts = self._eng.node_action(0, slice(base, base + k)).tolist()
        return {a: Node(0)._bind(self._eng, base + i, self._epoch, self) for i, a in enumerate(acts)}

    def is_terminal(self):
        return self.terminal_value is not None

    def expanded(self):
        return self.num_children() > 0

    def value(self):
        n = self.visit_count
        return 0.0 if n == 0 else self.value_sum / n

    def num_children(self):
        return self._link()[1] if self._live() else 0

    def get_visit_count(self):
        return self.visit_count

    def get_child(self, action):
        return self.children[action]


class Explorer:
    """Search/Explorer.py:35-67, 212-214."""

    def __init__(self, search_config, training, device="cuda:0", pool_nodes=None, rng_tape=None, seed=0, game_args=None):
        """game_args: (game_config_path, seed) of the SCS scenario when `run_mcts` is handed the REFERENCE's SCS_Game objects
        (they do not remember how they were built; nuzero_b200.games.adopt)."""
        self._game_args = game_args
        self.config = search_config
        self.training = training
        self.device = device
        self._pool_nodes = pool_nodes
        self._tape = rng_tape  # (gamma [M, K], uniforms [M, 3]) for parity runs; None -> device Philox
        self._seed = seed
        self._engines = {}

    def set_search_config(self, search_config):
        self.config = search_config
        self._engines = {}

    def _engine_for(self, game, network):
        spec = game.spec()
        is_prob = bool(getattr(network, "outputs_probabilities", False))
        key = (spec.name, is_prob)
        if key not in self._engines:
            tm, tw = (0, 0) if self._tape is None else (self._tape[0].shape[0], self._tape[0].shape[1])
            sims = int(self.config["Simulation"]["mcts_simulations"])
            pool = self._pool_nodes or (2 + sims * (spec.max_moves + 1)) * min(spec.max_children, 64)
            eng = SearchEngine(spec, self.config, 1, self.training, device=self.device, pool_nodes=pool,
                               policy_is_prob=is_prob, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                               auto_advance=False, max_sims_per_launch=1 << 20, tape_moves=tm, tape_width=tw,
                               seed=self._seed, arena_words=1 << 16)
            if self._tape is not None:
                eng.set_tapes(self._tape[0][None], self._tape[1][None])
            eng._epoch = 0
            self._engines[key] = eng
        return self._engines[key]

    def run_mcts(self, game, network, root_node, recurrent_iterations=2, cache=None):
        from .games.adopt import to_device_game

        game = to_device_game(game, self._game_args, self.device)  # a reference tic_tac_toe / SCS_Game is mirrored on the device
        eng = self._engine_for(game, network)
        ctl = eng.ctl
        bound = isinstance(root_node, Node) and root_node._eng is eng and root_node._epoch == eng._epoch
        phase = int(ctl[0, _ffi.CTL_PHASE])
        want = game.compact_state().to(eng.device)
        fresh = True
        if bound and phase == _ffi.PHASE_MOVE_READY and root_node._idx != 0:
            # the caller re-rooted on a child (Training/Gamer.py:78-79, MctsAgent.py:30-31): commit that child's action
            _, _, action = root_node._link()
            eng.commit_moves([action])
            fresh = int(ctl[0, _ffi.CTL_PHASE]) != _ffi.PHASE_READY or not torch.equal(eng.gstate[0, 0], want)
        elif bound and phase == _ffi.PHASE_MOVE_READY and torch.equal(eng.gstate[0, 0], want):
            # same root searched again (MctsAgent.update_subtree, MctsAgent.py:35-39)
            eng.research_root(0)
            fresh = False
        if fresh:
            # Node(0), or a node that is not the position of `game` (the caller skipped update_subtree on the opponent's
            # move): a new tree rooted at the position of `game`.  The reference would search on with a tree that belongs
            # to another position; rebuilding is the only faithful thing a kept sub-tree can fall back to.
            if hasattr(game, "_map") and game._map is not None:
                eng.set_maps([game._map])
            eng.reset()
            eng._epoch += 1
            eng.gstate[0, 0].copy_(want)
            ctl[0, _ffi.CTL_MOVE] = int(game.get_length())
        while True:
            eng.advance()
            phase = int(ctl[0, _ffi.CTL_PHASE])
            if phase == _ffi.PHASE_LEAF_PENDING:
                p, v = network.inference(eng.leaf[0:1], False, recurrent_iterations)  # Explorer.py:151/158
                eng.policy[0].copy_(torch.as_tensor(p).reshape(-1))
                eng.value[0] = float(torch.as_tensor(v).reshape(-1)[0])
            elif phase == _ffi.PHASE_MOVE_READY:
                break
            else:
                eng.raise_on_error()
                raise Exception("Explorer.run_mcts: unexpected engine phase %d" % phase)
        root_idx = int(ctl[0, _ffi.CTL_ROOT])
        if isinstance(root_node, Node):
            root_node._bind(eng, root_idx, eng._epoch)
            root_node.to_play = game.get_current_player()
            root_view = root_node
        else:
            # the reference's own Search/Node.Node(0) as the root container (its MctsAgent starts with one): give it the fields
            # the callers read — children (views of the device nodes), visit_count, value_sum, to_play
            root_view = Node(0)._bind(eng, root_idx, eng._epoch)
            root_view.to_play = game.get_current_player()
            if hasattr(root_node, "children"):
                root_node.children = root_view.children
                root_node.visit_count, root_node.value_sum = root_view.visit_count, root_view.value_sum
                root_node.to_play = root_view.to_play
        base, k, _ = root_view._link()
        child_i = int(ctl[0, _ffi.CTL_CHOSEN])
        action = int(eng.node_action(0, base + child_i))
        n_root = int(eng.node_N[0, root_idx])
        base_c, init_c = self.config["UCT"]["pb_c_base"], self.config["UCT"]["pb_c_init"]
        bias = math.log((n_root + base_c + 1) / base_c) + init_c  # calculate_exploration_bias (:103-108)
        return action, Node(0)._bind(eng, base + child_i, eng._epoch, root_view), bias
