"""Host mirrors of the reference's game classes.  Each object owns one compact game state in device
memory; every rule evaluation (step, legal mask, plane encoding, terminal scoring) is a kernel call
through the C ABI (nz_env_*).  The method set is the duck type the reference's search, worker and
replay buffer use (Games/Game.py:3-106; Search/Explorer.py:51-57,138-165; Training/Gamer.py:64-90;
Training/ReplayBuffer.py:31-33).  These objects are the one-game compatibility surface — thousands
of games at once go through nuzero_b200.gamer.Gamer.play_games / SearchEngine instead.
"""
import numpy as np
import torch

from .. import _ffi
from ..engine import EnvOps, SearchEngine, tic_tac_toe_spec

_DEFAULT_CFG = {
    "Simulation": {"mcts_simulations": 1, "keep_subtree": True},
    "UCT": {"pb_c_base": 10000, "pb_c_init": 1.15},
    "Exploration": {"number_of_softmax_moves": 0, "epsilon_softmax_exploration": 0.0,
                    "epsilon_random_exploration": 0.0, "value_factor": 1, "root_exploration_distribution": "gamma",
                    "root_exploration_fraction": 0.2, "root_dist_alpha": 0.15, "root_dist_beta": 1},
}
_env_cache = {}


def env_for(key, spec_fn, device="cuda:0"):
    """One tiny engine per game family / scenario, used only for its environment kernels."""
    k = (key, str(device))
    if k not in _env_cache:
        eng = SearchEngine(spec_fn(), _DEFAULT_CFG, 1, False, device=device, pool_nodes=8, max_depth=4,
                           ctable_len=4, arena_words=64)
        _env_cache[k] = EnvOps(eng)
    return _env_cache[k]


class DeviceGame:
    """Shared implementation; subclasses set `_env`, `_map`, shapes and names."""

    def _init_common(self):
        e = self._env.e
        self.action_space_shape = e.action_shape
        self.num_actions = e.A
        self.game_state_shape = e.state_shape
        self.child_policy, self.state_history, self.player_history, self.action_history = [], [], [], []
        self._state = self._env.reset(1, self._maps())
        self._refresh()

    def _maps(self):
        return None if self._map is None else [self._map]

    def _refresh(self):
        t, tv, pl, ln = self._env.status(self._state, self._maps())[0].tolist()
        self.terminal, self.terminal_value, self.agent_selection, self.length = bool(t), tv, pl, ln

    # -- getters (Games/Game.py) -----------------------------------------------------------------
    def get_state_shape(self):
        return self.game_state_shape

    def state_shape(self):
        return self.game_state_shape

    def get_action_space_shape(self):
        return self.action_space_shape

    def get_num_actions(self):
        return self.num_actions

    def get_current_player(self):
        return self.agent_selection

    def is_terminal(self):
        return self.terminal

    def get_terminal_value(self):
        return self.terminal_value

    def get_length(self):
        return self.length

    def get_winner(self):
        return 2 if self.terminal_value < 0 else (1 if self.terminal_value > 0 else 0)

    def get_state_from_history(self, i):
        return self.state_history[i]

    def store_state(self, state):
        self.state_history.append(state)

    def store_player(self, player):
        self.player_history.append(player)

    def store_action(self, action_coords):
        self.action_history.append(action_coords)

    def get_action_coords(self, action_i):  # Games/Game.py:96-98
        return np.unravel_index(action_i, self.get_action_space_shape())

    def get_action_index(self, action_coords):  # Games/Game.py:100-102
        return np.ravel_multi_index(action_coords, self.get_action_space_shape())

    # -- rules: all on the device -------------------------------------------------------------------
    def _mask_flat(self):
        return self._env.mask(self._state, self._maps())[0].cpu().numpy()

    def step(self, action_coords):
        if action_coords is None:
            return
        idx = int(self.get_action_index(tuple(int(x) for x in action_coords)))
        self.store_action(action_coords)
        try:
            self._env.step(self._state, [idx], self._maps())
        except _ffi.NzError:
            self.action_history.pop()
            raise Exception("Tried to play an illegal action!")  # Games/SCS/SCS_Game.py:382
        self._refresh()
        return self.terminal

    def generate_network_input(self):
        return self._env.encode(self._state, self._maps())[0:1].cpu()

    def generate_state_image(self):
        return self.generate_network_input()

    def store_search_statistics(self, node):  # tic_tac_toe.py:177-182 / SCS_Game.py:1517-1521
        kids = node.children
        total = sum(c.visit_count for c in kids.values())
        self.child_policy.append([kids[a].visit_count / total if a in kids else 0 for a in range(self.num_actions)])

    def make_target(self, i):
        return (self.terminal_value, self.child_policy[i])

    def shallow_clone(self):
        c = self.__class__.__new__(self.__class__)
        c.__dict__.update({k: v for k, v in self.__dict__.items()
                           if k not in ("child_policy", "state_history", "player_history", "action_history", "_state")})
        c.child_policy, c.state_history, c.player_history, c.action_history = [], [], [], []
        c._state = self._state.clone()
        return c

    def clone(self):
        c = self.shallow_clone()
        c.child_policy, c.state_history = list(self.child_policy), list(self.state_history)
        c.player_history, c.action_history = list(self.player_history), list(self.action_history)
        return c

    def reset(self, seed=None, options=None):
        self._init_common()

    def compact_state(self):
        """The 32-bit words the search engine takes as a root state."""
        return self._state[0]


class tic_tac_toe(DeviceGame):
    """Games/Tic_Tac_Toe/tic_tac_toe.py: 3x3, players 1 and 2, state planes [P1 stones, P2 stones]."""
    WIDTH, HEIGHT, TURNS, N_PLAYERS = 3, 3, 9, 2

    def __init__(self, device="cuda:0"):
        self._env = env_for("ttt", tic_tac_toe_spec, device)
        self._map = None
        self.total_action_planes = 1
        self._init_common()

    def spec(self):
        return tic_tac_toe_spec()

    def get_name(self):
        return "Tic_Tac_Toe"

    def get_dirname(self):
        return "Tic_Tac_Toe"

    def getBoardWidth(self):
        return self.WIDTH

    def getBoardHeight(self):
        return self.HEIGHT

    def possible_actions(self):
        return self._mask_flat().astype(np.float64).reshape(self.HEIGHT, self.WIDTH)  # tic_tac_toe.py:121-129

    @property
    def board(self):
        w = int(self._state[0, 0].item()) & 0xFFFFFFFF
        return [[1 if (w >> (r * 3 + c)) & 1 else (2 if (w >> (9 + r * 3 + c)) & 1 else 0) for c in range(3)] for r in range(3)]


class SCS_Game(DeviceGame):
    """Games/SCS/SCS_Game.py: SCS_Game(game_config_path, seed) — players 0 and 1."""
    N_PLAYERS = 2

    def __init__(self, game_config_path="", seed=None, device="cuda:0"):
        from .scs_config import ScsScenario

        if game_config_path == "":
            raise Exception("SCS_Game needs a game config path")
        self.scenario = ScsScenario(game_config_path, [seed])
        self._env = env_for(("scs", game_config_path, seed), self.scenario.spec, device)
        self._map = 0
        self.rows, self.columns = self.scenario.rows, self.scenario.cols
        self.turns, self.stacking_limit = self.scenario.turns, self.scenario.S
        self.total_action_planes = self.scenario.planes
        self.title = self.scenario.raw.get("Name", "Default_Game")
        self._init_common()

    def spec(self):
        return self.scenario.spec()

    def get_dirname(self):
        return "SCS"

    def get_name(self):
        return "SCS"

    def get_title(self):
        return self.title

    def getBoardColumns(self):
        return self.columns

    def getBoardRows(self):
        return self.rows

    def possible_actions(self):
        return self._mask_flat().astype(np.int8).reshape(self.action_space_shape)  # SCS_Game.py:395-484

    def generate_state(self):
        return self.generate_network_input()[0]
