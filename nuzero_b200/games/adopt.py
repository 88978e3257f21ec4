"""Reference game objects on the drop-in surface: `Explorer.run_mcts(game, ...)` and the agents accept the reference's own
`Games/Tic_Tac_Toe/tic_tac_toe.py` and `Games/SCS/SCS_Game.py` instances (Games/Game.py:3-106 duck type) as well as this
package's device games.  A reference object is mirrored by a device game:
* Tic-Tac-Toe: the compact state word is built from `board`, `length` and the terminal fields (the reference's `step` does
  not record its actions, tic_tac_toe.py:161-167);
* SCS: the mirror replays `action_history` (SCS_Game.step stores every action, SCS_Game.py:384) through the environment
  kernels — incrementally, so a game that is searched move after move costs one kernel call per new move.
The mirror lives on the reference object (`_nz_mirror`).

The reference's SCS_Game does not remember its config path or seed (SCS_Game.py:78-140): pass them as
`Explorer(..., game_args=(path, seed))` / `to_device_game(game, (path, seed))`, or set `game._nz_args = (path, seed)`.
"""
import numpy as np


def is_device_game(game):
    return hasattr(game, "compact_state") and hasattr(game, "spec")


def action_index(game, coords):
    """Flat action index of an entry of `action_history` (coords over the game's action-space shape, Games/Game.py:100-102)."""
    if np.isscalar(coords):
        return int(coords)
    return int(np.ravel_multi_index(tuple(int(x) for x in coords), tuple(game.get_action_space_shape())))


def _family(game):
    name = type(game).__name__
    if name == "tic_tac_toe" or (hasattr(game, "board") and np.asarray(game.board).shape == (3, 3) and not hasattr(game, "stacking_limit")):
        return "ttt"
    if name == "SCS_Game" or hasattr(game, "stacking_limit"):
        return "scs"
    raise TypeError("cannot mirror a %s on the device: only Tic-Tac-Toe and SCS games have device kernels" % name)


def to_device_game(game, game_args=None, device="cuda:0"):
    """`game` itself when it already is a device game; otherwise the device mirror of a reference game object, brought up
    to date with the moves played on it since the last call."""
    if is_device_game(game):
        return game
    from .device_game import SCS_Game, tic_tac_toe

    if _family(game) == "ttt":
        return _mirror_ttt(game, device)
    history = [action_index(game, a) for a in game.action_history]
    mirror = getattr(game, "_nz_mirror", None)
    if mirror is not None and (len(history) < len(mirror._applied) or history[:len(mirror._applied)] != mirror._applied):
        mirror = None  # the object was reset or is another line of play: start again
    if mirror is None:
        args = getattr(game, "_nz_args", None) or game_args
        if not args:
            raise ValueError("a reference SCS_Game does not remember its scenario: pass game_args=(game_config_path, seed)")
        mirror = SCS_Game(args[0], args[1] if len(args) > 1 else None, device=device)
        mirror._applied = []
        try:
            game._nz_mirror = mirror
        except AttributeError:
            pass
    for a in history[len(mirror._applied):]:
        mirror.step(mirror.get_action_coords(a))
        mirror._applied.append(a)
    if int(mirror.get_length()) != int(game.get_length()) or bool(mirror.is_terminal()) != bool(game.is_terminal()):
        raise ValueError("the device mirror disagrees with the reference game object (length %d vs %d): its action_history does "
                         "not describe the position" % (mirror.get_length(), game.get_length()))
    return mirror


def ttt_state_word(board, length, terminal, terminal_value):
    """csrc/game_ttt.cuh: bits 0-8 player 1's stones (cell r * 3 + c), 9-17 player 2's, 18-21 length, 22 terminal,
    23-24 terminal value + 1."""
    w = 0
    for r in range(3):
        for c in range(3):
            if board[r][c] == 1:
                w |= 1 << (r * 3 + c)
            elif board[r][c] == 2:
                w |= 1 << (9 + r * 3 + c)
    return w | (int(length) << 18) | ((1 if terminal else 0) << 22) | ((int(terminal_value) + 1) << 23)


def _mirror_ttt(game, device):
    from .device_game import tic_tac_toe

    mirror = getattr(game, "_nz_mirror", None)
    if mirror is None:
        mirror = tic_tac_toe(device=device)
        try:
            game._nz_mirror = mirror
        except AttributeError:
            pass
    terminal = bool(game.is_terminal())
    word = ttt_state_word(game.board, game.get_length(), terminal, game.get_terminal_value() if terminal else 0)
    if (int(mirror._state[0, 0].item()) & 0xFFFFFFFF) != word:
        mirror._state[0, 0] = word - (1 << 32) if word >= (1 << 31) else word
        mirror._refresh()
    return mirror
