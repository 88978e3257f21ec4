"""TEST INFRASTRUCTURE — Tic-Tac-Toe restated from Games/Tic_Tac_Toe/tic_tac_toe.py on two
9-bit boards (bit r*3+c).  Players are 1 and 2 (ref :27); player 1 moves on even lengths."""
import numpy as np

_LINES = (0o007, 0o070, 0o700, 0o111, 0o222, 0o444, 0o421, 0o124)


class TicTacToe:
    action_space_shape = (1, 3, 3)  # ref :30
    num_actions = 9  # ref :31
    state_shape = (2, 3, 3)  # ref :33
    prior_is_f64 = True  # mask is np.ones float64 (ref :122) -> f64 priors (Explorer.py:168-179)

    def __init__(self):
        self.stones = [0, 0]
        self.length = 0
        self.terminal = False
        self.terminal_value = 0

    # -- interface used by the search (ref :58-92) -------------------------------------------
    def get_current_player(self):
        return self.length % 2 + 1  # ref :165

    def is_terminal(self):
        return self.terminal

    def get_terminal_value(self):
        return self.terminal_value

    def get_length(self):
        return self.length

    def get_num_actions(self):
        return 9

    def clone(self):  # ref :267-273 (board, player, length only: terminal flags restart False)
        c = TicTacToe()
        c.stones = list(self.stones)
        c.length = self.length
        return c

    def legal_mask(self):
        """ref :121-129 — float64 ones with occupied cells zeroed, flattened."""
        occ = self.stones[0] | self.stones[1]
        return np.array([0.0 if (occ >> i) & 1 else 1.0 for i in range(9)], dtype=np.float64)

    def step(self, action):
        """ref :161-167 + check_terminal :198-262.  `action` is the flat index (plane 0)."""
        me = self.length % 2
        self.stones[me] |= 1 << action
        self.length += 1
        value, done = 0, False
        if any((self.stones[0] & l) == l for l in _LINES):  # P1 line tested first (ref :243)
            value, done = 1, True
        elif any((self.stones[1] & l) == l for l in _LINES):
            value, done = -1, True
        if self.length == 9:  # ref :252
            done = True
        if done:
            self.terminal, self.terminal_value = True, value
        return done

    def encode(self):
        """ref :135-159 — planes [P1 stones, P2 stones]; the player plane is built but dropped."""
        out = np.zeros((1, 2, 3, 3), dtype=np.float32)
        for p in range(2):
            for i in range(9):
                if (self.stones[p] >> i) & 1:
                    out[0, p, i // 3, i % 3] = 1.0
        return out

    def get_winner(self):  # ref :169-171
        return {1: 1, -1: 2, 0: 0}[self.terminal_value]
