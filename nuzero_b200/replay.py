"""In-process sink with the interface of Training/ReplayBuffer.py:10-62 (the reference's Ray actor):
a window of games, entries `(state, (value_target, policy_target), game_index)`."""
import random

import numpy as np


class ReplayBuffer:
    def __init__(self, window_size, batch_size):
        self.window_size, self.batch_size = window_size, batch_size
        self.buffer, self.n_games, self.full = [], 0, False

    def save_game(self, game, game_index):  # ReplayBuffer.py:24-36
        if self.n_games >= self.window_size:
            self.full = True
        else:
            self.full = False
            self.n_games += 1
        for i in range(len(game.state_history)):
            entry = (game.get_state_from_history(i), game.make_target(i), game_index)
            if self.full:
                self.buffer.pop(0)
            self.buffer.append(entry)

    def shuffle(self):
        random.shuffle(self.buffer)

    def get_slice(self, start_index, last_index):
        return self.buffer[start_index:last_index]

    def get_sample(self, batch_size, replace, probs):
        args = [len(self.buffer), batch_size, replace] + ([probs] if probs != [] else [])
        return [self.buffer[i] for i in np.random.choice(*args)]

    def get_buffer(self):
        return self.buffer

    def len(self):
        return len(self.buffer)

    def played_games(self):
        return self.n_games
