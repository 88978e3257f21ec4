"""Evaluation-side callers of the search (Testing/Agents/Generic/*.py in the reference): MctsAgent is
the second caller of `Explorer.run_mcts` (training=False — the noise-free mode the parity fixtures use),
RandomAgent / PolicyAgent are the opponents the reference's Tester pits it against."""
import numpy as np
import torch

from .search import Explorer, Node


class Agent:
    def choose_action(self, game):
        raise NotImplementedError

    def new_game(self, *args, **kwargs):
        return None

    def name(self):
        return "Agent"


class MctsAgent(Agent):
    """Testing/Agents/Generic/MctsAgent.py:13-78: most-visited action of a noise-free search."""

    def __init__(self, search_config, network, recurrent_iterations=2, cache=None, device="cuda:0", pool_nodes=None):
        self.explorer = Explorer(search_config, False, device=device, pool_nodes=pool_nodes)
        self.keep_subtree = search_config["Simulation"]["keep_subtree"]
        self.root_node = Node(0)
        self.network, self.recurrent_iterations, self.cache = network, recurrent_iterations, cache

    def choose_action(self, game):
        action_i, chosen_child, _ = self.explorer.run_mcts(game, self.network, self.root_node,
                                                           self.recurrent_iterations, self.cache)
        if self.keep_subtree:
            self.root_node = chosen_child
        return game.get_action_coords(action_i)

    def update_subtree(self, game, action_i):
        # MctsAgent.py:35-39: search once more on the opponent's turn, then follow the move actually played
        self.explorer.run_mcts(game, self.network, self.root_node, self.recurrent_iterations, self.cache)
        self.root_node = self.root_node.get_child(action_i)

    def new_game(self, *args, cache=None):
        self.root_node = Node(0)
        if cache is not None:
            self.cache = cache

    def set_search_config(self, search_config=None, **kwargs):
        if search_config is None:
            raise Exception("No search config provided.")
        self.explorer.set_search_config(search_config)
        self.keep_subtree = search_config["Simulation"]["keep_subtree"]

    def set_network(self, network):
        self.network = network

    def set_recurrent_iterations(self, recurrent_iterations):
        self.recurrent_iterations = recurrent_iterations

    def name(self):
        return "MCTS Agent"


class RandomAgent(Agent):
    """Testing/Agents/Generic/RandomAgent.py: uniform over the legal actions."""

    def choose_action(self, game):
        mask = game.possible_actions().flatten()
        return game.get_action_coords(np.random.choice(game.num_actions, p=mask / mask.sum()))

    def name(self):
        return "Random Agent"


class PolicyAgent(Agent):
    """Testing/Agents/Generic/PolicyAgent.py: arg-max of the network policy restricted to legal actions."""

    def __init__(self, network, recurrent_iterations=2, cache=None):
        self.network, self.recurrent_iterations, self.cache = network, recurrent_iterations, cache

    def choose_action(self, game):
        logits, _ = self.network.inference(game.generate_network_input(), False, self.recurrent_iterations)
        probs = torch.softmax(torch.as_tensor(logits).float().flatten(), 0).cpu().numpy()
        raw = int(np.argmax(probs))
        mask = game.possible_actions().flatten()
        if mask[raw]:
            return game.get_action_coords(raw)
        probs = probs * mask
        total = np.sum(probs)
        if total != 0:
            probs = probs / total
            np.random.choice(game.num_actions, p=probs)  # PolicyAgent.py:52 draws (and discards) a sample here
            return game.get_action_coords(int(np.argmax(probs)))
        return game.get_action_coords(np.random.choice(game.num_actions, p=mask / mask.sum()))

    def new_game(self, *args, cache=None, **kwargs):
        if cache is not None:
            self.cache = cache

    def set_network(self, network):
        self.network = network

    def set_recurrent_iterations(self, recurrent_iterations):
        self.recurrent_iterations = recurrent_iterations

    def name(self):
        return "Policy Agent"


def play_match(game, agents, max_moves=10_000):
    """One game between two agents, indexed by the order in which the players first move
    (cf. Testing/Tester.py:46-121).  Returns the winner (0 draw, 1, 2) and the number of moves."""
    for a in agents:
        a.new_game()
    first = game.get_current_player()
    moves = 0
    while not game.is_terminal() and moves < max_moves:
        mover = 0 if game.get_current_player() == first else 1
        agent, opponent = agents[mover], agents[1 - mover]
        action_coords = agent.choose_action(game)
        # Tester.py:92-94: an MctsAgent that keeps its sub-tree searches on the opponent's turn too and follows the move
        if isinstance(opponent, MctsAgent) and opponent.keep_subtree:
            opponent.update_subtree(game, int(game.get_action_index(action_coords)))
        game.step(action_coords)
        moves += 1
    return game.get_winner(), moves
