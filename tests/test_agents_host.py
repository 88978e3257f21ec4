"""Host-side agent logic (no GPU): the random agent's inverse-CDF rule against numpy's own np.random.choice, and the
single-game PolicyAgent of the drop-in API against the oracle's restatement (which the reference fixtures pin)."""
import numpy as np
import torch


def test_choice_from_uniform_is_numpy_random_choice():
    from nuzero_b200.tester import _choice_from_uniform
    from oracle.match import random_choice

    rng = np.random.default_rng(0)
    for trial in range(300):
        n = int(rng.integers(1, 40))
        mask = (rng.random(n) < 0.4).astype(np.int8)
        if mask.sum() == 0:
            mask[int(rng.integers(0, n))] = 1
        seed = int(rng.integers(0, 2 ** 31))
        np.random.seed(seed)
        u = np.random.random()
        np.random.seed(seed)
        want = int(np.random.choice(n, p=mask / mask.sum()))   # RandomAgent.py:10-15
        assert _choice_from_uniform(mask, u) == want
        assert random_choice(mask, u) == want


class _FakeGame:
    def __init__(self, mask):
        self.mask = np.asarray(mask, dtype=np.int8)
        self.num_actions = len(mask)

    def possible_actions(self):
        return self.mask.reshape(1, 1, -1)

    def generate_network_input(self):
        return torch.zeros(1, 1, 1, self.num_actions)

    def get_action_coords(self, action_i):
        return (0, 0, int(action_i))


class _FakeNet:
    def __init__(self, logits):
        self.logits = torch.as_tensor(logits, dtype=torch.float32)

    def inference(self, state, training, iters):
        return self.logits.reshape(1, -1), torch.zeros(1, 1)


def test_policy_agent_follows_the_reference_rule_and_its_random_draws():
    """PolicyAgent.py:21-68: raw arg-max when legal; otherwise the masked arg-max AFTER one discarded np.random.choice."""
    from nuzero_b200.agents import PolicyAgent
    from oracle.match import policy_choice

    rng = np.random.default_rng(1)
    for trial in range(200):
        n = int(rng.integers(2, 30))
        mask = (rng.random(n) < 0.5).astype(np.int8)
        if mask.sum() == 0:
            mask[0] = 1
        logits = rng.normal(size=n).astype(np.float32) * 3
        probs = torch.softmax(torch.from_numpy(logits), 0).numpy()
        seed = int(rng.integers(0, 2 ** 31))
        np.random.seed(seed)
        u = np.random.random()
        want, used = policy_choice(probs, mask, u)
        np.random.seed(seed)
        got = PolicyAgent(_FakeNet(logits), 2).choose_action(_FakeGame(mask))[2]
        after = np.random.random()
        assert got == want
        np.random.seed(seed)
        seq = [np.random.random() for _ in range(2)]
        assert after == seq[used]      # exactly `used` uniforms were consumed


def test_policy_agent_with_no_mass_on_legal_actions_plays_a_random_legal_one():
    from nuzero_b200.agents import PolicyAgent

    mask = np.array([0, 1, 0, 1, 1], dtype=np.int8)
    logits = np.array([50.0, -200.0, 60.0, -200.0, -200.0], dtype=np.float32)   # softmax underflows on the legal actions
    for seed in range(20):
        np.random.seed(seed)
        want = int(np.random.choice(5, p=mask / mask.sum()))
        np.random.seed(seed)
        assert PolicyAgent(_FakeNet(logits), 2).choose_action(_FakeGame(mask))[2] == want
