"""TEST INFRASTRUCTURE — numpy statement of the deterministic "dyadic" stub network used by the
parity protocol: probabilities are h/256 with integer h in [1,255] and the value is
k/128 with integer k in [-127,127], both pure integer functions of the encoded state and of a
per-game integer `salt` (so that deterministic games differ from one another).  Masked sums
of such numbers are exact in f32 in any order, so `prior = p/total` is a single IEEE division on
both sides.  The product has its own CUDA / torch statements of the same hash
(`nuzero_b200/csrc/stubnet.cuh`, `nuzero_b200/stubnet.py`); this file is the independent checker.
"""
import numpy as np

P = 65521


def feature_weights(n_features):
    i = np.arange(n_features, dtype=np.int64)
    return (i * 37 + 11) % 251 + 1, (i * 101 + 7) % 241 + 1


def action_multipliers(n_actions):
    a = np.arange(n_actions, dtype=np.int64)
    return (a * 40503 + 12345) % P, (a * 30011 + 54321) % P, (a * 977 + 101) % P


def stub_forward(state, n_actions, salt=0):
    """state: float array of any shape holding ONE position -> (probs f32 [A], value f32)."""
    s = np.asarray(state, dtype=np.float32).reshape(-1)
    q = np.rint(s * np.float32(64.0)).astype(np.int64)  # rint = round-half-even, like torch.round
    w1, w2 = feature_weights(s.size)
    s1 = int(((q * w1).sum() + salt) % P)
    s2 = int(((q * w2).sum() + 3 * salt) % P)
    m1, m2, m3 = action_multipliers(n_actions)
    h = ((s1 * m1 + s2 * m2 + m3) % P) % 255 + 1
    k = ((s1 * 7 + s2 * 13 + 5) % P) % 255 - 127
    return (h.astype(np.float32) / np.float32(256.0)), np.float32(k) / np.float32(128.0)


class StubNetwork:
    """Duck-types the reference `Network_Manager` (Neural_Networks/Network_Manager.py:46-64):
    `.inference(state, training, iters) -> (p, v)` with torch tensors, `.check_devices()`."""

    def __init__(self, action_shape, salt=0):
        self.salt = int(salt)
        self.action_shape = tuple(action_shape)
        self.n_actions = int(np.prod(action_shape))
        self.calls = 0

    def check_devices(self):
        return None

    def is_recurrent(self):
        return True

    def inference(self, state, training=False, iters_to_do=2, interim_thought=None):
        import torch

        self.calls += 1
        p, v = stub_forward(state.detach().cpu().numpy(), self.n_actions, self.salt)
        return (torch.from_numpy(p.reshape((1,) + self.action_shape)),
                torch.tensor([[float(v)]], dtype=torch.float32))
