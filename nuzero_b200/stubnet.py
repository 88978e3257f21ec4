"""Deterministic "dyadic" stub network of the parity protocol.

`DyadicStubNet` runs the one-launch CUDA kernel `nz_stubnet_forward`; `torch_reference` is the same
integer hash written with torch ops (used by tests to cross-check the kernel on the device).  The
independent CPU checker is oracle/stubnet_np.py.
"""
import ctypes as C

import torch

from . import _ffi
from ._ffi import check, lib

P = 65521


class DyadicStubNet:
    """Network stand-in for the engine: fills engine.policy / engine.value from engine.leaf."""

    def __init__(self, engine, salt=None, uid_mul=0):
        self.e = engine
        self.salt = None
        if salt is not None:
            self.salt = torch.as_tensor(salt, dtype=torch.int32, device=engine.device).contiguous()
        self.uid_mul = int(uid_mul)
        self.F = int(engine.leaf[0].numel())

    def __call__(self):
        e = self.e
        uid = None
        if self.uid_mul:
            if e.V != 1:
                raise _ffi.NzError("DyadicStubNet: per-game salts from the uid need one leaf row per game (virtual_loss = 1)")
            uid = C.c_void_p(e.ctl.data_ptr() + 4 * _ffi.CTL_UID)
        check(lib().nz_stubnet_forward(
            C.c_void_p(e.leaf.data_ptr()), e.c.leaf_dtype,
            None if self.salt is None else C.c_void_p(self.salt.data_ptr()), uid, _ffi.CTL_WORDS, self.uid_mul,
            e.rows, self.F, e.A, C.c_void_p(e.policy.data_ptr()), e.c.policy_dtype,
            C.c_void_p(e.value.data_ptr()), e._stream()))


def torch_reference(leaf, n_actions, salt=None):
    """Same hash in torch ops: leaf [n, ...] -> (probs f32 [n, A], value f32 [n])."""
    n = leaf.shape[0]
    x = leaf.reshape(n, -1).to(torch.float32)
    F = x.shape[1]
    q = torch.round(x * 64.0).to(torch.int64)
    i = torch.arange(F, device=x.device, dtype=torch.int64)
    w1, w2 = (i * 37 + 11) % 251 + 1, (i * 101 + 7) % 241 + 1
    sl = torch.zeros(n, dtype=torch.int64, device=x.device) if salt is None else salt.to(torch.int64)
    s1 = torch.remainder((q * w1).sum(1) + sl, P)
    s2 = torch.remainder((q * w2).sum(1) + 3 * sl, P)
    a = torch.arange(n_actions, device=x.device, dtype=torch.int64)
    m1, m2, m3 = (a * 40503 + 12345) % P, (a * 30011 + 54321) % P, (a * 977 + 101) % P
    h = ((s1[:, None] * m1 + s2[:, None] * m2 + m3) % P) % 255 + 1
    k = ((s1 * 7 + s2 * 13 + 5) % P) % 255 - 127
    return h.to(torch.float32) / 256.0, k.to(torch.float32) / 128.0


class StubNetworkManager:
    """Network_Manager-shaped wrapper of the stub for the drop-in API (Explorer.run_mcts / Gamer):
    `.inference(state, training, iters)` evaluates one position, `.bind_engine(engine)` gives the
    batched one-launch form.  `outputs_probabilities` tells the engine to skip the softmax."""
    outputs_probabilities = True

    def __init__(self, action_shape, salt=0, uid_mul=0):
        self.action_shape = tuple(action_shape)
        self.salt, self.uid_mul = int(salt), int(uid_mul)
        self.calls = 0

    def check_devices(self):
        return None

    def is_recurrent(self):
        return True

    def inference(self, state, training=False, iters_to_do=2, interim_thought=None):
        self.calls += 1
        n = state.shape[0]
        A = 1
        for d in self.action_shape:
            A *= d
        salt = torch.full((n,), self.salt, dtype=torch.int64, device=state.device)
        p, v = torch_reference(state, A, salt)
        return p.reshape((n,) + self.action_shape), v.reshape(n, 1)

    def bind_engine(self, engine):
        salts = torch.full((engine.G,), self.salt, dtype=torch.int32)
        return DyadicStubNet(engine, salt=salts, uid_mul=self.uid_mul)
