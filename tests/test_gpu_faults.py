"""GPU: fault paths and edge cases of the C ABI — the places where the reference raises (illegal action,
Games/SCS/SCS_Game.py:382; max() of an empty child list) and the engine's own resource limits (node pool, path depth,
record arena).  Device-side faults never trap: they set per-slot error bits that the host turns into exceptions."""
import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu


def _cfg(sims=50):
    cfg = {k: dict(v) if isinstance(v, dict) else v for k, v in golden_io.load("ttt_p0_s25_salt0")["cfg"].items()}
    cfg["Simulation"]["mcts_simulations"] = sims
    return cfg


def _ttt_engine(**kw):
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec

    args = dict(policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32, auto_advance=True, games_per_slot=1,
                pool_nodes=4000)
    args.update(kw)
    return SearchEngine(tic_tac_toe_spec(), _cfg(args.pop("sims", 50)), args.pop("G", 8), args.pop("training", False), **args)


def _run(e, launches=400):
    from nuzero_b200.stubnet import DyadicStubNet

    net = DyadicStubNet(e, salt=list(range(e.G)))
    for _ in range(launches):
        e.advance()
        net()
    torch.cuda.synchronize()


def test_node_pool_exhaustion_is_reported_not_trapped():
    from nuzero_b200 import _ffi

    e = _ttt_engine(pool_nodes=40, compact=False)  # 50 simulations need more than 40 nodes
    _run(e)
    assert bool((e.errors() & _ffi.ERR_POOL_FULL).any())
    assert bool((e.phases()[(e.errors() & _ffi.ERR_POOL_FULL) != 0] == _ffi.PHASE_ERROR).all())
    with pytest.raises(_ffi.NzError, match="node pool full"):
        e.raise_on_error()
    # the device is still usable: a fresh engine on the same context plays to the end
    e2 = _ttt_engine()
    _run(e2)
    e2.raise_on_error()
    assert bool((e2.phases() == _ffi.PHASE_IDLE).all())


def test_path_deeper_than_max_depth_is_reported():
    from nuzero_b200 import _ffi

    e = _ttt_engine(max_depth=3, sims=200)
    _run(e, 800)
    assert bool((e.errors() & _ffi.ERR_DEPTH).any())
    with pytest.raises(_ffi.NzError, match="max_depth"):
        e.raise_on_error()


def test_record_arena_overflow_counts_dropped_records():
    from nuzero_b200 import _ffi
    from nuzero_b200.replay import DeviceReplayBuffer

    e = _ttt_engine(arena_words=64, G=16)  # room for two or three records
    _run(e)
    recs, dropped = e.drain_records()
    assert dropped > 0 and len(recs) >= 1
    assert bool((e.errors() & _ffi.ERR_ARENA_FULL).any())
    e.raise_on_error()  # a dropped record is not fatal for the search itself
    e3 = _ttt_engine(arena_words=64, G=16)
    _run(e3)
    with pytest.raises(_ffi.NzError, match="dropped"):
        DeviceReplayBuffer(e3, 10, 4, capacity=200).ingest()


def test_env_step_raises_on_illegal_actions_like_the_reference():
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import EnvOps

    e = _ttt_engine(auto_advance=False)
    env = EnvOps(e)
    st = env.reset(3)
    env.step(st, [4, 4, 0])
    with pytest.raises(_ffi.NzError, match="illegal"):
        env.step(st.clone(), [4, 1, 1])  # occupied cell in game 0
    with pytest.raises(_ffi.NzError, match="illegal"):
        env.step(st.clone(), [0, 9, 1])  # action outside the action space
    with pytest.raises(_ffi.NzError, match="illegal"):
        env.step(st.clone(), [0, -1, 1])
    # a finished game accepts no further action
    s = env.reset(1)
    for a in [0, 3, 1, 4, 2]:  # player 1 completes the top row
        env.step(s, [a])
    assert env.status(s).cpu().numpy()[0].tolist()[:2] == [1, 1]
    with pytest.raises(_ffi.NzError, match="illegal"):
        env.step(s, [8])
    # empty batches are a no-op
    assert env.mask(st[:0]).shape == (0, 9) and env.encode(st[:0]).shape[0] == 0 and env.status(st[:0]).shape == (0, 4)


def test_commit_of_an_action_that_is_not_a_root_child_is_an_error():
    from nuzero_b200 import _ffi
    from nuzero_b200.stubnet import DyadicStubNet

    e = _ttt_engine(auto_advance=False, G=2)
    net = DyadicStubNet(e, salt=[0, 1])
    while not bool((e.phases() == _ffi.PHASE_MOVE_READY).all()):
        e.advance()
        net()
    e.commit_moves([4, 4])
    e.raise_on_error()
    while not bool((e.phases() == _ffi.PHASE_MOVE_READY).all()):
        e.advance()
        net()
    e.commit_moves([4, 0])  # cell 4 is taken in both games: slot 0 faults, slot 1 moves on
    assert int(e.errors()[0]) & _ffi.ERR_ILLEGAL and int(e.errors()[1]) == 0
    assert int(e.phases()[0]) == _ffi.PHASE_ERROR and int(e.phases()[1]) == _ffi.PHASE_READY
    with pytest.raises(_ffi.NzError, match="illegal action"):
        e.raise_on_error()


def test_unbound_or_bad_arguments_fail_with_a_message():
    import ctypes as C

    from nuzero_b200 import _ffi

    L = _ffi.lib()
    e = _ttt_engine()
    assert L.nz_replay_decode(e.h, None, None, None, None, None, 4, None) != 0 and b"null" in L.nz_last_error()
    assert L.nz_env_step(e.h, C.c_void_p(e.gstate.data_ptr()), None, None, None, 1, None) != 0
    x = torch.zeros(256, 64, device="cuda", dtype=torch.bfloat16)
    nbr = torch.zeros(1, 1, device="cuda", dtype=torch.int32)
    for bad in (dict(cin=60), dict(n_pad=24), dict(n_pad=512), dict(taps=10)):
        a = dict(cin=64, n_pad=64, taps=1)
        a.update(bad)
        rc = L.nz_hexconv_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(nbr.data_ptr()), C.c_void_p(x.data_ptr()), None,
                               C.c_void_p(x.data_ptr()), 256, 1, a["taps"], a["cin"], a["n_pad"], 64, 0, 0, None)
        assert rc != 0 and b"nz_hexconv_bf16" in L.nz_last_error()
