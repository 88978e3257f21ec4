"""GPU: the throughput mode with several leaves of one game waiting at the network (nz_config.virtual_loss_width > 1).
It is NOT the reference's sequencing, so there is no bit-exact oracle for it; what is checked are the invariants every
PUCT tree satisfies once nothing is pending, that width 1 is untouched (the parity suites), and that the search still
plays well."""
import os

import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu


def _cfg(sims):
    cfg = {k: dict(v) if isinstance(v, dict) else v for k, v in golden_io.load("ttt_p0_s25_salt0")["cfg"].items()}
    cfg["Simulation"]["mcts_simulations"] = sims
    return cfg


def _check_tree(e, g):
    """N(node) = 1 + sum N(children) for every expanded node reachable from the root; |W| <= N; nothing pending."""
    from nuzero_b200 import _ffi

    N = e.node_N[g].cpu().numpy().astype(np.int64)
    W = e.node_W[g].cpu().numpy()
    link = np.stack([e.node_base[g].cpu().numpy().astype(np.int64) & 0xFFFFFFFF, e.node_K[g].cpu().numpy().astype(np.int64)], 1)
    root = int(e.ctl[g, _ffi.CTL_ROOT])
    assert root == 0 and (link[0, 1] == 0 or link[0, 0] == 1)  # the slot layout: root at node 0, its children from node 1
    assert int(e.ctl[g, _ffi.CTL_N_PENDING]) == 0
    stack, seen = [root], 0
    while stack:
        n = stack.pop()
        seen += 1
        base, k = int(link[n, 0]), int(link[n, 1] & 0xFFFF)
        assert abs(W[n]) <= N[n] + 1e-9
        if k:
            kids = list(range(base, base + k))
            assert N[n] == 1 + int(N[kids].sum()), (g, n, N[n], N[kids])
            stack += [c for c in kids if N[c] > 0]
    return int(N[root]), seen


@pytest.mark.parametrize("V", [2, 4, 8])
def test_ttt_tree_invariants_hold_with_parked_leaves(V):
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.stubnet import DyadicStubNet

    G, sims = 16, 120
    e = SearchEngine(tic_tac_toe_spec(), _cfg(sims), G, False, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                     auto_advance=False, pool_nodes=8000, max_sims_per_launch=16, virtual_loss=V)
    assert e.leaf.shape[0] == G * V and e.policy.shape == (G * V, 9)
    net = DyadicStubNet(e, salt=[g for g in range(G) for _ in range(V)])
    launches = 0
    carried = np.zeros(G, dtype=np.int64)
    for move in range(3):
        while not bool((e.phases() == _ffi.PHASE_MOVE_READY).all()):
            e.advance()
            net()
            launches += 1
            assert launches < 5000
        e.raise_on_error()
        for g in range(G):
            root_n, _ = _check_tree(e, g)
            assert root_n == carried[g] + sims
            chosen = int(e.ctl[g, _ffi.CTL_CHOSEN])
            carried[g] = int(e.node_N[g, 1 + chosen])
        e.commit_moves()
        e.raise_on_error()
    # several simulations per network round trip: far fewer launches than one leaf per game per launch would need
    assert launches < 3 * sims * 0.8


def test_ttt_selfplay_with_parked_leaves_finishes_and_feeds_the_replay_buffer():
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.replay import DeviceReplayBuffer
    from nuzero_b200.selfplay import run_until_idle
    from nuzero_b200.stubnet import DyadicStubNet

    G, V, sims = 64, 4, 100
    e = SearchEngine(tic_tac_toe_spec(), _cfg(sims), G, True, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                     auto_advance=True, games_per_slot=2, pool_nodes=8000, max_sims_per_launch=8, virtual_loss=V, seed=3)
    run_until_idle(e, DyadicStubNet(e, salt=[g for g in range(G) for _ in range(V)]))
    drb = DeviceReplayBuffer(e, 1000, 8, capacity=G * 2 * 9)
    n = drb.ingest()
    assert drb.played_games() == 2 * G and 5 * 2 * G <= n <= 9 * 2 * G
    rows = drb._rows()
    assert torch.allclose(drb.policy[rows].sum(1), torch.ones(n, device=rows.device), atol=1e-6)
    assert set(drb.value[rows].cpu().tolist()) <= {-1.0, 0.0, 1.0}
    c = e.counters()
    assert c["sims"] >= n * sims and c["games"] == 2 * G


def test_parked_leaves_do_not_cost_playing_strength():
    """MctsAgent (keep_subtree) against RandomAgent on Tic-Tac-Toe: with terminal values doing the work (the stub network is
    uninformative) the search must not lose as the first player, with one or with four leaves in flight."""
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import tic_tac_toe_spec
    from nuzero_b200.stubnet import DyadicStubNet
    from nuzero_b200.tester import BatchedTester

    G = 128
    tape = np.random.default_rng(0).random((G, 12))
    res = {}
    for V in (1, 4):
        t = BatchedTester(tic_tac_toe_spec(), _cfg(200), G, lambda e: DyadicStubNet(e, salt=[g for g in range(G) for _ in range(e.V)]),
                          policy_is_prob=True, leaf_dtype=_ffi.F32, pool_nodes=8000, virtual_loss=V, max_sims_per_launch=8)
        res[V] = t.win_rates(t.play(1, unif_tape=tape), mcts_is_first=True)
    for V in (1, 4):
        mcts, rnd, draws = res[V]
        assert rnd <= 0.02 and mcts >= 0.85, (V, res[V])
    assert abs(res[1][0] - res[4][0]) < 0.08


def test_scs_tree_invariants_hold_with_parked_leaves():
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.stubnet import DyadicStubNet

    scn = ScsScenario(os.path.join(golden_io.SCS_CONFIGS, "randomized_config_5.yml"), [1, 2])
    G, V, sims = 8, 3, 60
    e = SearchEngine(scn.spec(), _cfg(sims), G, False, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                     auto_advance=False, pool_nodes=30000, max_depth=128, max_sims_per_launch=8, virtual_loss=V)
    e.set_maps([g % 2 for g in range(G)])
    e.reset()
    net = DyadicStubNet(e, salt=[g for g in range(G) for _ in range(V)])
    for move in range(4):
        for _ in range(5000):
            if bool((e.phases() == _ffi.PHASE_MOVE_READY).all()):
                break
            e.advance()
            net()
        e.raise_on_error()
        assert bool((e.phases() == _ffi.PHASE_MOVE_READY).all())
        for g in range(G):
            _check_tree(e, g)
        e.commit_moves()
        e.raise_on_error()
