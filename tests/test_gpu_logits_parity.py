"""GPU: the branch of Explorer.evaluate every REAL network takes (Search/Explorer.py:152-179: scipy.special.softmax of the
network's logits over ALL actions in f32, mask, np.sum, divide) against the oracle running the real scipy softmax.

The network is the deterministic stub turned into a logit network: logit = (h - 128) / 32 with the stub's integer hash h
(exact in f32 and in bf16), so both sides see bit-identical logits and the only difference is the softmax arithmetic
(device: expf / tile sum, scipy: SIMD exp / pairwise sum).  north_star tolerance: root value estimates and policy targets
within 1e-5 relative; visit counts / trajectories equal wherever no selection had a score gap below 1e-6."""
import os

import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu


class LogitStub:
    """DyadicStubNet + (p - 0.5) * 8 on the policy rows: exact, elementwise, on the current stream."""

    def __init__(self, engine, salt):
        from nuzero_b200.stubnet import DyadicStubNet

        self.e, self.stub = engine, DyadicStubNet(engine, salt=salt)

    def __call__(self):
        self.stub()
        self.e.policy.sub_(0.5).mul_(8.0)


def _oracle_net(A, salt):
    from oracle.stubnet_np import stub_forward

    def net(state):
        p, v = stub_forward(state, A, salt)
        return ((p - np.float32(0.5)) * np.float32(8.0)).astype(np.float32), v

    return net


def _compare(got, ref, gaps_per_move, report):
    """Moves are compared in order while the two games coincide; a difference is only accepted after a near tie."""
    for m in range(ref["length"]):
        same_kids = m < len(got["child_actions"]) and np.array_equal(got["child_actions"][m], ref["child_actions"][m])
        if same_kids:
            # priors at the root's expansion and after noise: the softmax difference, within 1e-5 relative (measured: ~1e-7)
            np.testing.assert_allclose(got["child_prior"][m], ref["child_prior"][m], rtol=1e-5, atol=0)
            report["max_prior_rel"] = max(report["max_prior_rel"], float(np.max(np.abs(got["child_prior"][m] - ref["child_prior"][m]) /
                                                                             np.abs(ref["child_prior"][m]))))
        if same_kids and np.array_equal(got["child_N"][m], ref["child_N"][m]) and got["root_N"][m] == ref["root_N"][m]:
            # identical visit counts: value sums can differ only through the priors' effect on ... nothing: W sums are sums of
            # the same network values along the same paths -> equal; root value estimate and policy target follow
            np.testing.assert_allclose(got["root_W"][m] / got["root_N"][m], ref["root_W"][m] / ref["root_N"][m], rtol=1e-5, atol=1e-12)
            pol_g = got["child_N"][m] / got["child_N"][m].sum()
            pol_r = ref["child_N"][m] / ref["child_N"][m].sum()
            np.testing.assert_allclose(pol_g, pol_r, rtol=1e-5, atol=0)
            report["moves_equal"] += 1
            if got["actions"][m] != ref["actions"][m]:
                raise AssertionError("same statistics, different action at move %d" % m)
            continue
        # the searches parted ways: legitimate only if some selection of this move was a near tie
        gap = min(gaps_per_move[m]) if gaps_per_move[m] else 0.0
        assert gap < 1e-6, "move %d differs although the smallest score gap of its search was %.3g" % (m, gap)
        report["diverged_after_near_tie"] += 1
        return
    assert got["length"] == ref["length"] and got["terminal_value"] == ref["terminal_value"]
    report["games_equal"] += 1


@pytest.mark.parametrize("policy_dtype", ["f32", "bf16"])
@pytest.mark.parametrize("training", [False, True])
def test_ttt_logit_network_matches_scipy_softmax_oracle(policy_dtype, training):
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.selfplay import game_record, group_games, run_until_idle
    from oracle import mcts, selfplay
    from oracle.ttt import TicTacToe

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 100
    G = 32
    rng = np.random.Generator(np.random.Philox(11))
    gm, un = rng.gamma(0.15, 1.0, size=(G, 10, 9)), rng.random(size=(G, 10, 3))
    e = SearchEngine(tic_tac_toe_spec(), cfg, G, training, policy_is_prob=False, leaf_dtype=_ffi.F32,
                     policy_dtype=_ffi.BF16 if policy_dtype == "bf16" else _ffi.F32, auto_advance=True, games_per_slot=1,
                     record_detail=True, pool_nodes=20000, tape_moves=10 if training else 0, tape_width=9 if training else 0)
    if training:
        e.set_tapes(gm, un)
    salts = list(range(300, 300 + G))
    run_until_idle(e, LogitStub(e, salts))
    games = group_games(e.drain_records()[0])
    assert len(games) == G
    report = dict(moves_equal=0, games_equal=0, diverged_after_near_tie=0, max_prior_rel=0.0)
    for gi in range(G):
        mcts.GAP_LOG = []
        marks = []

        class Net:
            def __init__(self, f):
                self.f = f

            def __call__(self, s):
                return self.f(s)

        tape = mcts.ReplayTape(gm[gi], un[gi]) if training else None
        # gaps per move: run_mcts is called once per move, note where each call starts in the log
        orig = mcts.run_mcts

        def traced(*a, **k):
            marks.append(len(mcts.GAP_LOG))
            return orig(*a, **k)

        mcts.run_mcts = traced
        try:
            ref = selfplay.play_game(TicTacToe(), _oracle_net(9, salts[gi]), cfg, training, False, tape)
        finally:
            mcts.run_mcts = orig
        log, mcts.GAP_LOG = mcts.GAP_LOG, None
        marks.append(len(log))
        gaps = [log[marks[i]:marks[i + 1]] for i in range(len(marks) - 1)]
        _compare(game_record(games[gi]), ref, gaps, report)
    print("logit parity (TTT, %s, training=%s): %s" % (policy_dtype, training, report))
    assert report["games_equal"] + report["diverged_after_near_tie"] == G
    assert report["games_equal"] >= G - 2 and report["max_prior_rel"] < 1e-5


def test_scs_logit_network_matches_scipy_softmax_oracle():
    """The f32 chain (int8 mask -> float32 priors, SCS_Game.py:399-408) with 525 logits per leaf."""
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.selfplay import game_record, group_games, run_until_idle
    from oracle import mcts, scs as oscs, selfplay

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 30
    path = os.path.join(golden_io.SCS_CONFIGS, "mirrored_config_5.yml")
    scn = ScsScenario(path, [None])
    G = 6
    e = SearchEngine(scn.spec(), cfg, G, False, policy_is_prob=False, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                     auto_advance=True, games_per_slot=1, record_detail=True, pool_nodes=60000, max_depth=256)
    e.set_maps([0] * G)
    e.reset()
    salts = list(range(40, 40 + G))
    run_until_idle(e, LogitStub(e, salts), max_launches=400000)
    games = group_games(e.drain_records()[0])
    assert len(games) == G
    sc = oscs.load_scenario(path, None)
    report = dict(moves_equal=0, games_equal=0, diverged_after_near_tie=0, max_prior_rel=0.0)
    for gi in range(G):
        mcts.GAP_LOG, marks = [], []
        orig = mcts.run_mcts

        def traced(*a, **k):
            marks.append(len(mcts.GAP_LOG))
            return orig(*a, **k)

        mcts.run_mcts = traced
        try:
            ref = selfplay.play_game(oscs.SCS(sc), _oracle_net(sc.A, salts[gi]), cfg, False, False, None)
        finally:
            mcts.run_mcts = orig
        log, mcts.GAP_LOG = mcts.GAP_LOG, None
        marks.append(len(log))
        gaps = [log[marks[i]:marks[i + 1]] for i in range(len(marks) - 1)]
        _compare(game_record(games[gi]), ref, gaps, report)
    print("logit parity (SCS mirrored_5): %s" % report)
    assert report["games_equal"] + report["diverged_after_near_tie"] == G
    assert report["moves_equal"] > 100 and report["max_prior_rel"] < 1e-5
