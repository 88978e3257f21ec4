"""GPU: the CUDA search path against (a) the golden fixtures generated from the real reference and
(b) the CPU oracle on fresh seeded games.  Bit-exact on tree statistics, priors, trajectories."""
import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu


def _engine_for(cfg, training, n_games, tapes=None, **kw):
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec

    tm, tw = (0, 0) if tapes is None else tapes[0].shape[1:3]
    e = SearchEngine(tic_tac_toe_spec(), cfg, n_games, training, policy_is_prob=True, leaf_dtype=_ffi.F32,
                     policy_dtype=_ffi.F32, auto_advance=True, games_per_slot=1, record_detail=True,
                     tape_moves=tm, tape_width=tw, pool_nodes=40000, **kw)
    if tapes is not None:
        e.set_tapes(*tapes)
    return e


def _play(e, salts):
    from nuzero_b200.engine import EnvOps
    from nuzero_b200.selfplay import game_record, group_games, run_until_idle
    from nuzero_b200.stubnet import DyadicStubNet

    run_until_idle(e, DyadicStubNet(e, salt=salts))
    recs, dropped = e.drain_records()
    assert dropped == 0
    games = group_games(recs)
    env = EnvOps(e)
    return {uid: game_record(m, env) for uid, m in games.items()}


@pytest.mark.parametrize("name", golden_io.names("ttt_"))
def test_ttt_engine_matches_reference_golden(name):
    g = golden_io.load(name)
    G = 3
    tapes = None
    if g["training"]:
        gm = np.zeros((G, 16, 16))
        un = np.zeros((G, 16, 3))
        gm[:, : g["gamma_tape"].shape[0], : g["gamma_tape"].shape[1]] = g["gamma_tape"]
        un[:, : g["unif_tape"].shape[0]] = g["unif_tape"]
        tapes = (gm, un)
    e = _engine_for(g["cfg"], g["training"], G, tapes)
    out = _play(e, [g["salt"]] * G)
    assert sorted(out) == list(range(G))
    for uid in range(G):
        golden_io.assert_record_matches(out[uid], g, check_trees=False)


@pytest.mark.parametrize("training", [False, True])
def test_ttt_engine_matches_oracle_many_games(training):
    from oracle import mcts, selfplay
    from oracle.stubnet_np import stub_forward
    from oracle.ttt import TicTacToe

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 48
    cfg["Exploration"]["epsilon_softmax_exploration"] = 0.2
    cfg["Exploration"]["epsilon_random_exploration"] = 0.2
    cfg["Exploration"]["number_of_softmax_moves"] = 1
    G = 24
    rng = np.random.Generator(np.random.Philox(7))
    gm = rng.gamma(0.15, 1.0, size=(G, 10, 9))
    un = rng.random(size=(G, 10, 3))
    e = _engine_for(cfg, training, G, (gm, un) if training else None)
    salts = list(range(100, 100 + G))
    out = _play(e, salts)
    assert len(out) == G
    for gi in range(G):
        tape = mcts.ReplayTape(gm[gi], un[gi]) if training else None
        ref = selfplay.play_game(TicTacToe(), lambda s, sl=salts[gi]: stub_forward(s, 9, sl), cfg, training, True, tape)
        got = out[gi]
        assert got["actions"] == ref["actions"], "game %d" % gi
        assert got["root_N"] == ref["root_N"]
        assert got["terminal_value"] == ref["terminal_value"]
        for m in range(ref["length"]):
            np.testing.assert_array_equal(got["child_N"][m], ref["child_N"][m])
            np.testing.assert_array_equal(got["child_W"][m], ref["child_W"][m])
            np.testing.assert_array_equal(got["child_prior"][m], ref["child_prior"][m])
            np.testing.assert_array_equal(got["states"][m], ref["states"][m])
        np.testing.assert_array_equal(np.array(got["root_W"]), np.array(ref["root_W"]))
        np.testing.assert_array_equal(np.array(got["bias"]), np.array(ref["bias"]))


def test_stubnet_kernel_matches_numpy_and_torch():
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.stubnet import DyadicStubNet, torch_reference
    from oracle.stubnet_np import stub_forward

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    G = 64
    e = SearchEngine(tic_tac_toe_spec(), cfg, G, False, policy_is_prob=True, leaf_dtype=_ffi.F32, pool_nodes=64)
    torch.manual_seed(0)
    e.leaf.copy_((torch.rand_like(e.leaf) < 0.4).float())
    salts = torch.arange(G, dtype=torch.int32) * 13
    DyadicStubNet(e, salt=salts)()
    p_t, v_t = torch_reference(e.leaf, 9, salts.to(e.device))
    assert torch.equal(p_t, e.policy) and torch.equal(v_t, e.value)
    leaf = e.leaf.cpu().numpy()
    for i in range(G):
        p, v = stub_forward(leaf[i], 9, int(salts[i]))
        np.testing.assert_array_equal(p, e.policy[i].cpu().numpy())
        assert v == e.value[i].item()


def test_ttt_env_kernels_match_oracle_random_playouts():
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import EnvOps, SearchEngine, tic_tac_toe_spec
    from oracle.ttt import TicTacToe

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    e = SearchEngine(tic_tac_toe_spec(), cfg, 1, False, pool_nodes=64)
    env = EnvOps(e)
    n = 256
    rng = np.random.default_rng(3)
    games = [TicTacToe() for _ in range(n)]
    st = env.reset(n)
    for step in range(9):
        mask = env.mask(st).cpu().numpy()
        enc = env.encode(st).cpu().numpy()
        status = env.status(st).cpu().numpy()
        acts = np.zeros(n, dtype=np.int32)
        for i, gme in enumerate(games):
            np.testing.assert_array_equal(enc[i], gme.encode()[0])
            assert status[i].tolist() == [int(gme.terminal), gme.terminal_value, gme.get_current_player(), gme.length]
            if gme.terminal:
                acts[i] = -1
                continue
            np.testing.assert_array_equal(mask[i], gme.legal_mask().astype(np.uint8))
            acts[i] = rng.choice(np.flatnonzero(mask[i]))
            gme.step(int(acts[i]))
        live = acts >= 0
        if not live.any():
            break
        idx = torch.from_numpy(np.flatnonzero(live)).to(e.device)
        sub = st[idx].contiguous()
        env.step(sub, acts[live])
        st[idx] = sub
    with pytest.raises(Exception):
        bad = env.reset(1)
        env.step(bad, [4])
        env.step(bad, [4])


def test_device_noise_generator_has_gamma_moments():
    """Throughput-mode root noise (Philox + Marsaglia-Tsang) is unpinned against numpy's stream; its
    distribution is checked instead: mean alpha*beta, variance alpha*beta^2, and numpy's quantiles."""
    import ctypes as C

    from nuzero_b200 import _ffi

    n = 1 << 20
    for alpha, beta in [(0.15, 1.0), (0.3, 1.0), (2.5, 0.5)]:
        out = torch.zeros(n, dtype=torch.float64, device="cuda")
        _ffi.check(_ffi.lib().nz_noise_probe(C.c_void_p(out.data_ptr()), n, alpha, beta, 12345, None))
        x = out.cpu().numpy()
        assert np.all(x >= 0) and np.isfinite(x).all()
        assert abs(x.mean() - alpha * beta) < 0.01 * max(alpha * beta, 0.1)
        assert abs(x.var() - alpha * beta * beta) < 0.03 * max(alpha * beta * beta, 0.1)
        ref = np.random.default_rng(0).gamma(alpha, beta, n)
        for q in (0.5, 0.9, 0.99):
            assert abs(np.quantile(x, q) - np.quantile(ref, q)) < 0.03 * max(np.quantile(ref, q), 0.05)


def test_training_mode_without_tape_plays_legal_varied_games():
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.selfplay import group_games, run_until_idle
    from nuzero_b200.stubnet import DyadicStubNet

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 32
    cfg["Exploration"]["epsilon_random_exploration"] = 0.2
    G = 256
    e = SearchEngine(tic_tac_toe_spec(), cfg, G, True, policy_is_prob=True, leaf_dtype=_ffi.F32, games_per_slot=2,
                     pool_nodes=8192, seed=7)
    run_until_idle(e, DyadicStubNet(e, salt=[3] * G))  # same network everywhere: only the noise differs
    recs, _ = e.drain_records()
    games = group_games(recs)
    assert len(games) == 2 * G
    trajectories = {tuple(m["action"] for m in g) for g in games.values()}
    assert len(trajectories) > 20, "root noise / random moves must diversify identical starting positions"
    for moves in games.values():
        seen = set()
        for m in moves:
            assert m["action"] not in seen and m["action"] in m["child_actions"]
            seen.add(m["action"])
