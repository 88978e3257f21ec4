"""GPU: the device-resident replay buffer (nz_replay_decode + DeviceReplayBuffer) against the list-based restatement of
Training/ReplayBuffer.py fed with the same games through the reference-shaped objects (FinishedGame): entries, order,
window behaviour, slices.  Exact on states, policy targets (float32 of the reference's float64 fractions) and values."""
import os

import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu


def _list_buffer_from_records(e, recs, window):
    """The reference path: records -> per-game objects -> ReplayBuffer.save_game, in the order the games finished."""
    from nuzero_b200.engine import EnvOps
    from nuzero_b200.gamer import finished_game_of
    from nuzero_b200.replay import ReplayBuffer
    from nuzero_b200.selfplay import game_record, group_games

    games = group_games(recs)
    order = [r["uid"] for r in recs if r["game_end"] and r["uid"] in games]
    env = EnvOps(e)
    rb = ReplayBuffer(window, 8)
    for uid in order:
        moves = games[uid]
        rec = game_record(moves, env, None if e.spec.kind == 0 else moves[0]["map"])
        rb.save_game(finished_game_of(rec, e.A), 7)
    return rb, order


def _assert_same(drb, rb):
    assert drb.len() == rb.len()
    got = drb.get_buffer()
    want = rb.get_buffer()
    for (s0, (v0, p0), g0), (s1, (v1, p1), g1) in zip(got, want):
        assert torch.equal(s0, torch.as_tensor(s1).float())
        assert v0 == float(v1)
        assert p0 == torch.tensor(p1, dtype=torch.float32).tolist()
        assert g0 == g1


@pytest.mark.parametrize("incremental", [False, True])
def test_ttt_device_replay_matches_list_replay(incremental):
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, parse_records, tic_tac_toe_spec
    from nuzero_b200.replay import DeviceReplayBuffer
    from nuzero_b200.stubnet import DyadicStubNet

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    G, per_slot, window = 48, 3, 100
    e = SearchEngine(tic_tac_toe_spec(), cfg, G, True, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                     auto_advance=True, games_per_slot=per_slot, pool_nodes=4000, seed=5, max_sims_per_launch=4)
    net = DyadicStubNet(e, uid_mul=1)
    drb = DeviceReplayBuffer(e, window, 8, capacity=window * 9, game_index=7)
    all_words = []
    it = 0
    while True:
        e.advance()
        net()
        it += 1
        if it % 16 == 0:
            done = bool((e.phases() == _ffi.PHASE_IDLE).all())
            if incremental or done:
                top = e.arena_top.cpu()
                all_words.append(e.arena[: int(top[0])].cpu().numpy().view(np.uint32).copy())
                drb.ingest()
            if done:
                break
    e.raise_on_error()
    recs = []
    for w in all_words:
        recs += parse_records(w, e.state_words)
    rb, order = _list_buffer_from_records(e, recs, window)
    assert len(order) == G * per_slot and drb.played_games() == rb.played_games() == window
    assert drb.pend_hdr.shape[0] == 0 and drb.pend_words.numel() == 0
    _assert_same(drb, rb)
    # slices and tensors
    st, v, p, g = drb.get_slice_tensors(5, 40)
    want = rb.get_slice(5, 40)
    assert st.shape[0] == len(want) == 35 and st.is_cuda
    assert torch.equal(st.cpu(), torch.cat([torch.as_tensor(w[0]).float() for w in want]))
    assert torch.equal(v.cpu(), torch.tensor([float(w[1][0]) for w in want]))
    assert torch.allclose(p.sum(1), torch.ones_like(p.sum(1)), atol=1e-6)
    # sampling returns rows of the window; shuffle keeps the multiset of rows
    before = sorted((float(a), tuple(b.tolist())) for a, b in zip(drb.value[drb._rows()].cpu(), drb.policy[drb._rows()].cpu()))
    s2 = drb.get_sample_tensors(64, True)
    assert s2[0].shape[0] == 64
    drb.shuffle()
    after = sorted((float(a), tuple(b.tolist())) for a, b in zip(drb.value[drb._rows()].cpu(), drb.policy[drb._rows()].cpu()))
    assert before == after


def test_scs_device_replay_matches_list_replay():
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, parse_records
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.replay import DeviceReplayBuffer
    from nuzero_b200.stubnet import DyadicStubNet

    cfg = dict(golden_io.load("ttt_p0_s25_salt0")["cfg"])
    cfg = {k: dict(v) if isinstance(v, dict) else v for k, v in cfg.items()}
    cfg["Simulation"]["mcts_simulations"] = 12
    scn = ScsScenario(os.path.join(golden_io.SCS_CONFIGS, "randomized_config_5.yml"), [1, 2, 3])
    G = 12
    e = SearchEngine(scn.spec(), cfg, G, False, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                     auto_advance=True, games_per_slot=1, pool_nodes=60000, max_depth=256, max_sims_per_launch=4)
    e.set_maps([i % 3 for i in range(G)])
    e.reset()
    net = DyadicStubNet(e, uid_mul=1)
    drb = DeviceReplayBuffer(e, 1000, 8, capacity=G * 160, game_index=7)
    words = []
    for it in range(200000):
        e.advance()
        net()
        if (it + 1) % 256 == 0:
            top = e.arena_top.cpu()
            words.append(e.arena[: int(top[0])].cpu().numpy().view(np.uint32).copy())
            drb.ingest()
            if bool((e.phases() == _ffi.PHASE_IDLE).all()):
                break
    e.raise_on_error()
    recs = []
    for w in words:
        recs += parse_records(w, e.state_words)
    rb, order = _list_buffer_from_records(e, recs, 1000)
    assert len(order) == G and drb.len() == rb.len() > 0
    _assert_same(drb, rb)


def test_selfplay_runner_pipelined_collect_matches_one_shot():
    """SelfPlayRunner (records of step i read on a side stream while step i+1 searches; append-only arena with resets)
    delivers exactly the games a synchronous drain delivers."""
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.replay import DeviceReplayBuffer
    from nuzero_b200.selfplay import SelfPlayRunner
    from nuzero_b200.stubnet import DyadicStubNet

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    G, per_slot = 64, 4
    out = []
    for pipelined in (False, True):
        e = SearchEngine(tic_tac_toe_spec(), cfg, G, True, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                         auto_advance=True, games_per_slot=per_slot, pool_nodes=4000, seed=11, max_sims_per_launch=2,
                         arena_words=1 << 13)  # small arena: forces resets of the append-only arena
        net = DyadicStubNet(e, uid_mul=1)
        drb = DeviceReplayBuffer(e, 10000, 8, capacity=G * per_slot * 9)
        runner = SelfPlayRunner(e, net, drb, launches_per_step=8, use_graph=pipelined)
        for _ in range(4000):
            if pipelined:
                runner.step()
            else:
                runner.play()
                runner.collect()
            if bool((e.phases() == _ffi.PHASE_IDLE).all()):
                break
        runner.flush()
        e.raise_on_error()
        assert drb.played_games() == G * per_slot and drb.pend_hdr.shape[0] == 0
        rows = drb._rows()
        key = torch.argsort(drb.uid[rows] * 16 + torch.arange(rows.numel(), device=rows.device) % 1, stable=True)
        out.append((drb.uid[rows][key].cpu(), drb.states[rows][key].cpu(), drb.policy[rows][key].cpu(), drb.value[rows][key].cpu()))
    for a, b in zip(out[0], out[1]):
        assert torch.equal(a, b)


def test_ingest_parts_equals_one_ingest_per_part():
    """The multi-rank path (all ranks' records in one pass) stores what ingesting rank after rank stores."""
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.replay import DeviceReplayBuffer
    from nuzero_b200.selfplay import run_until_idle
    from nuzero_b200.stubnet import DyadicStubNet

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    parts = []
    for rank in range(3):
        e = SearchEngine(tic_tac_toe_spec(), cfg, 16, True, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                         auto_advance=True, games_per_slot=2, pool_nodes=4000, seed=100 + rank)
        run_until_idle(e, DyadicStubNet(e, uid_mul=1))
        top = e.arena_top.cpu()
        parts.append((e.arena[: int(top[0])].clone(), e.rec_index[: int(top[2])].to(torch.int64)))
    a = DeviceReplayBuffer(e, 1000, 8, capacity=3 * 32 * 9)
    b = DeviceReplayBuffer(e, 1000, 8, capacity=3 * 32 * 9)
    n_a = a.ingest_parts(parts, uid_mul=3)
    n_b = sum(b.ingest_words(w, o, uid_mul=3, uid_add=r) for r, (w, o) in enumerate(parts))
    assert n_a == n_b == a.len() == b.len() and a.played_games() == b.played_games() == 96
    ra, rb = a._rows(), b._rows()
    assert sorted(a.uid[ra].cpu().tolist()) == sorted(b.uid[rb].cpu().tolist())
    assert len(set(a.uid[ra].cpu().tolist())) == 96
    ka, kb = torch.argsort(a.uid[ra], stable=True), torch.argsort(b.uid[rb], stable=True)
    for x, y in ((a.states, b.states), (a.policy, b.policy), (a.value, b.value)):
        assert torch.equal(x[ra][ka], y[rb][kb])


@pytest.mark.parametrize("game", ["ttt", "scs"])
def test_replay_decoder_against_the_oracle(game):
    """States, policy targets and value targets of the device window (and of its pinned-host mirror) against the ORACLE's
    self-play of the same games — the reference's save_game input (ReplayBuffer.py:31-33: state_history[i], make_target(i))
    computed without any kernel of this package."""
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.replay import DeviceReplayBuffer
    from nuzero_b200.selfplay import run_until_idle
    from nuzero_b200.stubnet import DyadicStubNet
    from oracle import scs as oscs
    from oracle import selfplay as oselfplay
    from oracle.stubnet_np import stub_forward
    from oracle.ttt import TicTacToe

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    if game == "ttt":
        G, A, sims = 6, 9, 30
        spec, kw, new_game = tic_tac_toe_spec(), dict(pool_nodes=8000), TicTacToe
    else:
        G, sims = 3, 6
        path = os.path.join(golden_io.SCS_CONFIGS, "mirrored_config_5.yml")
        scn = ScsScenario(path, [None])
        sc = oscs.load_scenario(path, None)
        A = sc.A
        spec, kw, new_game = scn.spec(), dict(pool_nodes=30000, max_depth=128), lambda: oscs.SCS(sc)
    cfg["Simulation"]["mcts_simulations"] = sims
    salts = [11 + g for g in range(G)]
    e = SearchEngine(spec, cfg, G, False, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32, auto_advance=True,
                     games_per_slot=1, max_sims_per_launch=4, **kw)
    run_until_idle(e, DyadicStubNet(e, salt=salts), max_launches=200000)
    drb = DeviceReplayBuffer(e, 100, 8, capacity=G * 400, game_index=3, host_mirror=True)
    n_pos = drb.ingest()
    rows = drb._rows()
    uid = drb.uid[rows].cpu().numpy()
    h_st, h_v, h_p, h_g = drb.host_arrays()
    total = 0
    for g in range(G):
        ref = oselfplay.play_game(new_game(), lambda s, sl=salts[g]: stub_forward(s, A, sl), cfg, False, True)
        pick = np.nonzero(uid == g)[0]
        sel = rows[torch.from_numpy(pick).to(rows.device)]
        want_states = np.stack(ref["states"]).astype(np.float32)
        want_policy = oselfplay.policy_targets(ref, A).astype(np.float32)
        assert len(pick) == ref["length"]
        np.testing.assert_array_equal(drb.states[sel].cpu().numpy().reshape(want_states.shape), want_states)
        np.testing.assert_array_equal(drb.policy[sel].cpu().numpy(), want_policy)
        assert bool((drb.value[sel].cpu() == float(ref["terminal_value"])).all())
        np.testing.assert_array_equal(h_st[pick].reshape(want_states.shape), want_states)
        np.testing.assert_array_equal(h_p[pick], want_policy)
        assert bool((h_v[pick] == float(ref["terminal_value"])).all()) and bool((h_g[pick] == 3).all())
        total += ref["length"]
    assert n_pos == total
