"""GPU: BASELINE.json's full sizes.  The oracle cannot replay 10^8 simulations, so whole-batch results
are checked through size-independent properties (visit-count conservation, carried sub-tree counts,
arg-max choice, legality and outcome under an independent replay of every trajectory), and a sample
of the same batch is compared with the oracle bit for bit."""
import os

import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu


def _group(recs):
    games = {}
    for r in recs:
        games.setdefault(r["uid"], []).append(r)
    for m in games.values():
        m.sort(key=lambda r: r["move"])
    return games


def _check_search_invariants(moves, sims, training):
    L = len(moves)
    assert [m["move"] for m in moves] == list(range(L)) and moves[-1]["game_end"] and moves[-1]["length"] == L
    for i, m in enumerate(moves):
        n = m["child_N"]
        assert int(n.sum()) == m["root_N"] - 1, "every visit but the expanding one went through a child"
        carried = 0 if i == 0 else int(moves[i - 1]["child_N"][moves[i - 1]["child"]])
        assert m["root_N"] == sims + carried, "keep_subtree carries the chosen child's visits"
        assert m["child_actions"][m["child"]] == m["action"]
        assert np.all(np.diff(m["child_actions"]) > 0), "children in ascending action order"
        if not training:
            assert m["child"] == int(np.argmax(n)), "max_action: first maximum"
        assert abs(m["root_W"]) <= m["root_N"] + 1e-9


def test_ttt_full_size_16384_games_800_sims():
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.selfplay import run_until_idle
    from nuzero_b200.stubnet import DyadicStubNet
    from oracle import selfplay
    from oracle.stubnet_np import stub_forward
    from oracle.ttt import TicTacToe, _LINES

    cfg = golden_io.load("ttt_p0_s800_salt3")["cfg"]
    G, sims = 16384, 800
    e = SearchEngine(tic_tac_toe_spec(), cfg, G, False, policy_is_prob=True, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.F32,
                     auto_advance=True, games_per_slot=1, max_sims_per_launch=4, pool_nodes=32768, arena_words=1 << 23)
    salts = torch.arange(G, dtype=torch.int32) * 7 + 1
    run_until_idle(e, DyadicStubNet(e, salt=salts), max_launches=200000, check_every=256)
    # the same records through the device-resident replay buffer (nz_replay_decode), before the host drains them
    from nuzero_b200.replay import DeviceReplayBuffer

    top = e.arena_top.cpu()
    words, offs = e.arena[: int(top[0])].clone(), e.rec_index[: int(top[2])].to(torch.int64)
    recs, dropped = e.drain_records()
    assert dropped == 0
    games = _group(recs)
    assert sorted(games) == list(range(G))
    drb = DeviceReplayBuffer(e, G, 2048, capacity=9 * G)
    assert drb.ingest_words(words, offs) == len(recs) == drb.len() and drb.played_games() == G
    rows = drb._rows()
    st, pol, val, uid = drb.states[rows], drb.policy[rows], drb.value[rows], drb.uid[rows]
    stones = st.sum(dim=(1, 2, 3)).to(torch.int64)              # position after k moves shows k stones
    assert torch.allclose(pol.sum(1), torch.ones_like(val), atol=1e-6)
    assert bool(((pol > 0).sum(1) <= 9 - stones).all()), "policy mass only on empty cells"
    assert bool((pol * st.sum(1).reshape(-1, 9) == 0).all()), "no policy mass on occupied cells"
    first = torch.ones_like(uid, dtype=torch.bool)
    first[1:] = uid[1:] != uid[:-1]
    assert int(first.sum()) == G and bool((stones[first] == 0).all()), "games are contiguous and start from the empty board"
    assert bool((stones[~first] == stones[torch.nonzero(~first)[:, 0] - 1] + 1).all()), "one stone per move, in order"
    tv_by_uid = torch.tensor([games[u][-1]["terminal_value"] for u in range(G)], dtype=torch.float32, device=val.device)
    assert torch.equal(val, tv_by_uid[uid]), "every position carries its game's terminal value (make_target)"
    c = e.counters()
    assert c["games"] == G and c["sims"] == sims * c["moves"] == sims * len(recs)
    values = np.zeros(3, dtype=np.int64)
    for uid, moves in games.items():
        _check_search_invariants(moves, sims, training=False)
        # independent replay of the trajectory on a plain bit board
        stones, tv = [0, 0], 0
        for i, m in enumerate(moves):
            occ = stones[0] | stones[1]
            assert m["player"] == i % 2 + 1 and int(m["state"][0]) & 0x3FFFF == stones[0] | (stones[1] << 9)
            assert m["child_actions"].tolist() == [a for a in range(9) if not (occ >> a) & 1], "children = legal moves"
            stones[i % 2] |= 1 << m["action"]
        if any(stones[0] & l == l for l in _LINES):
            tv = 1
        elif any(stones[1] & l == l for l in _LINES):
            tv = -1
        assert moves[-1]["terminal_value"] == tv and (tv != 0 or len(moves) == 9 or True)
        values[tv + 1] += 1
    assert values.sum() == G and values.min() >= 0
    # a sample of the very same batch against the oracle, bit for bit
    for uid in (0, 1, 4097, 16383):
        ref = selfplay.play_game(TicTacToe(), lambda s, sl=int(salts[uid]): stub_forward(s, 9, sl), cfg, False, True,
                                 keep_states=False)
        moves = games[uid]
        assert [m["action"] for m in moves] == ref["actions"] and [m["root_N"] for m in moves] == ref["root_N"]
        for m, n, w in zip(moves, ref["child_N"], ref["root_W"]):
            np.testing.assert_array_equal(m["child_N"], n)
            assert m["root_W"] == w


def test_scs_full_size_4096_games_200_sims():
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.selfplay import run_until_idle
    from nuzero_b200.stubnet import DyadicStubNet
    from oracle import scs as oscs

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 200
    path = os.path.join(golden_io.SCS_CONFIGS, "randomized_config_5.yml")
    seeds = list(range(1, 17))
    scn = ScsScenario(path, seeds)
    G = 4096
    # 4096 x 400 000 nodes x 32 B = 52 GB of node pools + 26-30 GB of per-run game states (node_state_cache)
    e = SearchEngine(scn.spec(), cfg, G, True, policy_is_prob=True, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.F32,
                     auto_advance=True, games_per_slot=1, max_sims_per_launch=4, pool_nodes=400000, max_depth=200,
                     arena_words=1 << 25, seed=5)
    maps = [i % len(seeds) for i in range(G)]
    e.set_maps(maps)
    e.reset()
    run_until_idle(e, DyadicStubNet(e, salt=torch.arange(G, dtype=torch.int32)), max_launches=400000, check_every=512)
    recs, dropped = e.drain_records()
    assert dropped == 0
    games = _group(recs)
    assert sorted(games) == list(range(G))
    scen = {s: oscs.load_scenario(path, s) for s in seeds}
    lengths = []
    for uid, moves in games.items():
        _check_search_invariants(moves, 200, training=True)
        lengths.append(len(moves))
        if uid % 8:  # replay one game in eight through the oracle's rules (no search): legality, players, outcome
            continue
        g = oscs.SCS(scen[seeds[maps[moves[0]["slot"]]]])
        for m in moves:
            assert not g.is_terminal() and m["player"] == g.get_current_player()
            assert m["child_actions"].tolist() == np.flatnonzero(g.legal_mask()).tolist(), "children = legal actions"
            g.step(m["action"])
        assert g.is_terminal() and g.get_terminal_value() == moves[-1]["terminal_value"] and g.length == len(moves)
    assert min(lengths) >= 20 and max(lengths) <= 400
