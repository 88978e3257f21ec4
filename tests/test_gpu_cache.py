"""GPU: the device inference cache (nz_cache_lookup / nz_cache_insert, CachedForward) in both forms — consulted by a kernel of
its own after the search launch, and consulted INSIDE the search kernel (nz_engine_attach_cache: the reference's order,
Explorer.py:146-155; hits are expanded within the launch, the missed leaves form a dense batch).  The search must be
bit-identical with and without it — a hit returns exactly what the network returned for that state — and it must actually hit."""
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


def _play(e, net):
    from nuzero_b200.selfplay import group_games, run_until_idle

    run_until_idle(e, net)
    if hasattr(net, "drain"):
        net.drain()
    recs, dropped = e.drain_records()
    assert dropped == 0
    return group_games(recs)


def _same(a, b):
    assert sorted(a) == sorted(b)
    for uid in a:
        assert [m["action"] for m in a[uid]] == [m["action"] for m in b[uid]]
        assert [m["root_N"] for m in a[uid]] == [m["root_N"] for m in b[uid]]
        for x, y in zip(a[uid], b[uid]):
            assert np.array_equal(x["child_N"], y["child_N"]) and x["root_W"] == y["root_W"]


@pytest.mark.parametrize("capacity_log2,in_kernel,budget,pipeline", [(16, False, 2, False), (5, False, 2, False), (16, True, 2, False),
                                                                     (16, True, 12, False), (5, True, 6, False), (16, True, 8, True),
                                                                     (5, True, 3, True)])
def test_ttt_real_network_search_is_identical_with_the_cache(capacity_log2, in_kernel, budget, pipeline):
    from nuzero_b200 import _ffi
    from nuzero_b200.cache import CachedForward
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.fastnet import FusedRecurrentForward
    from nuzero_b200.nets import RecurrentNet, initialize_parameters

    torch.manual_seed(0)
    model = RecurrentNet(2, 1, 64, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=False)
    initialize_parameters(model)
    out = []
    for cached in (False, True):
        e = SearchEngine(tic_tac_toe_spec(), _cfg(60), 96, True, policy_is_prob=False, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.BF16,
                         auto_advance=True, games_per_slot=2, pool_nodes=4000, seed=9, max_sims_per_launch=budget if cached else 2,
                         record_detail=True)
        if cached:
            net = CachedForward(e, lambda v: FusedRecurrentForward(v, model, 2, use_graph=True), capacity_log2=capacity_log2, min_rows=32,
                                in_kernel=in_kernel, pipeline=pipeline)  # pipeline: the network call of launch k runs beside launch k + 1
        else:
            net = FusedRecurrentForward(e, model, 2, use_graph=True)
        out.append(_play(e, net))
        if cached and capacity_log2 >= 16:
            assert net.hit_rate() > 0.6, net.hit_rate()   # warming up: Tic-Tac-Toe has 5478 reachable positions
        if cached and capacity_log2 == 5:
            assert 0.0 < net.hit_rate() < 0.9            # 32 slots: most states do not fit, the results must still agree
        if cached and in_kernel and not pipeline:  # (in the two-lane pipeline a game plays in every other launch)
            assert e.launches < launches_plain, (e.launches, launches_plain)  # hits do not wait for a launch of their own
        launches_plain = e.launches
    _same(out[0], out[1])


@pytest.mark.parametrize("in_kernel,pipeline", [(False, False), (True, False), (True, True)])
def test_scs_search_is_identical_with_the_cache_and_maps_do_not_alias(in_kernel, pipeline):
    from nuzero_b200 import _ffi
    from nuzero_b200.cache import CachedForward
    from nuzero_b200.engine import SearchEngine
    from nuzero_b200.fastnet import FusedRecurrentForward
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.nets import RecurrentNet, initialize_parameters

    scn = ScsScenario(os.path.join(golden_io.SCS_CONFIGS, "randomized_config_5.yml"), [1, 2, 3])
    torch.manual_seed(0)
    model = RecurrentNet(scn.C, scn.planes, 64, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=True)
    initialize_parameters(model)
    out = []
    for cached in (False, True):
        e = SearchEngine(scn.spec(), _cfg(16), 24, False, policy_is_prob=False, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.BF16,
                         auto_advance=True, games_per_slot=1, pool_nodes=30000, max_depth=128, max_sims_per_launch=8 if cached and in_kernel else 2)
        e.set_maps([g % 3 for g in range(24)])  # 8 games per map: identical games on the same map, different ones across maps
        e.reset()
        if cached:
            net = CachedForward(e, lambda v: FusedRecurrentForward(v, model, 2, use_graph=True), capacity_log2=18, min_rows=8, in_kernel=in_kernel,
                                pipeline=pipeline)
        else:
            net = FusedRecurrentForward(e, model, 2, use_graph=True)
        out.append(_play(e, net))
        if cached:
            # the 8 games of a map run in lock-step, so a new state misses in all 8 rows of the SAME batch (duplicates inside a
            # batch are evaluated, not shared); the hits are the temporal re-visits
            assert net.hit_rate() > 0.1, net.hit_rate()
    _same(out[0], out[1])
    acts = {uid: tuple(m["action"] for m in moves) for uid, moves in out[1].items()}
    assert len({acts[g] for g in range(0, 24, 3)}) == 1, "same map, same deterministic game"


@pytest.mark.parametrize("game,G,capacity_log2", [("ttt", 2048, 14), ("ttt", 2048, 8), ("scs", 192, 16)])
def test_in_kernel_cache_under_contention_matches_the_plain_search(game, G, capacity_log2):
    """Many games, few distinct states, a cheap deterministic network (the stub): every launch sees hits, claims, shared rows
    and published expansions racing in the same table (Tic-Tac-Toe: four games per warp).  The records must equal the plain
    search's, with a table that holds everything and with one that is far too small."""
    from nuzero_b200 import _ffi
    from nuzero_b200.cache import CachedForward
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.stubnet import DyadicStubNet

    if game == "ttt":
        spec, kw, sims, maps = tic_tac_toe_spec(), dict(pool_nodes=4000), 40, None
    else:
        scn = ScsScenario(os.path.join(golden_io.SCS_CONFIGS, "randomized_config_5.yml"), [1, 2])
        spec, kw, sims, maps = scn.spec(), dict(pool_nodes=30000, max_depth=128), 12, [g % 2 for g in range(G)]
    out = []
    for cached in (False, True):
        e = SearchEngine(spec, _cfg(sims), G, True, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32, auto_advance=True,
                         games_per_slot=2 if game == "ttt" else 1, seed=21, max_sims_per_launch=8 if cached else 2, record_detail=True, **kw)
        if maps is not None:
            e.set_maps(maps)
            e.reset()
        if cached:
            net = CachedForward(e, lambda v: DyadicStubNet(v), capacity_log2=capacity_log2, min_rows=64, in_kernel=True)
        else:
            net = DyadicStubNet(e)
        out.append(_play(e, net))
        if cached:
            c = e.counters()
            assert c["cache_hits"] > 0 and c["cache_shared"] > 0, c
    _same(out[0], out[1])
