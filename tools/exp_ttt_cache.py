"""Experiment: the headline TTT workload (16384 games, 800 sims/move, stub network) with the in-kernel inference cache."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, yaml
from nuzero_b200 import _ffi
from nuzero_b200.cache import CachedForward
from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
from nuzero_b200.stubnet import DyadicStubNet

cfg = yaml.safe_load(open(os.path.join(ROOT, "nuzero_b200", "configs", "a1_search_config.yaml")))
cfg["Simulation"]["mcts_simulations"] = 800
G = 16384
for budget in [int(x) for x in (sys.argv[1:] or ["4", "16", "64"])]:
    e = SearchEngine(tic_tac_toe_spec(), cfg, G, True, pool_nodes=32768, policy_is_prob=True, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.F32,
                     auto_advance=True, games_per_slot=0, max_sims_per_launch=budget, seed=1234, arena_words=1 << 24)
    net = CachedForward(e, lambda v: DyadicStubNet(v), capacity_log2=16, min_rows=64, in_kernel=True)
    for i in range(3000):
        e.advance(); net()
        if i % 256 == 255: e.arena_top.zero_()
    torch.cuda.synchronize()
    e.raise_on_error()
    c0 = e.counters(); n = 600
    t0 = time.perf_counter()
    for i in range(n):
        e.advance(); net()
        if i % 64 == 63: e.arena_top.zero_()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c1 = e.counters()
    print("budget", budget, "sims/s %.3e" % ((c1["sims"] - c0["sims"]) / dt), "us/launch pair %.1f" % (dt / n * 1e6), "sims/launch %.0f" % ((c1["sims"] - c0["sims"]) / n),
          "hit rate %.4f" % net.hit_rate(), "misses/launch %.1f" % (net.misses / net.calls))
    e.close(); del e, net
    torch.cuda.empty_cache()
