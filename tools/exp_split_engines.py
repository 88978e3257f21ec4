"""Experiment: the 16384 concurrent games as K independent engines on K streams inside one CUDA graph, so that one
engine's network kernel and launch gaps overlap with another engine's search kernel.  Results are unaffected (games are
independent); prints simulations/s for K = 1, 2, 4."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yaml

from nuzero_b200 import _ffi
from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
from nuzero_b200.stubnet import DyadicStubNet

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = yaml.safe_load(open(os.path.join(root, "nuzero_b200", "configs", "a1_search_config.yaml")))
cfg["Simulation"]["mcts_simulations"] = 800
G, INNER, STEPS = 16384, 256, 20
for K in (1, 2, 4):
    engines = [SearchEngine(tic_tac_toe_spec(), cfg, G // K, True, pool_nodes=32768, policy_is_prob=True, leaf_dtype=_ffi.BF16,
                            policy_dtype=_ffi.F32, auto_advance=True, games_per_slot=0, max_sims_per_launch=1, seed=1 + k,
                            arena_words=1 << 22) for k in range(K)]
    nets = [DyadicStubNet(e, uid_mul=1) for e in engines]
    streams = [torch.cuda.Stream() for _ in range(K)]

    def pairs(n):
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for i in range(n):
            for e, net, s in zip(engines, nets, streams):
                with torch.cuda.stream(s):
                    e.advance()
                    net()
            if i % 2048 == 2047:
                for e, s in zip(engines, streams):
                    with torch.cuda.stream(s):
                        e.arena_top.zero_()
        for s in streams:
            cur.wait_stream(s)

    warm = torch.cuda.Stream()
    with torch.cuda.stream(warm):
        pairs(20000)
    torch.cuda.synchronize()
    for e in engines:
        e.arena_top.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        pairs(INNER)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    c0 = sum(e.counters()["sims"] for e in engines)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(STEPS):
        g.replay()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    c1 = sum(e.counters()["sims"] for e in engines)
    for e in engines:
        e.raise_on_error()
    print("K=%d engines x %d games: %.3e sims/s, %.2f us per launch pair of all engines" % (K, G // K, (c1 - c0) / ms * 1e3, ms * 1e3 / STEPS / INNER))
    del g, engines, nets
    torch.cuda.empty_cache()
