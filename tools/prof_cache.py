"""Where the time of a cached SCS-5 step goes: search launch, cache look-up (+ the host read of the miss count), gather,
network on the missed rows, scatter + insert.  CUDA events, mid-game."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yaml

from nuzero_b200 import _ffi
from nuzero_b200.cache import CachedForward
from nuzero_b200.engine import SearchEngine
from nuzero_b200.fastnet import FusedRecurrentForward
from nuzero_b200.games.scs_config import ScsScenario
from nuzero_b200.nets import RecurrentNet, initialize_parameters

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = yaml.safe_load(open(os.path.join(root, "nuzero_b200", "configs", "a1_search_config.yaml")))
cfg["Simulation"]["mcts_simulations"] = 200
scn = ScsScenario(os.path.join(root, "nuzero_b200", "configs", "scs", "mirrored_config_5.yml"), [None])
G = 4096
min_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
e = SearchEngine(scn.spec(), cfg, G, True, pool_nodes=131072, policy_is_prob=False, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.BF16,
                 auto_advance=True, games_per_slot=0, max_sims_per_launch=1, seed=7, arena_words=1 << 24, max_depth=256)
e.set_maps([0] * G)
e.reset()
torch.manual_seed(0)
model = RecurrentNet(scn.C, scn.planes, 256, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=True)
initialize_parameters(model)
net = CachedForward(e, lambda v: FusedRecurrentForward(v, model, 6, use_graph=True), capacity_log2=22, min_rows=min_rows)
print("bucket sizes", [v.rows for v in net.views])
for _ in range(3000):
    e.advance()
    net()
torch.cuda.synchronize()
e.arena_top.zero_()
h0, m0 = net.hits, net.misses
N = 400
t0 = time.perf_counter()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * N + 1)]
for i in range(N):
    ev[2 * i].record()
    e.advance()
    ev[2 * i + 1].record()
    net()
ev[2 * N].record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N * 1e3
adv = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(N)) / N
netms = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(N)) / N
h, m = net.hits - h0, net.misses - m0
print("min_rows %d: wall %.3f ms/step, search %.3f ms, cached network call %.3f ms (device time), hit rate %.3f, misses/step %.0f, pending/step %.0f"
      % (min_rows, wall, adv, netms, h / max(1, h + m), m / N, (h + m) / N))
