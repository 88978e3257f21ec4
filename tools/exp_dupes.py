"""Experiment: how many of the leaves one launch hands to the network are duplicates of each other (same compact state)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, yaml
from nuzero_b200 import _ffi
from nuzero_b200.cache import CachedForward
from nuzero_b200.engine import SearchEngine
from nuzero_b200.fastnet import FusedRecurrentForward
from nuzero_b200.games.scs_config import ScsScenario
from nuzero_b200.nets import RecurrentNet, initialize_parameters

cfg = yaml.safe_load(open(os.path.join(ROOT, "nuzero_b200", "configs", "a1_search_config.yaml")))
cfg["Simulation"]["mcts_simulations"] = 200
scn = ScsScenario(os.path.join(ROOT, "nuzero_b200", "configs", "scs", "mirrored_config_5.yml"), [None])
G = 4096
e = SearchEngine(scn.spec(), cfg, G, True, pool_nodes=131072, policy_is_prob=False, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.BF16,
                 auto_advance=True, games_per_slot=1, max_sims_per_launch=16, seed=7, arena_words=1 << 24, max_depth=256)
torch.manual_seed(0)
model = RecurrentNet(scn.C, scn.planes, 256, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=True)
initialize_parameters(model)
net = CachedForward(e, lambda v: FusedRecurrentForward(v, model, 6, use_graph=True), capacity_log2=22, min_rows=512, in_kernel=True, miss_target=1024)
tot = uniq = calls = 0
hist = []
t0 = time.time()
for it in range(100000):
    e.advance()
    n = int(e.dense_count[0])
    if n:
        slots = e.dense_rows[:n].long()
        keys = e.gstate[slots, 1]
        u = torch.unique(keys, dim=0).shape[0]
        tot += n; uniq += u; calls += 1
        if calls % 400 == 0:
            hist.append((calls, tot, uniq))
    net()
    e.arena_top.zero_()
    if it % 64 == 0 and bool((e.phases() == _ffi.PHASE_IDLE).all()):
        break
print("launches", it, "rows", tot, "unique within their batch", uniq, "dup fraction %.3f" % (1 - uniq / max(1, tot)), "hit rate", net.hit_rate(), "sec", time.time() - t0)
print(hist)
