import sys, time, cProfile, pstats
sys.path.insert(0,'/root/repo')
import torch, yaml
from nuzero_b200 import _ffi
from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
from nuzero_b200.stubnet import DyadicStubNet
from nuzero_b200.replay import DeviceReplayBuffer
from nuzero_b200.selfplay import SelfPlayRunner
cfg = yaml.safe_load(open('/root/repo/nuzero_b200/configs/a1_search_config.yaml')); cfg["Simulation"]["mcts_simulations"]=800
e = SearchEngine(tic_tac_toe_spec(), cfg, 16384, True, pool_nodes=32768, policy_is_prob=True, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.F32, auto_advance=True, games_per_slot=0, max_sims_per_launch=1, seed=1, arena_words=1<<24)
net = DyadicStubNet(e, uid_mul=1)
for _ in range(3000): e.advance(); net()
torch.cuda.synchronize(); e.arena_top.zero_()
drb = DeviceReplayBuffer(e, 400000, 2048, 400000*9+9, drop_incomplete=True)
r = SelfPlayRunner(e, net, drb, launches_per_step=256)
for _ in range(5): r.step()
r.flush()
pr = cProfile.Profile(); pr.enable()
t0=time.perf_counter()
for _ in range(20): r.step()
r.flush(); torch.cuda.synchronize()
dt=time.perf_counter()-t0
pr.disable()
print("ms/step", dt/20*1e3)
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
