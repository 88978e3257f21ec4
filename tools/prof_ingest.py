"""Where the time of SelfPlayRunner.step() goes (TTT headline workload): host phases and the duration of each CUDA-graph
replay with and without concurrent replay-side work."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yaml

from nuzero_b200 import _ffi
from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
from nuzero_b200.replay import DeviceReplayBuffer
from nuzero_b200.selfplay import SelfPlayRunner
from nuzero_b200.stubnet import DyadicStubNet

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = yaml.safe_load(open(os.path.join(root, "nuzero_b200", "configs", "a1_search_config.yaml")))
cfg["Simulation"]["mcts_simulations"] = 800
e = SearchEngine(tic_tac_toe_spec(), cfg, 16384, True, pool_nodes=32768, policy_is_prob=True, leaf_dtype=_ffi.BF16,
                 policy_dtype=_ffi.F32, auto_advance=True, games_per_slot=0, max_sims_per_launch=1, seed=1, arena_words=1 << 24)
net = DyadicStubNet(e, uid_mul=1)
for _ in range(3000):
    e.advance()
    net()
torch.cuda.synchronize()
e.arena_top.zero_()
drb = DeviceReplayBuffer(e, 400000, 2048, 400000 * 9 + 9, drop_incomplete=True)
r = SelfPlayRunner(e, net, drb, launches_per_step=256)
for _ in range(5):
    r.step()
r.flush()
N = 20


def timed(label, body):
    evs = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        body(a, b)
        evs.append((a, b))
    r.flush()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / N * 1e3
    g = sorted(a.elapsed_time(b) for a, b in evs)
    print("%-28s wall %.2f ms/step   graph replay median %.2f  max %.2f ms" % (label, dt, g[N // 2], g[-1]))


def play_only(a, b):
    a.record(); r.play(); b.record()


def full_step(a, b):
    prev = r._snap
    a.record(); r.play(); b.record()
    r._snap = r._snapshot()
    if prev is not None:
        r._collect(prev)


timed("play only", play_only)
r.collect()
timed("step (pipelined ingest)", full_step)
timed("play only", play_only)

# host timeline of the pipelined step
import collections
acc = collections.Counter()
torch.cuda.synchronize()
T0 = time.perf_counter()
for _ in range(N):
    t = time.perf_counter()
    prev = r._snap
    r.play()
    t1 = time.perf_counter(); acc["replay()"] += t1 - t
    r._snap = r._snapshot()
    t2 = time.perf_counter(); acc["snapshot"] += t2 - t1
    if prev is not None:
        prev[1].synchronize()
        t3 = time.perf_counter(); acc["wait prev graph"] += t3 - t2
        r._collect(prev)
        acc["collect body"] += time.perf_counter() - t3
torch.cuda.synchronize()
tot = time.perf_counter() - T0
print("pipelined: %.2f ms/step; " % (tot / N * 1e3) + ", ".join("%s %.2f" % (k, v / N * 1e3) for k, v in acc.items()))


def step_and_sample(a, b):
    a.record(); b.record()
    r.step()
    with torch.cuda.stream(r.side):
        st_b, v_b, p_b, _g = drb.get_sample_tensors(256, True)
        float(v_b.sum().cpu())


timed("step + sample read-back", step_and_sample)
acc = collections.Counter()
torch.cuda.synchronize()
T0 = time.perf_counter()
for _ in range(N):
    t = time.perf_counter()
    r.step()
    t1 = time.perf_counter(); acc["step()"] += t1 - t
    with torch.cuda.stream(r.side):
        st_b, v_b, p_b, _g = drb.get_sample_tensors(256, True)
        t2 = time.perf_counter(); acc["sample enqueue"] += t2 - t1
        float(v_b.sum().cpu())
        acc["sample sync"] += time.perf_counter() - t2
torch.cuda.synchronize()
print("step+sample: %.2f ms/step; " % ((time.perf_counter() - T0) / N * 1e3) + ", ".join("%s %.2f" % (k, v / N * 1e3) for k, v in acc.items()))
