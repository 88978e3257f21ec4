import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import golden_io, torch
from nuzero_b200 import _ffi
from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
from nuzero_b200.stubnet import DyadicStubNet
cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
cfg["Simulation"]["mcts_simulations"] = 100
e = SearchEngine(tic_tac_toe_spec(), cfg, 1, False, policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                 auto_advance=True, games_per_slot=1, record_detail=True, pool_nodes=40000, max_sims_per_launch=8, compact=False)
net = DyadicStubNet(e, salt=[2])
def dump(tag):
    torch.cuda.synchronize()
    c = e.ctl[0].tolist()
    print(tag, "ctl", c[:16])
    for i in list(range(0, 11)) + [34, 35, 36]:
        print("   node", i, "prior %.4f W %.3f N %d act %d base %d K %d" % (float(e.node_prior[0, i]), float(e.node_W[0, i]), int(e.node_N[0, i]),
              int(e.node_flags[0, i]) >> 16, int(e.node_base[0, i]), int(e.node_K[0, i])))
prev = 0
for it in range(200):
    e.advance(); net()
    mv = int(e.ctl[0, _ffi.CTL_MOVE])
    sd = int(e.ctl[0, _ffi.CTL_SIMS_DONE])
    if it < 3 or mv != prev or (mv == 0 and sd >= 92):
        dump("launch %d move %d" % (it, mv))
    if mv >= 2: break
    prev = mv
