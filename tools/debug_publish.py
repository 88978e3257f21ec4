"""Debug: verify the published expansion lists of the in-kernel cache against a host recomputation (TTT, real network)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_io
from nuzero_b200 import _ffi
from nuzero_b200.cache import CachedForward
from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
from nuzero_b200.fastnet import FusedRecurrentForward
from nuzero_b200.nets import RecurrentNet, initialize_parameters
from nuzero_b200.selfplay import group_games, run_until_idle

cfg = {k: dict(v) if isinstance(v, dict) else v for k, v in golden_io.load("ttt_p0_s25_salt0")["cfg"].items()}
cfg["Simulation"]["mcts_simulations"] = 60
torch.manual_seed(0)
model = RecurrentNet(2, 1, 64, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=False)
initialize_parameters(model)
res = {}
REPS = int(os.environ.get("REPS", "1"))
runs = [("plain", None, 2)] + [("nopub%d" % i, 0, 12) for i in range(REPS)] + [("pub%d" % i, 64, 12) for i in range(REPS)]
for tag, pw, budget in runs:
    e = SearchEngine(tic_tac_toe_spec(), cfg, 96, True, policy_is_prob=False, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.BF16,
                     auto_advance=True, games_per_slot=2, pool_nodes=4000, seed=9, max_sims_per_launch=budget, record_detail=True)
    if pw is None:
        net = FusedRecurrentForward(e, model, 2, use_graph=True)
    else:
        net = CachedForward(e, lambda v: FusedRecurrentForward(v, model, 2, use_graph=True), capacity_log2=16, min_rows=32, in_kernel=True, publish_width=pw)
    run_until_idle(e, net)
    recs, _ = e.drain_records()
    res[tag] = group_games(recs)
    if tag.startswith("pub"):
        meta = net.meta.cpu().numpy(); em = net.exp_meta.cpu().numpy()
        keys = net.keys.cpu().numpy().view(np.uint32); pol = net.pol.float().cpu().numpy()
        acts = net.exp_act.cpu().numpy().view(np.uint16); pri = net.exp_prior.cpu().numpy()
        bad = 0; n = 0
        for p in np.nonzero(em > 0)[0]:
            K = em[p] - 1
            w = int(keys[p, 0]); occ = (w | (w >> 9)) & 0x1ff
            legal = [a for a in range(9) if not (occ >> a) & 1]
            x = pol[p].astype(np.float32)
            ex = np.exp(x - x.max()); pr = ex / ex.sum(dtype=np.float32)
            tot = np.float64(0)
            for a in legal: tot += np.float64(pr[a])
            want = [np.float64(pr[a]) / tot for a in legal]
            n += 1
            if K != len(legal) or list(acts[p, :K]) != legal or not np.allclose(pri[p, :K], want, rtol=1e-4):
                bad += 1
                if bad < 6: print("slot", p, "meta", meta[p], "K", K, "legal", legal, "acts", acts[p, :K], "pri", pri[p, :K], "want", want)
        print("published", n, "bad", bad, "entries ready", int((meta == 2).sum()), "pending", int((meta == 3).sum()), "busy", int((meta == 1).sum()))
def same(a, b):
    diffs = 0
    for uid in a:
        for x, y in zip(a[uid], b[uid]):
            if x["action"] != y["action"] or not np.array_equal(x["child_N"], y["child_N"]) or x["root_W"] != y["root_W"]:
                diffs += 1
    return diffs
print("plain vs nopub diffs", [same(res["plain"], res["nopub%d" % i]) for i in range(REPS)], " plain vs pub diffs", [same(res["plain"], res["pub%d" % i]) for i in range(REPS)])
