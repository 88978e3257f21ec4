"""GPU: the WHOLE search tree (Search/Node.py:3-32 — every node's visit count, value sum, prior, child set) against the
tree dumps the real reference wrote into the golden fixtures (`tree%d_i / tree%d_f`, oracle/gen_golden.py: dump_ref_tree):
north_star "tree shape ... bit-exact".  The engine runs in manual mode (the kernel stops at the end of every move's
search), the device tree is walked depth first in ascending action order — the reference's dict order — and compared row
for row: (depth, action, N, number of children) and (W, prior)."""
import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu


def dump_device_tree(e, g=0):
    """Pre-order rows like oracle.mcts.dump_tree, read from the node pool of slot g (root = node 0, its children from 1)."""
    N = e.node_N[g].cpu().numpy().astype(np.int64)
    W = e.node_W[g].cpu().numpy()
    prior = e.node_prior[g].cpu().numpy()
    act = (e.node_flags[g].cpu().numpy().astype(np.int64) >> 16) & 0xFFFF
    base = e.node_base[g].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    K = e.node_K[g].cpu().numpy().astype(np.int64)
    ints, flts = [], []
    stack = [(0, 0, -1)]
    while stack:
        n, depth, action = stack.pop()
        ints.append((depth, action, N[n], K[n]))
        flts.append((W[n], prior[n]))
        for c in range(int(K[n]) - 1, -1, -1):
            stack.append((int(base[n]) + c, depth + 1, int(act[base[n] + c])))
    return np.array(ints, dtype=np.int64), np.array(flts, dtype=np.float64)


def _spec_for(g):
    from nuzero_b200.engine import tic_tac_toe_spec
    from nuzero_b200.games.scs_config import ScsScenario
    import os

    game = str(g["game"])
    if game == "ttt":
        return tic_tac_toe_spec(), None, {}
    _, cfg_name, seed = game.split(":") if game.count(":") == 2 else (None, game[4:], int(g["seed"]))
    seed = None if str(seed) in ("None", "0") else int(seed)
    scn = ScsScenario(os.path.join(golden_io.SCS_CONFIGS, cfg_name), [seed])
    return scn.spec(), [0], dict(max_depth=256)


@pytest.mark.parametrize("name", golden_io.names("ttt_") + golden_io.names("scs_p"))
def test_whole_tree_matches_reference_dump(name):
    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine
    from nuzero_b200.stubnet import DyadicStubNet

    g = golden_io.load(name)
    tree_moves = set(g["tree_moves"].tolist())
    assert tree_moves, "fixture without tree dumps"
    spec, maps, kw = _spec_for(g)
    tm, tw = (0, 0)
    if g["training"]:
        tm, tw = g["gamma_tape"].shape[0] + 1, max(8, g["gamma_tape"].shape[1])
    e = SearchEngine(spec, g["cfg"], 1, g["training"], policy_is_prob=True, leaf_dtype=_ffi.F32, policy_dtype=_ffi.F32,
                     auto_advance=False, max_sims_per_launch=64, tape_moves=tm, tape_width=tw, pool_nodes=400000, **kw)
    if g["training"]:
        gm, un = np.zeros((1, tm, tw)), np.zeros((1, tm, 3))
        gm[0, : g["gamma_tape"].shape[0], : g["gamma_tape"].shape[1]] = g["gamma_tape"]
        un[0, : g["unif_tape"].shape[0]] = g["unif_tape"]
        e.set_tapes(gm, un)
    if maps is not None:
        e.set_maps(maps)
        e.reset()
    net = DyadicStubNet(e, salt=[g["salt"]])
    checked, nodes = 0, 0
    actions = g["actions"].tolist()
    for move in range(len(actions)):
        for _ in range(100000):
            e.advance()
            net()
            ph = int(e.ctl[0, _ffi.CTL_PHASE])
            if ph != _ffi.PHASE_READY and ph != _ffi.PHASE_LEAF_PENDING:
                break
        e.raise_on_error()
        assert ph == _ffi.PHASE_MOVE_READY, (move, ph)
        if move in tree_moves:
            ti, tf = dump_device_tree(e)
            np.testing.assert_array_equal(ti, g["tree%d_i" % move], err_msg="tree ints, move %d" % move)
            np.testing.assert_array_equal(tf, g["tree%d_f" % move], err_msg="tree floats, move %d" % move)
            checked += 1
            nodes += len(ti)
        chosen = int(e.ctl[0, _ffi.CTL_CHOSEN])
        assert int(e.node_action(0, 1 + chosen)) == actions[move], "move %d" % move
        e.commit_moves()
        e.raise_on_error()
    assert int(e.ctl[0, _ffi.CTL_PHASE]) == _ffi.PHASE_IDLE  # the game is over where the reference's ended
    assert checked == len(tree_moves) and nodes > 0
