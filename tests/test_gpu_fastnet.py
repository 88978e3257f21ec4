"""GPU: the im2col + GEMM fast path of the network forward against the nn.Module forward (fp32)."""
import numpy as np
import pytest
import torch

import golden_io

pytestmark = pytest.mark.gpu


def test_neighbour_tables_follow_the_board_geometry():
    from nuzero_b200.fastnet import hex_neighbour_table
    from oracle.scs import SCS, Scenario

    for R, C in [(5, 5), (4, 7), (15, 15)]:
        sc = Scenario()
        sc.rows, sc.cols = R, C
        g = SCS.__new__(SCS)
        g.sc = sc
        tab = hex_neighbour_table(R, C).numpy()
        for t in range(R * C):
            assert tab[t, 0] == t and tab[t, 1:].tolist() == g.neighbours(t)


@pytest.mark.parametrize("hexa,shape", [(True, "scs"), (False, "ttt")])
def test_fast_forward_matches_module(hexa, shape):
    import os

    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine, tic_tac_toe_spec
    from nuzero_b200.fastnet import FastRecurrentForward
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.nets import RecurrentNet, initialize_parameters

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    if shape == "scs":
        scn = ScsScenario(os.path.join(golden_io.GOLDEN, "scs_configs", "randomized_config_5.yml"), [3])
        spec, cin, planes = scn.spec(), scn.C, scn.planes
    else:
        spec, cin, planes = tic_tac_toe_spec(), 2, 1
    G = 64
    e = SearchEngine(spec, cfg, G, False, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.F32, pool_nodes=64)
    torch.manual_seed(0)
    model = RecurrentNet(cin, planes, 64, 2, recall=True, policy_head="conv", value_head="reduce",
                         value_activation="relu", hex=hexa).to(e.device)
    initialize_parameters(model)
    e.leaf.copy_((torch.rand_like(e.leaf.float()) < 0.3).float() * torch.randint(1, 4, e.leaf.shape, device=e.device))
    fast = FastRecurrentForward(e, model, iters_to_do=3, use_graph=True)
    fast()
    with torch.no_grad():
        (p, v), _ = model(e.leaf.float(), 3)
    p = p.reshape(G, -1)
    scale = float(p.abs().max())
    assert float((e.policy - p).abs().max()) < 0.03 * scale + 1e-3
    # bf16 activations through 3 recurrent iterations: a few 1e-2 absolute on a tanh output
    assert float((e.value - v.reshape(-1)).abs().max()) < 0.1
    assert float((e.value - v.reshape(-1)).abs().mean()) < 0.03
    # same arg-max policy entry on (almost) every row
    agree = (e.policy.argmax(1) == p.argmax(1)).float().mean().item()
    assert agree > 0.9
