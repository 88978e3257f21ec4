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


def test_fused_tcgen05_conv_matches_reference_gemm():
    """nz_hexconv_bf16 (tcgen05.mma, TMEM accumulators, gather in the loader) against gather + fp32 matmul."""
    import ctypes as C

    from nuzero_b200 import _ffi
    from nuzero_b200.fastnet import hex_neighbour_table, ortho_neighbour_table

    torch.manual_seed(0)
    dev = "cuda"
    for (B, R, Cc, cin, cout, n_pad, relu, res, table) in [
            (11, 5, 5, 64, 16, 16, 0, False, hex_neighbour_table), (64, 5, 5, 128, 256, 256, 0, False, hex_neighbour_table),
            (300, 5, 5, 256, 256, 256, 1, True, hex_neighbour_table), (7, 15, 15, 64, 21, 32, 0, False, hex_neighbour_table),
            (3, 30, 30, 128, 138, 192, 1, False, hex_neighbour_table), (500, 3, 3, 64, 64, 64, 1, True, ortho_neighbour_table)]:
        RC, rows = R * Cc, B * R * Cc
        nbr = table(R, Cc).to(dev)
        taps = nbr.shape[1]
        x = (torch.randn(rows, cin, device=dev) * 0.5).to(torch.bfloat16)
        w = (torch.randn(taps * cin, cout, device=dev) / (taps * cin) ** 0.5).to(torch.bfloat16)
        wt = torch.zeros(n_pad, taps * cin, device=dev, dtype=torch.bfloat16)
        wt[:cout] = w.t()
        resid = torch.randn(rows, n_pad, device=dev).to(torch.bfloat16) if res else None
        out = torch.full((rows, n_pad), 7.0, device=dev, dtype=torch.bfloat16)
        _ffi.check(_ffi.lib().nz_hexconv_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(nbr.data_ptr()), C.c_void_p(wt.data_ptr()),
                                              None if resid is None else C.c_void_p(resid.data_ptr()), C.c_void_p(out.data_ptr()),
                                              rows, RC, taps, cin, n_pad, n_pad, 0, relu, None))
        xp = torch.cat([x.float().view(B, RC, cin), torch.zeros(B, 1, cin, device=dev)], 1)
        idx = nbr.long().clone()
        idx[idx < 0] = RC
        ref = xp[:, idx.view(-1)].view(rows, taps * cin) @ w.float()
        if res:
            ref = ref + resid.float()[:, :cout]
        if relu:
            ref = torch.relu(ref)
        assert float((out.float()[:, :cout] - ref).abs().max()) < 0.02 * max(1.0, float(ref.abs().max()))
        if n_pad > cout:
            assert float(out.float()[:, cout:].abs().max()) == 0.0
        # variants (include/nz_engine.h): bit 1 = taps innermost in K (L1-allocating gathers), bit 2 = one CTA per tile
        # instead of the two-CTA tcgen05 pair.  The pairing does not change the summation order, the K order does.
        outs = {}
        for flag in (2, 4, 6):
            o = torch.full((rows, n_pad), 7.0, device=dev, dtype=torch.bfloat16)
            _ffi.check(_ffi.lib().nz_hexconv_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(nbr.data_ptr()), C.c_void_p(wt.data_ptr()),
                                                  None if resid is None else C.c_void_p(resid.data_ptr()), C.c_void_p(o.data_ptr()),
                                                  rows, RC, taps, cin, n_pad, n_pad, flag, relu, None))
            outs[flag] = o
        assert torch.equal(outs[4], out)
        assert torch.equal(outs[2], outs[6])
        assert float((outs[2].float()[:, :cout] - ref).abs().max()) < 0.02 * max(1.0, float(ref.abs().max()))
        assert float((outs[2].float() - out.float()).abs().max()) < 0.01 * max(1.0, float(ref.abs().max()))
    with pytest.raises(_ffi.NzError):
        _ffi.check(_ffi.lib().nz_hexconv_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(nbr.data_ptr()), C.c_void_p(wt.data_ptr()),
                                              None, C.c_void_p(out.data_ptr()), rows, RC, taps, 60, n_pad, n_pad, 0, 0, None))


def test_fused_conv_shape_sweep_pair_vs_single_cta():
    """Random shapes (ragged row counts, every n_pad the network's heads use, 1..9 taps): the two-CTA pair kernel, the
    single-CTA kernel and the fp32 reference agree; rows beyond the tensor are never written."""
    import ctypes as C

    from nuzero_b200 import _ffi
    from nuzero_b200.fastnet import hex_neighbour_table, ortho_neighbour_table

    g = torch.Generator().manual_seed(1)
    dev = "cuda"
    L = _ffi.lib()
    for trial in range(14):
        R, Cc = [(5, 5), (3, 3), (7, 4), (15, 15), (10, 10), (30, 30), (1, 1)][trial % 7]
        table = ortho_neighbour_table if trial % 5 == 4 else hex_neighbour_table
        B = int(torch.randint(1, 40, (1,), generator=g)) if R * Cc < 400 else 1
        cin = [64, 128, 192, 256, 320][int(torch.randint(0, 5, (1,), generator=g))]
        n_pad = [16, 32, 48, 64, 96, 128, 160, 192, 224, 256][int(torch.randint(0, 10, (1,), generator=g))]
        relu, res = trial % 2, trial % 3 == 0
        RC, rows = R * Cc, B * R * Cc
        nbr = table(R, Cc).to(dev)
        taps = nbr.shape[1]
        x = (torch.randn(rows, cin, generator=g) * 0.5).to(dev).to(torch.bfloat16)
        wt = (torch.randn(n_pad, taps * cin, generator=g) / (taps * cin) ** 0.5).to(dev).to(torch.bfloat16)
        resid = torch.randn(rows, n_pad, generator=g).to(dev).to(torch.bfloat16) if res else None
        outs = []
        for flag in (0, 4, 16, 48):  # pair / single CTA / pair without the small-batch forms: TMA row gather, cp.async gathers
            guard = torch.full((rows + 64, n_pad), 7.0, device=dev, dtype=torch.bfloat16)
            _ffi.check(L.nz_hexconv_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(nbr.data_ptr()), C.c_void_p(wt.data_ptr()),
                                         None if resid is None else C.c_void_p(resid.data_ptr()), C.c_void_p(guard.data_ptr()),
                                         rows, RC, taps, cin, n_pad, n_pad, flag, relu, None))
            assert float((guard[rows:].float() - 7.0).abs().max()) == 0.0, "rows beyond the tensor were written"
            outs.append(guard[:rows])
        assert all(torch.equal(outs[0], o) for o in outs[1:]), (trial, R, Cc, B, cin, n_pad)
        xp = torch.cat([x.float().view(B, RC, cin), torch.zeros(B, 1, cin, device=dev)], 1)
        idx = nbr.long().clone()
        idx[idx < 0] = RC
        ref = xp[:, idx.view(-1)].view(rows, taps * cin) @ wt.float().t()
        if res:
            ref = ref + resid.float()
        if relu:
            ref = torch.relu(ref)
        assert float((outs[0].float() - ref).abs().max()) < 0.02 * max(1.0, float(ref.abs().max())), (trial, R, Cc, B, cin, n_pad)


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
        scn = ScsScenario(os.path.join(golden_io.SCS_CONFIGS, "randomized_config_5.yml"), [3])
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
    from nuzero_b200.fastnet import FusedRecurrentForward
    fused = FusedRecurrentForward(e, model, iters_to_do=3, use_graph=True)
    fused()
    p_fused, v_fused = e.policy.clone(), e.value.clone()
    e.policy.zero_()
    e.value.zero_()
    fast = FastRecurrentForward(e, model, iters_to_do=3, use_graph=True)
    fast()
    # the two fast paths agree with each other as well as bf16 allows
    assert float((p_fused - e.policy).abs().max()) < 0.03 * float(e.policy.abs().max()) + 1e-3
    assert float((v_fused - e.value).abs().max()) < 0.05
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


@pytest.mark.parametrize("cfg_name,iters,G", [("mirrored_config_5.yml", 6, 256), ("solo_soldier_config_15.yml", 20, 32)])
def test_fused_forward_at_baseline_shapes_against_fp32_module(cfg_name, iters, G):
    """the bf16 tcgen05 forward at the shapes BASELINE.json names — RecurrentNet(C, planes, 256 filters,
    2 blocks, recall, hex) x 6 iterations on the 5x5 board (configs[2]) and x 20 on 15x15 (configs[3]) — against the fp32
    nn.Module on REAL leaf positions (the search's own leaf tensor after a few hundred launches), where the bf16 rounding of
    up to 20 recurrent passes accumulates.  hexagdly itself is absent: the module is this repo's restatement."""
    import os

    from nuzero_b200 import _ffi
    from nuzero_b200.engine import SearchEngine
    from nuzero_b200.fastnet import FusedRecurrentForward
    from nuzero_b200.games.scs_config import ScsScenario
    from nuzero_b200.nets import RecurrentNet, initialize_parameters

    cfg = golden_io.load("ttt_p0_s25_salt0")["cfg"]
    cfg["Simulation"]["mcts_simulations"] = 24
    scn = ScsScenario(os.path.join(golden_io.SCS_CONFIGS, cfg_name), [None if "mirrored" in cfg_name else 1])
    e = SearchEngine(scn.spec(), cfg, G, True, leaf_dtype=_ffi.BF16, policy_dtype=_ffi.F32, pool_nodes=20000, max_depth=128,
                     policy_is_prob=False, seed=3)
    e.set_maps([0] * G)
    e.reset()
    torch.manual_seed(0)
    model = RecurrentNet(scn.C, scn.planes, 256, 2, recall=True, policy_head="conv", value_head="reduce",
                         value_activation="relu", hex=True).to(e.device)
    initialize_parameters(model)
    # Xavier-initialised weights are not a trained network: over 6-20 recurrent passes the activations grow to 1e2 .. 1e8
    # and every soft-max saturates.  One common factor on all parameters (found by bisection on the fp32 module) brings the
    # logits to the scale of a trained policy head (|logit| ~ 10), where a comparison of probabilities means something.
    probe = (torch.rand(8, scn.C, scn.rows, scn.cols, device=e.device) < 0.2).float()
    base = [q.detach().clone() for q in model.parameters()]
    lo, hi = 0.05, 1.0
    for _ in range(16):
        c = (lo * hi) ** 0.5
        with torch.no_grad():
            for q, b in zip(model.parameters(), base):
                q.copy_(b * c)
            (pp, _), _ = model(probe, iters)
        if float(pp.abs().max()) > 10.0:
            hi = c
        else:
            lo = c
    fused = FusedRecurrentForward(e, model, iters_to_do=iters, use_graph=True)
    for _ in range(150):  # games spread over openings and middle games: the leaf rows are real positions
        e.advance()
        fused()
    e.raise_on_error()
    waiting = (e.phases() == _ffi.PHASE_LEAF_PENDING)
    assert int(waiting.sum()) > G // 2
    fused()
    with torch.no_grad():
        (p, v), _ = model(e.leaf.float(), iters)
    p, v = p.reshape(G, -1)[waiting], v.reshape(-1)[waiting]
    pf, vf = e.policy[waiting], e.value[waiting]
    scale = float(p.abs().max())
    perr, verr = (pf - p).abs(), (vf - v).abs()
    soft_err = (torch.softmax(pf, 1) - torch.softmax(p, 1)).abs().max(1).values
    agree = (pf.argmax(1) == p.argmax(1)).float().mean().item()
    report = dict(rows=int(waiting.sum()), logit_scale=scale, logit_max_err=float(perr.max()), logit_mean_err=float(perr.mean()),
                  prob_max_err=float(soft_err.max()), prob_mean_err=float(soft_err.mean()), value_max_err=float(verr.max()),
                  value_mean_err=float(verr.mean()), argmax_agreement=agree)
    print("fused forward vs fp32 module, %s x%d: %s" % (cfg_name, iters, report))
    assert 1.0 < scale < 100.0
    assert report["logit_max_err"] < 0.05 * scale + 5e-3
    assert report["logit_mean_err"] < 0.01 * scale + 1e-3
    assert report["prob_max_err"] < 0.05 and report["prob_mean_err"] < 0.01
    assert report["value_max_err"] < 0.1 and report["value_mean_err"] < 0.02
    assert agree > 0.9
