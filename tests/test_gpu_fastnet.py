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
