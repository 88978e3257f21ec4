"""Inference-only fast paths for RecurrentNet (hex or orthogonal) on a cells-major / channels-last layout
[B*R*C, channels]:

* `FastRecurrentForward`: every convolution is a neighbour-table im2col (own CUDA kernel, `nz_im2col_bf16`)
  followed by ONE library GEMM (torch.matmul -> cuBLAS, bf16, tensor cores);
* `FusedRecurrentForward`: every convolution is ONE launch of the hand-written tcgen05 kernel
  `nz_hexconv_bf16` (csrc/hexgemm.cuh): the gather feeds the MMA pipeline directly, accumulators live in
  tensor memory, residual add and ReLU happen in the epilogue — no im2col matrix, no elementwise kernels.

Compared with running the nn.Module (three small cuDNN convolutions per hexagonal layer) this does
exactly the 7 taps of the hexagonal stencil in a single large GEMM, and uses the linearity of the
recall convolution: conv([thought, x]) = conv_t(thought) + conv_x(x), where conv_x(x) does not change
between recurrent iterations and is computed once.  The weights are read from the module, the module
itself is untouched.  Results equal the module's forward up to bf16 rounding (tests/test_gpu_fastnet.py).
"""
import ctypes as C

import torch
import torch.nn.functional as F
from torch import nn

from ._ffi import check, lib
from .nets import BasicBlock, HexConv2d, RecurrentNet


def hex_neighbour_table(rows, cols):
    """[R*C, 7] source cells for taps (centre, n, ne, se, s, sw, nw); even columns are shifted up
    (Games/SCS/SCS_Game.py:1048-1094, 1199-1243).  -1 = off the board."""
    out = []
    for r in range(rows):
        for c in range(cols):
            even = c % 2 == 0
            cand = [(r, c), (r - 1, c), (r - 1 if even else r, c + 1), (r if even else r + 1, c + 1), (r + 1, c),
                    (r if even else r + 1, c - 1), (r - 1 if even else r, c - 1)]
            out.append([rr * cols + cc if 0 <= rr < rows and 0 <= cc < cols else -1 for rr, cc in cand])
    return torch.tensor(out, dtype=torch.int32)


def ortho_neighbour_table(rows, cols):
    """[R*C, 9] source cells of a zero-padded 3x3 cross-correlation, taps in (dr, dc) row-major order."""
    out = []
    for r in range(rows):
        for c in range(cols):
            out.append([(r + dr) * cols + (c + dc) if 0 <= r + dr < rows and 0 <= c + dc < cols else -1
                        for dr in (-1, 0, 1) for dc in (-1, 0, 1)])
    return torch.tensor(out, dtype=torch.int32)


def _pad8(n):
    return (n + 7) // 8 * 8


def conv_matrix(conv, cin_pad, cout_pad):
    """Weights of one convolution as the [taps * cin_pad, cout_pad] GEMM operand matching the tables above."""
    if isinstance(conv, HexConv2d):
        k0, k1 = conv.kernel0.detach(), conv.kernel1.detach()
        taps = [k0[:, :, 1, 0], k0[:, :, 0, 0], k1[:, :, 0, 1], k1[:, :, 1, 1], k0[:, :, 2, 0], k1[:, :, 1, 0], k1[:, :, 0, 0]]
    else:
        w = conv.weight.detach()
        taps = [w[:, :, i, j] for i in range(3) for j in range(3)]
    cout, cin = taps[0].shape
    m = taps[0].new_zeros(len(taps), cin_pad, cout_pad)
    for t, w in enumerate(taps):
        m[t, :cin, :cout] = w.t()
    return m.reshape(len(taps) * cin_pad, cout_pad)


class FastRecurrentForward:
    """Callable with the GraphedForward contract: reads engine.leaf, writes engine.policy / engine.value."""

    def __init__(self, engine, network, iters_to_do=2, use_graph=True):
        model = network.get_model() if hasattr(network, "get_model") else network
        if not isinstance(model, RecurrentNet) or not model.recall:
            raise ValueError("FastRecurrentForward handles RecurrentNet(recall=True)")
        self.e, self.iters = engine, iters_to_do
        dev, dt = engine.device, torch.bfloat16
        C_in, R, Cc = engine.state_shape
        self.B, self.RC, self.cin = engine.rows, R * Cc, C_in
        first = model.projection[0]
        hexa = isinstance(first, HexConv2d)
        self.nbr = (hex_neighbour_table(R, Cc) if hexa else ortho_neighbour_table(R, Cc)).to(dev)
        self.K = self.nbr.shape[1]
        Fw = model.num_filters
        cin8 = _pad8(C_in)
        self.cin8, self.F = cin8, Fw
        if Fw % 8:
            raise ValueError("num_filters must be a multiple of 8")

        def mat(conv, ci, co):
            return conv_matrix(conv, ci, co).to(dev).to(dt).contiguous()

        self.w_proj = mat(first, cin8, Fw)
        recall = model.recur_module[0]
        full = conv_matrix(recall, Fw + C_in, Fw).reshape(self.K, Fw + C_in, Fw)  # channels: [thought | x] (RecurrentNet.py:90-91)
        self.w_rec_t = full[:, :Fw].reshape(self.K * Fw, Fw).to(dev).to(dt).contiguous()
        wx = full.new_zeros(self.K, cin8, Fw)
        wx[:, :C_in] = full[:, Fw:]
        self.w_rec_x = wx.reshape(self.K * cin8, Fw).to(dev).to(dt).contiguous()
        self.blocks = []
        for blk in list(model.recur_module)[1:]:
            assert isinstance(blk, BasicBlock)
            self.blocks.append((mat(blk.before_shortcut[0], Fw, Fw), mat(blk.before_shortcut[-1], Fw, Fw)))

        def head(layers):
            convs = [m for m in layers if isinstance(m, (HexConv2d, nn.Conv2d))]
            acts = []
            mods = list(layers)
            for i, m in enumerate(mods):
                if isinstance(m, (HexConv2d, nn.Conv2d)):
                    nxt = mods[i + 1] if i + 1 < len(mods) else None
                    acts.append("relu" if isinstance(nxt, nn.ReLU) else ("tanh" if isinstance(nxt, nn.Tanh) and i + 1 < len(mods) - 1 else None))
            out, ci = [], Fw
            for cv, act in zip(convs, acts):
                co = cv.out_channels
                out.append((mat(cv, _pad8(ci), _pad8(co)), co, act))
                ci = co
            return out

        self.policy_layers = head(model.policy_head.layers)
        self.value_layers = head(model.value_head.layers)
        self.P = self.policy_layers[-1][1]
        kmax = self.K * max(Fw, cin8, max(_pad8(l[1]) for l in self.policy_layers + self.value_layers))
        self.col = torch.empty(self.B * self.RC, kmax, dtype=dt, device=dev)  # im2col scratch, reused by every layer
        self.graph = None
        with torch.no_grad():
            if use_graph:
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(3):
                        self._run()
                torch.cuda.current_stream(dev).wait_stream(side)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._run()

    def _im2col(self, x, channels, relu=False):
        """x [B*RC, channels] -> view of the scratch [B*RC, K*channels]"""
        col = self.col.view(-1)[: self.B * self.RC * self.K * channels].view(self.B * self.RC, self.K * channels)
        check(lib().nz_im2col_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(self.nbr.data_ptr()), C.c_void_p(col.data_ptr()),
                                   self.B, self.RC, self.K, channels, int(relu), self.e._stream()))
        return col

    def _head(self, t, layers):
        ci = self.F
        for w, co, act in layers:
            t = self._im2col(t, _pad8(ci)) @ w
            if act == "relu":
                t = torch.relu_(t)
            elif act == "tanh":
                t = torch.tanh_(t)
            ci = co
        return t

    def _run(self):
        e = self.e
        B, RC = self.B, self.RC
        x = e.leaf.to(torch.bfloat16).permute(0, 2, 3, 1).reshape(B * RC, self.cin)
        if self.cin8 != self.cin:
            x = F.pad(x, (0, self.cin8 - self.cin))
        x = x.contiguous()
        gx = self._im2col(x, self.cin8)
        thought = torch.relu_(gx @ self.w_proj)           # projection (RecurrentNet.py:82-83)
        rx = gx @ self.w_rec_x                            # the x-half of the recall convolution, once
        for _ in range(self.iters):
            t = torch.addmm(rx, self._im2col(thought, self.F), self.w_rec_t)
            for w1, w2 in self.blocks:                    # BasicBlock (blocks.py:36-40)
                h = self._im2col(t, self.F) @ w1
                t = torch.relu_(torch.addmm(t, self._im2col(h, self.F, relu=True), w2))
            thought = t
        p = self._head(thought, self.policy_layers)[:, : self.P]           # [B*RC, planes]
        e.policy.copy_(p.reshape(B, RC, self.P).permute(0, 2, 1).reshape(B, e.A))
        v = self._head(thought, self.value_layers)[:, :1]                   # [B*RC, 1]
        e.value.copy_(torch.tanh(v.reshape(B, RC).float().mean(1)))        # AdaptiveAvgPool3d(1) -> Tanh

    def __call__(self):
        with torch.no_grad():
            if self.graph is not None:
                self.graph.replay()
            else:
                self._run()


def _pad64(n):
    return (n + 63) // 64 * 64


def _pad16(n):
    return max(16, (n + 15) // 16 * 16)


class FusedRecurrentForward:
    """Same contract as FastRecurrentForward, convolutions by the fused tcgen05 kernel.  Channel counts are
    padded to multiples of 64 (inputs) / 16 (outputs) with zero weights."""

    def __init__(self, engine, network, iters_to_do=2, use_graph=True):
        model = network.get_model() if hasattr(network, "get_model") else network
        if not isinstance(model, RecurrentNet) or not model.recall:
            raise ValueError("FusedRecurrentForward handles RecurrentNet(recall=True)")
        self.e, self.iters = engine, iters_to_do
        dev, dt = engine.device, torch.bfloat16
        C_in, R, Cc = engine.state_shape
        self.B, self.RC, self.cin = engine.rows, R * Cc, C_in
        self.rows = self.B * self.RC
        first = model.projection[0]
        hexa = isinstance(first, HexConv2d)
        self.nbr = (hex_neighbour_table(R, Cc) if hexa else ortho_neighbour_table(R, Cc)).to(dev)
        self.K = self.nbr.shape[1]
        Fw = model.num_filters
        if Fw % 64:
            raise ValueError("num_filters must be a multiple of 64 for the fused kernel")
        self.F, self.cin64 = Fw, _pad64(C_in)

        def wt(conv, ci_pad, co_pad, rows_from=None):
            m = conv_matrix(conv, ci_pad, co_pad)          # [taps * ci_pad, co_pad]
            return m.t().contiguous().to(dev).to(dt)        # W^T [co_pad, taps * ci_pad], K contiguous

        self.w_proj = wt(first, self.cin64, Fw)
        recall = model.recur_module[0]
        full = conv_matrix(recall, Fw + C_in, Fw).reshape(self.K, Fw + C_in, Fw)
        self.w_rec_t = full[:, :Fw].reshape(self.K * Fw, Fw).t().contiguous().to(dev).to(dt)
        wx = full.new_zeros(self.K, self.cin64, Fw)
        wx[:, :C_in] = full[:, Fw:]
        self.w_rec_x = wx.reshape(self.K * self.cin64, Fw).t().contiguous().to(dev).to(dt)
        self.blocks = [(wt(b.before_shortcut[0], Fw, Fw), wt(b.before_shortcut[-1], Fw, Fw))
                       for b in list(model.recur_module)[1:]]

        def head(layers):
            mods = list(layers)
            out, ci = [], Fw
            for i, m in enumerate(mods):
                if isinstance(m, (HexConv2d, nn.Conv2d)):
                    nxt = mods[i + 1] if i + 1 < len(mods) else None
                    act = "relu" if isinstance(nxt, nn.ReLU) else ("tanh" if isinstance(nxt, nn.Tanh) and i + 1 < len(mods) - 1 else None)
                    co = m.out_channels
                    # a layer whose output feeds another convolution is padded to 64 channels
                    last = not any(isinstance(x, (HexConv2d, nn.Conv2d)) for x in mods[i + 1:])
                    co_pad = _pad16(co) if last else _pad64(co)
                    out.append((wt(m, _pad64(ci), co_pad), co, co_pad, act))
                    ci = co
            return out

        self.policy_layers = head(model.policy_head.layers)
        self.value_layers = head(model.value_head.layers)
        self.P = self.policy_layers[-1][1]
        self.graph = None
        with torch.no_grad():
            if use_graph:
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(3):
                        self._run()
                torch.cuda.current_stream(dev).wait_stream(side)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._run()

    def _conv(self, x, w, residual=None, relu=False):
        n_pad = w.shape[0]
        out = torch.empty(self.rows, n_pad, dtype=torch.bfloat16, device=x.device)
        check(lib().nz_hexconv_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(self.nbr.data_ptr()), C.c_void_p(w.data_ptr()),
                                    None if residual is None else C.c_void_p(residual.data_ptr()), C.c_void_p(out.data_ptr()),
                                    self.rows, self.RC, self.K, x.shape[1], n_pad, n_pad, 0, int(relu), self.e._stream()))
        return out

    def _head(self, t, layers):
        for w, co, co_pad, act in layers:
            t = self._conv(t, w, relu=(act == "relu"))
            if act == "tanh":
                t = torch.tanh_(t)
        return t

    def _run(self):
        e = self.e
        B, RC = self.B, self.RC
        x = e.leaf.to(torch.bfloat16).permute(0, 2, 3, 1).reshape(B * RC, self.cin)
        x = F.pad(x, (0, self.cin64 - self.cin)).contiguous()
        thought = self._conv(x, self.w_proj, relu=True)          # projection + ReLU (RecurrentNet.py:82-83)
        rx = self._conv(x, self.w_rec_x)                          # x-half of the recall convolution, once
        for _ in range(self.iters):
            t = self._conv(thought, self.w_rec_t, residual=rx)    # conv([thought, x])
            for w1, w2 in self.blocks:                            # BasicBlock (blocks.py:36-40)
                h = self._conv(t, w1, relu=True)
                t = self._conv(h, w2, residual=t, relu=True)
            thought = t
        p = self._head(thought, self.policy_layers)[:, : self.P]
        e.policy.copy_(p.reshape(B, RC, self.P).permute(0, 2, 1).reshape(B, e.A))
        v = self._head(thought, self.value_layers)[:, :1]
        e.value.copy_(torch.tanh(v.reshape(B, RC).float().mean(1)))

    def __call__(self):
        with torch.no_grad():
            if self.graph is not None:
                self.graph.replay()
            else:
                self._run()
