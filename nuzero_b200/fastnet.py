"""Inference-only fast path for RecurrentNet (hex or orthogonal): every convolution is a neighbour-table
im2col (own CUDA kernel, `nz_im2col_bf16`) followed by ONE library GEMM (torch.matmul -> cuBLAS, bf16,
tensor cores), on a cells-major / channels-last layout [B*R*C, channels].

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
        self.B, self.RC, self.cin = engine.G, R * Cc, C_in
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
