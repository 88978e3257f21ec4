#!/usr/bin/env python
"""Micro-benchmark of nz_hexconv_bf16 (one 256 -> 256 channel hexagonal convolution over 4096 x 5 x 5 cells, the
layer that makes up 97 % of the SCS-5 network forward) for each kernel variant, CUDA events, L2 flushed between runs.

    python tools/bench_hexconv.py [--boards 4096] [--rows 5] [--cols 5] [--cin 256] [--cout 256] [--flags 0,2,8]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nuzero_b200 import _ffi  # noqa: E402
from nuzero_b200.fastnet import hex_neighbour_table  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--boards", type=int, default=4096)
    ap.add_argument("--rows", type=int, default=5)
    ap.add_argument("--cols", type=int, default=5)
    ap.add_argument("--cin", type=int, default=256)
    ap.add_argument("--cout", type=int, default=256)
    ap.add_argument("--flags", default="0,2,8")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--trace", action="store_true")
    a = ap.parse_args()
    dev = "cuda"
    nbr = hex_neighbour_table(a.rows, a.cols).to(dev)
    taps, RC = nbr.shape[1], a.rows * a.cols
    rows = a.boards * RC
    torch.manual_seed(0)
    x = (torch.randn(rows, a.cin, device=dev) * 0.5).to(torch.bfloat16)
    wt = (torch.randn(a.cout, taps * a.cin, device=dev) / (taps * a.cin) ** 0.5).to(torch.bfloat16)
    res = torch.randn(rows, a.cout, device=dev).to(torch.bfloat16)
    out = torch.empty(rows, a.cout, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    L = _ffi.lib()
    flops = 2.0 * rows * a.cout * taps * a.cin
    ref = None
    for flag in [int(f) for f in a.flags.split(",")]:
        def run():
            _ffi.check(L.nz_hexconv_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(nbr.data_ptr()), C.c_void_p(wt.data_ptr()),
                                         C.c_void_p(res.data_ptr()), C.c_void_p(out.data_ptr()), rows, RC, taps, a.cin, a.cout,
                                         a.cout, flag, 1, None))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        err = float((out.float() - ref.float()).abs().max())
        cold, warm = [], []
        for _ in range(a.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            cold.append(e0.elapsed_time(e1) * 1e3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        back = e0.elapsed_time(e1) * 1e3 / a.reps
        cold.sort()
        print(json.dumps({"flag": flag, "rows": rows, "cin": a.cin, "cout": a.cout, "taps": taps, "us_cold_median": cold[len(cold) // 2],
                          "us_cold_min": cold[0], "us_back_to_back": back, "tflops_back_to_back": flops / back / 1e6,
                          "max_abs_diff_vs_first": err}))
        if a.trace:
            n_chunks = taps * a.cin // 64
            buf = torch.zeros(4 * n_chunks + 4, dtype=torch.int64, device=dev)
            L.nz_hexconv_set_trace(C.c_void_p(buf.data_ptr()))
            run()
            torch.cuda.synchronize()
            L.nz_hexconv_set_trace(None)
            t = buf.cpu().tolist()
            t0 = t[0]
            print("trace flag", flag, "chunk: free issue-start | published | mma-sees | issued   (cycles from first issue)")
            for kc in range(n_chunks):
                print("  %2d %7d %7d %7d %7d" % (kc, t[4 * kc] - t0, t[4 * kc + 1] - t0, t[4 * kc + 2] - t0, t[4 * kc + 3] - t0))
            print("  epilogue %d .. %d" % (t[4 * n_chunks] - t0, t[4 * n_chunks + 1] - t0))


if __name__ == "__main__":
    main()
