#!/usr/bin/env python
"""Time single layers of the main14b_2 tensor-core walk, with the developer switches that remove the stores (1) or the
MMAs (2) of pconv_tc_kernel — where a layer's time goes: python tools/pconv_layer_bench.py [B]"""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wmb200 import _lib as L  # noqa: E402
from wmb200 import pconv as PC  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
be = PC.CudaBackend()
lib = L.load()
torch.manual_seed(0)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def planar(C, T, split=1):
    p = PC.Planar(C, B, T, split, "cuda")
    p.store.zero_()
    return p


cases = {}


def case_ct(cin, cout, s, T):
    ct = nn.ConvTranspose1d(cin, cout, 2 * s, stride=s, padding=s // 2).cuda()
    g = PC.gemm_convT(ct.weight, ct.bias, s, s // 2)
    To = (T - 1) * s - 2 * (s // 2) + 2 * s
    x, y = planar(cin, T), planar(cout, To)
    return lambda: be.run(g, [(x, 0)], B, T, False, None, PC.OUT_CONVT, y, ct=(s, s // 2, cout), out_T=To)


def case_s1(cin, cout, K, T, res, fp32=False):
    c = nn.Conv1d(cin, cout, K, padding=K // 2).cuda()
    g = PC.gemm_conv_s1(c.weight, c.bias)
    x = planar(cin, T)
    if fp32:
        y = torch.empty(B, cout, T, device="cuda")
        return lambda: be.run(g, [(x, 0)], B, T, False, None, PC.OUT_FP32, y, out_T=T, cout=cout)
    y = planar(g.n_total, T)
    return lambda: be.run(g, [(x, 0)], B, T, True, x if res else None, PC.OUT_PLANAR, y)


def case_strided(cin, cout, s, T):
    c = nn.Conv1d(cin, cout, 3, stride=s, padding=1).cuda()
    g = PC.gemm_conv_strided(c.weight, c.bias, s)
    x, y = planar(cin, T // s, s), planar(cout, T // s)
    return lambda: be.run(g, [(x, s - 1), (x, 0), (x, 1)], B, T // s, True, None, PC.OUT_PLANAR, y)


cases["CT3 128->64 s4 T2001"] = case_ct(128, 64, 4, 2001)
cases["CT4 64->32 s2 T8004"] = case_ct(64, 32, 2, 8004)
cases["CT2 256->128 s5 T400"] = case_ct(256, 128, 5, 400)
cases["final 32->17 k7 T16008"] = case_s1(32, 17, 7, 16008, False, fp32=True)
cases["RB128 conv2 T2001 (+res)"] = case_s1(128, 128, 3, 2001, True)
cases["E2c1 64->128 s4 T8000"] = case_strided(64, 128, 4, 8000)
cases["E3c1 128->256 s5 T2000"] = case_strided(128, 256, 5, 2000)
out = {}
for name, fn in cases.items():
    r = {}
    for label, opt in (("full", 0), ("no stores", 1 << 8), ("no mma", 2 << 8), ("neither", 3 << 8)):
        lib.wm_debug_lstm_opts(opt)
        r[label] = round(timed(fn), 3)
    lib.wm_debug_lstm_opts(0)
    out[name] = r
print(json.dumps({"B": B, "ms": out}, indent=1))
