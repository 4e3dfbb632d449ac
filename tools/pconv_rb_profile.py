#!/usr/bin/env python
"""Per-role cycle breakdown of the fused main14b_2 ResidualBlock kernel (pconv_rb_kernel, block 0, per tile):
python tools/pconv_rb_profile.py C B T"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wmb200 import _lib as L  # noqa: E402
from wmb200 import main14b_2 as M  # noqa: E402
from wmb200 import pconv as PC  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
T = int(sys.argv[3]) if len(sys.argv) > 3 else 8004
torch.manual_seed(0)
blk = M.ResidualBlock(C, C).cuda()
x = torch.randn(B, C, T, device="cuda")
be = PC.CudaBackend()
xin = PC.Planar(C, B, T, 1, "cuda")
be.to_planar(x, xin)
out = PC.Planar(C, B, T, 1, "cuda")
g1, g2 = PC.gemm_conv_s1(blk.conv1.weight, blk.conv1.bias), PC.gemm_conv_s1(blk.conv2.weight, blk.conv2.bias)
lib = L.load()
prof = torch.zeros(32, dtype=torch.int64, device="cuda")


def run():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    be.run(g1, [(xin, 0)], B, T, True, xin, PC.OUT_PLANAR, out, g2=g2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


ms = [run() for _ in range(4)]
lib.wm_debug_lstm_profile(prof.data_ptr())
msp = [run() for _ in range(2)]
lib.wm_debug_lstm_profile(None)
p = prof.cpu().double()
n = max(p[31].item(), 1)
names = {0: "producer wait empty", 1: "producer issue", 8: "mma wait t1_empty", 9: "mma wait stage full", 10: "mma issue G1",
         11: "mma wait u_full", 12: "mma wait t2_empty", 13: "mma issue G2 (+skip)", 15: "mma loop overhead",
         16: "e1 wait t1_full", 17: "e1 tmem ld + bias", 18: "e1 wait u_empty", 19: "e1 elu/split/st.shared",
         20: "e1 fence + arrive", 21: "e1 loop head", 24: "e2 wait t2_full", 25: "e2 tmem ld", 26: "e2 bias/residual/elu/store",
         27: "e2 loop head + prefetch"}
print(json.dumps({"C": C, "B": B, "T": T, "ms": min(ms), "ms_profiled": min(msp), "tiles_block0": n,
                  "cycles_per_tile": {v: round(p[k].item() / n, 1) for k, v in names.items()}}, indent=1))
