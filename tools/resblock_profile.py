#!/usr/bin/env python
"""Per-phase cycle breakdown of the fused ResBlock kernel (block 0, averaged over its tiles)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wmb200
from wmb200 import _lib as L, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
T = 16000
dev = "cuda"
torch.manual_seed(0)
gen = wmb200.Generator(16).to(dev).eval()
blob = gen.packed()
lib = L.load()
x = ops.to_planar(torch.randn(B, T, 64, device=dev))
y = torch.empty_like(x)
prof = torch.zeros(32, dtype=torch.int64, device=dev)
lib.wm_debug_lstm_profile(prof.data_ptr())
img = blob[L.G_TC:]; b1 = blob[L.G_RB0 + L.RB_B1:]; b2 = blob[L.G_RB0 + L.RB_B2:]
st = torch.cuda.current_stream().cuda_stream
HB = os.environ.get("HB", "1") == "1"          # biases by value (the product variant) or from device memory
hb1, hb2 = b1[:64].cpu().contiguous(), b2[:64].cpu().contiguous()
def launch():
    if HB:
        L.check(lib.wm_resblock_tc_hostbias_fwd(x.data_ptr(), img.data_ptr(), hb1.data_ptr(), hb2.data_ptr(), y.data_ptr(), None, B, T, st), "rb")
    else:
        L.check(lib.wm_resblock_tc_fwd(x.data_ptr(), img.data_ptr(), b1.data_ptr(), b2.data_ptr(), y.data_ptr(), None, B, T, st), "rb")
lib.wm_debug_lstm_profile(None)
ms_plain = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch(); e1.record(); torch.cuda.synchronize()
    ms_plain.append(e0.elapsed_time(e1))
lib.wm_debug_lstm_profile(prof.data_ptr())
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record(); torch.cuda.synchronize()
lib.wm_debug_lstm_profile(None)
p = prof.cpu().double()
n = max(p[31].item(), 1)
names = {16: "mma wait x full(i+1)", 17: "mma issue conv1(i+1)", 18: "mma wait u_full(i) + d2_empty",
         19: "mma issue conv2(i)", 24: "g1 wait d1_full", 25: "g1 wait u_empty", 26: "g1 tmem->U + fence + arrive",
         27: "g2 prefetch + wait d2_full", 28: "g2 tmem + residual -> y"}
print(json.dumps({"B": B, "host_bias": HB, "ms_production_kernel": min(ms_plain), "ms": e0.elapsed_time(e1), "tiles_block0": n,
                  "cycles_per_tile": e0.elapsed_time(e1) * 1e-3 * 1.965e9 / n,
                  "phases_cycles_per_tile": {v: round(p[k].item() / n, 1) for k, v in names.items()}}, indent=1))
