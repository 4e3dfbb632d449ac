#!/usr/bin/env python
"""Per-stage device times of the embed+detect path (CUDA events, warm), for DESIGN.md / profiles.
Usage: python tools/stage_times.py [B]   (GPU box only)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wmb200
from wmb200 import _lib as L
from wmb200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
T = 16000
dev = "cuda"
torch.manual_seed(0)
gen = wmb200.Generator(16).to(dev).eval()
blob = gen.packed()
lib = L.load()
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


x32 = torch.randn(B, T, 64, device=dev)
y32 = torch.empty_like(x32)
s = torch.randn(B, T, device=dev) * 0.1
xp = ops.to_planar(x32)
yp = torch.empty_like(xp)
rp = ops.to_planar(y32.normal_())
w3 = blob[L.G_RB0 + L.RB_W1:L.G_RB0 + L.RB_W1 + 3 * 4096]
b3 = blob[L.G_RB0 + L.RB_B1:L.G_RB0 + L.RB_B1 + 64]
img3 = blob[L.G_TC:L.G_TC + L.TC_IMG3]
img7 = blob[L.G_TC_CT:L.G_TC_CT + L.TC_IMG7]
out = {"B": B}
out["conv64_fp32_k3_ms"] = timeit(lambda: lib.wm_conv64_fwd(x32.data_ptr(), w3.data_ptr(), b3.data_ptr(), None, None, y32.data_ptr(), B, T, 3, 1, st))
out["conv64_tc_k3_ms"] = timeit(lambda: lib.wm_conv64_tc_fwd(xp.data_ptr(), img3.data_ptr(), b3.data_ptr(), None, yp.data_ptr(), None, B, T, 3, 1, st))
out["conv64_tc_k3_res_ms"] = timeit(lambda: lib.wm_conv64_tc_fwd(xp.data_ptr(), img3.data_ptr(), b3.data_ptr(), rp.data_ptr(), yp.data_ptr(), None, B, T, 3, 1, st))
out["conv64_tc_k3_res_fp32out_ms"] = timeit(lambda: lib.wm_conv64_tc_fwd(xp.data_ptr(), img3.data_ptr(), b3.data_ptr(), rp.data_ptr(), None, y32.data_ptr(), B, T, 3, 1, st))
out["conv64_tc_k7_ms"] = timeit(lambda: lib.wm_conv64_tc_fwd(xp.data_ptr(), img7.data_ptr(), b3.data_ptr(), None, yp.data_ptr(), None, B, T, 7, 0, st))
img33 = blob[L.G_TC:L.G_TC + 2 * L.TC_IMG3]
b1 = blob[L.G_RB0 + L.RB_B1:]; b2 = blob[L.G_RB0 + L.RB_B2:]
out["resblock_tc_fused_ms"] = timeit(lambda: lib.wm_resblock_tc_fwd(xp.data_ptr(), img33.data_ptr(), b1.data_ptr(), b2.data_ptr(), yp.data_ptr(), None, B, T, st))
out["resblock_tc_fused_fp32out_ms"] = timeit(lambda: lib.wm_resblock_tc_fwd(xp.data_ptr(), img33.data_ptr(), b1.data_ptr(), b2.data_ptr(), None, y32.data_ptr(), B, T, st))
out["resblock_tc_cycles_per_tile_at_1965MHz"] = out["resblock_tc_fused_ms"] * 1e-3 * 1.965e9 * 148 / (B * 127)
out["resblock_tc_TFLOPs_algorithmic"] = 2 * 2 * 64 * 64 * 3 * T * B / (out["resblock_tc_fused_ms"] * 1e-3) / 1e12
out["to_planar_ms"] = timeit(lambda: lib.wm_to_planar(x32.data_ptr(), None, yp.data_ptr(), B, T, st))
wih = blob[L.G_LSTM_WIH:L.G_LSTM_WIH + 16384]
whh = blob[L.G_LSTM_WHH:L.G_LSTM_WHH + 16384]
bl = blob[L.G_LSTM_B:L.G_LSTM_B + 256]
out["lstm_fp32_ms"] = timeit(lambda: lib.wm_lstm_fwd(x32.data_ptr(), wih.data_ptr(), whh.data_ptr(), bl.data_ptr(), y32.data_ptr(), B, T, st), reps=2, warm=1)
wpk = blob[L.G_TC_LSTM_W:L.G_TC_LSTM_W + 4 * 256 * 64 // 2]
bpk = blob[L.G_TC_LSTM_B:L.G_TC_LSTM_B + 256]
out["lstm_tc_ms"] = timeit(lambda: lib.wm_lstm_tc_fwd(xp.data_ptr(), wpk.data_ptr(), bpk.data_ptr(), None, yp.data_ptr(), B, T, st), reps=2, warm=1)
out["lstm_tc_cycles_per_step_at_1965MHz"] = out["lstm_tc_ms"] * 1e-3 * 1.965e9 / T / max(1, -(-B // (32 * 148)))
hw = blob[L.G_HEAD_W:L.G_HEAD_W + 64]
hb = blob[L.G_HEAD_B:L.G_HEAD_B + 1]
d = torch.empty(B, T, device=dev)
out["head1_ms"] = timeit(lambda: lib.wm_head_fwd(x32.data_ptr(), hw.data_ptr(), hb.data_ptr(), d.data_ptr(), B, T, 1, st))
fir = wmb200.functional.fir_taps_on(torch.device(dev))
sw = torch.empty_like(s)
out["postprocess_ms"] = timeit(lambda: lib.wm_postprocess_fwd(d.data_ptr(), s.data_ptr(), fir.data_ptr(), None, sw.data_ptr(), None, B, T, 7, 0.02, 0.005, 1e-8, st))
det = wmb200.Detector(16).to(dev).eval()
dblob = det.packed()
pr = torch.empty(B, T, device=dev); cp = torch.empty(B, device=dev); ml = torch.empty(B, 16, device=dev)
out["head_detect17_ms"] = timeit(lambda: lib.wm_detect_heads_fwd(x32.data_ptr(), None, pr.data_ptr(), cp.data_ptr(), ml.data_ptr(), None, B, T // 17 * 17 // 17, 17, st)) if False else None
hwd = dblob[L.D_HEAD_W:L.D_HEAD_W + 17 * 64]
yp32 = ops.conv_in_k7(s, blob[L.G_IN_W:L.G_IN_W + 448], blob[L.G_IN_B:L.G_IN_B + 64])
out["conv_in_fp32_ms"] = timeit(lambda: lib.wm_conv_in_k7_fwd(s.data_ptr(), blob[L.G_IN_W:].data_ptr(), blob[L.G_IN_B:].data_ptr(), y32.data_ptr(), B, T, st))
r = lambda: det.detect(s.unsqueeze(1), want_votes=False)
out["detector_total_ms"] = timeit(r, reps=3, warm=1)
g = lambda: gen(s.unsqueeze(1), torch.zeros(B, dtype=torch.int64, device=dev))
out["generator_total_ms"] = timeit(g, reps=2, warm=1)
tiles = B * 125
out["conv64_tc_k3_cycles_per_tile_at_1965MHz"] = out["conv64_tc_k3_ms"] * 1e-3 * 1.965e9 * 148 / tiles
out["conv64_tc_k3_TFLOPs_algorithmic"] = 2 * 64 * 64 * 3 * T * B / (out["conv64_tc_k3_ms"] * 1e-3) / 1e12
out["conv64_tc_k3_GBps"] = 2 * wmb200._lib.load().wm_planar_bytes(B, T) / (out["conv64_tc_k3_ms"] * 1e-3) / 1e9
print(json.dumps(out, indent=1))
