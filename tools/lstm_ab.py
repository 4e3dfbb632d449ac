#!/usr/bin/env python
"""A/B of the LSTM kernel's scheduling variants (wm_debug_lstm_opts): time per launch, per-phase cycle sums of block 0,
and bit-equality of the output with variant 0.  usage: lstm_ab.py [B ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wmb200  # noqa: E402
from wmb200 import _lib as L, ops  # noqa: E402

Bs = [int(a) for a in sys.argv[1:]] or [32, 4096]
OPTS = [int(o) for o in os.environ.get("LSTM_OPTS", "0,1").split(",")]
T = 16000
dev = "cuda"
torch.manual_seed(0)
gen = wmb200.Generator(16).to(dev).eval()
blob = gen.packed()
lib = L.load()
wpk, bpk = blob[L.G_TC_LSTM_W:], blob[L.G_TC_LSTM_B:]
st = torch.cuda.current_stream().cuda_stream
names = ["wait acc", "tmem ld", "phase1", "bar", "phase2", "arrive", "store", "-", "mma wait h", "mma issue h", "mma wait acc", "mma x"]
out = []
for B in Bs:
    x = ops.to_planar(torch.randn(B, T, 64, device=dev))
    ref = None
    for o in OPTS:
        lib.wm_debug_lstm_opts(o)
        y = torch.zeros_like(x)
        ms = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.wm_lstm_tc_fwd(x.data_ptr(), wpk.data_ptr(), bpk.data_ptr(), None, y.data_ptr(), B, T, st), "lstm")
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        same = None
        if ref is None:
            ref = y.clone()
        else:
            same = bool(torch.equal(ref, y))
            if not same:
                nb = min(B, 64)
                same = float((ops.from_planar(ref, B, T)[:nb] - ops.from_planar(y, B, T)[:nb]).abs().max())
        prof = torch.zeros(16, dtype=torch.int64, device=dev)
        lib.wm_debug_lstm_profile(prof.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.wm_lstm_tc_fwd(x.data_ptr(), wpk.data_ptr(), bpk.data_ptr(), None, y.data_ptr(), B, T, st), "lstm")
        e1.record()
        torch.cuda.synchronize()
        ms_prof = e0.elapsed_time(e1)
        lib.wm_debug_lstm_profile(None)
        p = (prof.cpu().double() / T).tolist()
        rec = {"B": B, "opts": o, "ms_min": round(min(ms), 3), "ms": [round(m, 3) for m in ms], "same_as_opts0": same, "ms_prof_build": round(ms_prof, 3), "mhz_prof": round(sum(p[:7]) * T / ms_prof / 1e3),
               "phases": {n: round(v) for n, v in zip(names, p) if n != "-"}}
        print(json.dumps(rec), flush=True)
        out.append(rec)
lib.wm_debug_lstm_opts(-1)
