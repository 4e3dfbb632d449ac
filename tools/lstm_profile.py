#!/usr/bin/env python
"""Per-phase cycle breakdown of one step of the tensor-core LSTM (block 0, averaged over T steps)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wmb200
from wmb200 import _lib as L, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
dev = "cuda"
torch.manual_seed(0)
gen = wmb200.Generator(16).to(dev).eval()
blob = gen.packed()
lib = L.load()
x = ops.to_planar(torch.randn(B, T, 64, device=dev))
y = torch.empty_like(x)
prof = torch.zeros(16, dtype=torch.int64, device=dev)
wpk = blob[L.G_TC_LSTM_W:]; bpk = blob[L.G_TC_LSTM_B:]
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):   # production kernel (no counters)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    L.check(lib.wm_lstm_tc_fwd(x.data_ptr(), wpk.data_ptr(), bpk.data_ptr(), None, y.data_ptr(), B, T, st), "lstm")
    p1.record(); torch.cuda.synchronize()
plain_ms = p0.elapsed_time(p1)
lib.wm_debug_lstm_profile(prof.data_ptr())
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.wm_lstm_tc_fwd(x.data_ptr(), wpk.data_ptr(), bpk.data_ptr(), None, y.data_ptr(), B, T, st), "lstm")
    e1.record(); torch.cuda.synchronize()
lib.wm_debug_lstm_profile(None)
p = (prof.cpu().double() / T).tolist()
names = ["epi wait acc_full", "epi tmem ld + arrive", "epi phase1 exp", "epi named barrier", "epi phase2 + h store",
         "epi fence+arrive h_ready", "epi global store", "-", "mma wait h_ready", "mma issue h part + commit",
         "mma x part (waits + issue)"]
print(json.dumps({"B": B, "T": T, "ms_production_kernel": plain_ms, "ms": e0.elapsed_time(e1), "cycles_per_step_total": e0.elapsed_time(e1) * 1e-3 * 1.965e9 / T,
                  "phases_cycles": {n: round(v, 1) for n, v in zip(names, p) if n != "-"}}, indent=1))
