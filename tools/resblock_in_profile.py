#!/usr/bin/env python
"""Per-phase cycle breakdown of the fused input-stage ResBlock (block 0, averaged over its tiles): runs the
detector on B clips with the profiling instantiation of resblock_in_tc_kernel."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wmb200
from wmb200 import _lib as L, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = 16000
dev = "cuda"
torch.manual_seed(0)
det = wmb200.Detector(16).to(dev).eval()      # the detector path has no LSTM (which shares the counter buffer)
s = 0.1 * torch.randn(B, T, device=dev)
lib = L.load()
blob = det.packed()
ws_n = lib.wm_detector_workspace_bytes(B, T)
ws = torch.empty(ws_n, dtype=torch.uint8, device=dev)
pr = torch.empty(B, T, device=dev); cp = torch.empty(B, device=dev); ml = torch.empty(B, 16, device=dev)
prof = torch.zeros(32, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
def run():
    L.check(lib.wm_detect_fwd(blob.data_ptr(), s.data_ptr(), None, pr.data_ptr(), cp.data_ptr(), ml.data_ptr(), None,
                              ws.data_ptr(), ws_n, B, T, 17, st), "detect")
run(); torch.cuda.synchronize()
lib.wm_debug_lstm_profile(prof.data_ptr())
prof.zero_()
run(); torch.cuda.synchronize()
lib.wm_debug_lstm_profile(None)
p = prof.cpu().double()
n = max(p[15].item(), 1)
names = ["prod wait a_empty", "prod build A", "mma wait a_full", "mma issue gemm1", "mma wait u_full", "mma issue conv2",
         "g1 wait d1_full", "g1 wait u_empty", "g1 u part", "g1 wait d2_empty", "g1 residual st + arrive",
         "g2 wait d2_full", "g2 until d2 release", "g2 rest"]
print(json.dumps({"B": B, "tiles_block0": n, "phases_cycles_per_tile": {k: round(p[i].item() / n, 1) for i, k in enumerate(names)}}, indent=1))
