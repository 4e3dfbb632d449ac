"""Detector pass with and without the fused majority-vote epilogue, CUDA-event times (developer A/B)."""
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import wmb200
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 16000
torch.manual_seed(0)
det = wmb200.Detector(16).cuda().eval()
s = 0.1 * torch.randn(B, 1, T, device="cuda")
out = {"lib": os.environ.get("WMB200_LIB", "default")}
for votes in (False, True):
    for _ in range(2):
        r = det.detect(s, want_probs=True, want_votes=votes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = det.detect(s, want_probs=True, want_votes=votes)
    e1.record(); torch.cuda.synchronize()
    out["votes" if votes else "plain"] = e0.elapsed_time(e1) / 5
    if votes:
        out["vote_frac_sum"] = float(r["vote_frac"].double().sum())
print(json.dumps(out))
