"""One evaluate_model-style step (delta RMS + fused majority-vote bits) and one host-fed step, for ncu launch lists."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import wmb200
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 16000
gen = wmb200.Generator(16).cuda().eval(); det = wmb200.Detector(16).cuda().eval()
s = 0.1 * torch.randn(B, 1, T, device="cuda"); m = torch.randint(0, 65536, (B,), device="cuda")
for _ in range(2):
    r = wmb200.embed_detect(gen, det, s, m, want_delta=False, want_probs=True, want_votes=True, want_rms=True)
torch.cuda.synchronize()
