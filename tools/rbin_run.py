import os, sys, torch
sys.path.insert(0, os.getcwd())
import wmb200
from wmb200 import _lib as L, ops
B, T = 4096, 16000
gen = wmb200.Generator(16).to("cuda").eval(); det = wmb200.Detector(16).to("cuda").eval()
s = 0.1 * torch.randn(B, 1, T, device="cuda")
with torch.no_grad():
    for _ in range(3):
        r = det.detect(s, want_probs=True, want_votes=False)
    r = det.detect(s, want_probs=False, want_votes=True)
torch.cuda.synchronize()
