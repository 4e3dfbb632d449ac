"""Small shapes of the round's new kernels for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wmb200
from wmb200 import train as TR
from wmb200 import main14b_2 as M
torch.manual_seed(0)
tr = TR.Trainer(wmb200.Generator(message_bits=16).cuda(), wmb200.Detector(message_bits=16).cuda())
s = 0.1 * torch.randn(2, 1300, device="cuda")
msg = torch.tensor([5, 5], device="cuda")
out = tr.step(s, msg)
torch.cuda.synchronize()
print("train", {k: round(float(v), 4) for k, v in out.items()})
G, D = M.Generator().cuda().eval(), M.Detector().cuda().eval()
x = 0.1 * torch.randn(2, 1, 1777, device="cuda")
with torch.no_grad():
    y = D(x + G(x, torch.tensor([1, 2], device="cuda")))
torch.cuda.synchronize()
print("main14b_2", tuple(y.shape), float(y.abs().mean()))
