"""Two training iterations at batch B (default 16) for the ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_one_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wmb200
from wmb200 import train as TR
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
tr = TR.Trainer(wmb200.Generator(message_bits=16).cuda(), wmb200.Detector(message_bits=16).cuda())
s = 0.1 * torch.randn(B, 16000, device="cuda")
msg = torch.randint(0, 65536, (B,), device="cuda")
for _ in range(2):
    out = tr.step(s, msg)
torch.cuda.synchronize()
print({k: round(float(v), 5) for k, v in out.items()})
