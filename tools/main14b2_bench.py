#!/usr/bin/env python
"""Throughput of the main14b_2 stack (BASELINE config 3: 8192 clips over 8 GPUs = 1024 per GPU), generic fp32
kernels.  Usage: python tools/main14b2_bench.py [clips_per_gpu]   (GPU box only)"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wmb200 import main14b_2 as M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = "cuda"
torch.manual_seed(0)
G, D = M.Generator().to(dev).eval(), M.Detector().to(dev).eval()
s = (0.1 * torch.randn(B, 1, 16000, device=dev)).clamp(-0.99, 0.99)
msg = torch.randint(0, 65536, (B,), device=dev)
def step():
    d = G(s, msg)
    return D(s + d)
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"model": "main14b_2", "clips": B, "ms_per_step": ms, "clip_s_per_s": B / ms * 1e3,
                  "algorithmic_gflop_per_clip": 4.540, "tflops": 4.540e9 * B / ms * 1e3 / 1e12}))
