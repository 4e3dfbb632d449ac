#!/usr/bin/env python
"""Gradient error of the whole training step against the fp64 oracle, next to PyTorch's own fp32 autograd on the same
GPU, for the convolution variants of the training path (WMB200_TRAIN_FP32=1: fp32 FMA kernels; default: tcgen05)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wmb200
from wmb200 import train as TR
from oracle import wm_oracle_train as OT

def rel(a, b):
    b = b.to(a.device).to(torch.float64); a = a.to(torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

rows = []
for (B, T, seed) in [(2, 2400, 0), (3, 16000, 1), (8, 16000, 3)]:
    torch.manual_seed(seed)
    g, d = wmb200.Generator(message_bits=16), wmb200.Detector(message_bits=16)
    with torch.no_grad():
        for m in list(g.modules()) + list(d.modules()):
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(0.8 + 0.4 * torch.rand(64)); m.bias.copy_(0.1 * torch.randn(64))
        g.decoder[2].weight.mul_(0.05); g.decoder[2].bias.mul_(0.05)
    gsd, dsd = g.state_dict(), d.state_dict()
    tr = TR.Trainer(g.cuda(), d.cuda())
    o64 = OT.TrainOracle(gsd, dsd, dtype=torch.float64, device="cuda")
    o32 = OT.TrainOracle(gsd, dsd, dtype=torch.float32, device="cuda")
    t = torch.arange(T, device="cuda") / 16000.0
    s = 0.1 * torch.randn(B, T, device="cuda") + 0.2 * torch.sin(2 * np.pi * 300.0 * t)
    msg = torch.randint(0, 65536, (B,), device="cuda")
    want, base = o64.step(s, msg), o32.step(s, msg)
    got = tr.forward_backward(s, msg)
    gg, dg = tr.grad_dicts()
    worst = {"ours": 0.0, "torch32": 0.0, "ratio": 0.0, "key": None}
    for name, ours, w64, w32 in (("g", gg, want["g_grads"], base["g_grads"]), ("d", dg, want["d_grads"], base["d_grads"])):
        for k, v in w64.items():
            if k.endswith(("block.0.bias", "block.3.bias")):
                continue
            e_o, e_t = rel(ours[k], v), rel(w32[k], v)
            r = e_o / (3 * e_t + 1e-4)
            if r > worst["ratio"]:
                worst = {"ours": e_o, "torch32": e_t, "ratio": r, "key": name + "." + k}
    losses = {k: (abs(float(got[k]) - float(want[k])), abs(float(base[k]) - float(want[k]))) for k in ("l1", "mel", "loud", "loc", "bce", "hf")}
    rows.append({"B": B, "T": T, "mode": "fp32" if os.environ.get("WMB200_TRAIN_FP32") == "1" else "tcgen05",
                 "worst_gate_ratio (<1 passes)": round(worst["ratio"], 3), "worst_key": worst["key"],
                 "err_ours": worst["ours"], "err_torch_fp32": worst["torch32"],
                 "loss_err (ours, torch32)": {k: (float("%.2e" % a), float("%.2e" % b)) for k, (a, b) in losses.items()}})
    print(json.dumps(rows[-1]), flush=True)
