"""Training-step throughput (BASELINE config 4): wmb200.Trainer.step next to PyTorch eager autograd of the same step
(oracle definitions, fp32, cuDNN) on the same GPU.   python tools/train_bench.py [--batch 16] [--steps 5]
Also prints the two LSTM kernels alone (the sequential part of the step)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import wmb200
from wmb200 import train as TR


def timed(fn, steps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[16, 64])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-torch", action="store_true")
    a = ap.parse_args()
    T = 16000
    out = []
    for B in a.batch:
        torch.manual_seed(0)
        g, d = wmb200.Generator(message_bits=16), wmb200.Detector(message_bits=16)
        gsd, dsd = g.state_dict(), d.state_dict()
        tr = TR.Trainer(g.cuda(), d.cuda())
        s = 0.1 * torch.randn(B, T, device="cuda")
        msg = torch.randint(0, 65536, (B,), device="cuda")
        ms = timed(lambda: tr.step(s, msg), a.steps)
        row = {"batch": B, "wmb200_ms": round(ms, 2), "wmb200_it_s": round(1000 / ms, 2),
               "wmb200_clips_s": round(B * 1000 / ms, 1)}
        x = torch.randn(B, T, 64, device="cuda")
        p = {k: v.cuda() for k, v in gsd.items() if k.startswith("lstm.")}
        h, saved = TR.lstm_train_fwd(x, p["lstm.weight_ih_l0"], p["lstm.weight_hh_l0"], p["lstm.bias_ih_l0"], p["lstm.bias_hh_l0"])
        row["lstm_fwd_ms"] = round(timed(lambda: TR.lstm_train_fwd(x, p["lstm.weight_ih_l0"], p["lstm.weight_hh_l0"],
                                                                   p["lstm.bias_ih_l0"], p["lstm.bias_hh_l0"]), 3, 1), 2)
        row["lstm_bwd_ms"] = round(timed(lambda: TR.lstm_train_bwd(x, saved), 3, 1), 2)
        dt = TR.DetectorTrainer(d)
        xd = torch.cat([s, s])
        row["detector_step_ms"] = round(timed(lambda: dt.step(xd, msg), 3, 1), 2)
        if not a.no_torch:
            from oracle import wm_oracle_train as OT
            torch.backends.cudnn.allow_tf32 = True       # the reference's setting (py/main16.py:44)
            torch.backends.cuda.matmul.allow_tf32 = True
            o = OT.TrainOracle(gsd, dsd, device="cuda")
            ms_t = timed(lambda: o.step(s, msg), a.steps)
            row.update(torch_eager_tf32_ms=round(ms_t, 2), torch_eager_it_s=round(1000 / ms_t, 2))
        out.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
