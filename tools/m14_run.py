"""One Generator + Detector pass of main14b_2 at batch B (for ncu launch lists / timing): python tools/m14_run.py B [reps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from wmb200 import main14b_2 as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(0)
G, D = M.Generator().cuda().eval(), M.Detector().cuda().eval()
s = (0.1 * torch.randn(B, 1, 16000, device="cuda")).clamp(-0.99, 0.99)
msg = torch.randint(0, 65536, (B,), device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = D(s + G(s, msg))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        y = D(s + G(s, msg))
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("B", B, "ms/step", e0.elapsed_time(e1) / reps, "host issue ms", t_issue * 1e3 / reps, "host ms", (time.perf_counter() - t0) * 1e3 / reps,
          "clip-s/s", B * 1e3 * reps / e0.elapsed_time(e1), "finite", bool(torch.isfinite(y).all()))
    if len(sys.argv) > 3 and sys.argv[3] == "torch":
        # stock PyTorch eager (cuDNN) on the same GPU through the oracle's functional restatement of the reference
        # (informational: TF32 as the reference sets it, and off)
        from oracle import wm_oracle_14b2 as O
        gsd, dsd = G.state_dict(), D.state_dict()
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for _ in range(2):
                yt = O.detector_forward(dsd, s + O.generator_forward(gsd, s, msg))
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                yt = O.detector_forward(dsd, s + O.generator_forward(gsd, s, msg))
            e1.record()
            torch.cuda.synchronize()
            print("B", B, "torch eager tf32=%s ms/step" % tf32, e0.elapsed_time(e1) / reps, "clip-s/s",
                  B * 1e3 * reps / e0.elapsed_time(e1), "max |logit diff| vs wmb200", float((yt - y).abs().max()))
    if len(sys.argv) > 3 and sys.argv[3] == "graph":
        ge = M.GraphedEmbedDetect(G, D, B, 16000)
        for _ in range(2):
            ge(s, msg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            ge(s, msg)
        e1.record()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        print("B", B, "graph replay ms/step", e0.elapsed_time(e1) / reps, "host issue ms", t_issue * 1e3 / reps,
              "clip-s/s", B * 1e3 * reps / e0.elapsed_time(e1))
