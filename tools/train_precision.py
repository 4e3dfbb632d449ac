"""Gradient error of the detector training step against an fp64 oracle: this library (fp32 CUDA kernels) next to
PyTorch's own fp32 autograd on the same GPU.  GPU box:  python tools/train_precision.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wmb200
from wmb200 import train as TR
from oracle import wm_oracle_train as OT
from tests import test_train as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for name, x, msg, sd in [("golden", torch.from_numpy(T.G["x0"]).cuda(), torch.from_numpy(T.G["message0"]).cuda(), T.sd_of("init.")),
                         ("random", 0.1 * torch.randn(8, 16000, device="cuda"), torch.randint(0, 65536, (4,), device="cuda"), None)]:
    det = wmb200.Detector(message_bits=16)
    if sd is not None:
        det.load_state_dict(sd)
    sd = det.state_dict()
    n_wm = msg.numel()
    o64 = OT.DetectorTrainOracle(sd, dtype=torch.float64, device="cuda").step(x, msg, n_wm, update=False)
    o32 = OT.DetectorTrainOracle(sd, dtype=torch.float32, device="cuda").step(x, msg, n_wm, update=False)
    tr = TR.DetectorTrainer(det.cuda())
    r = tr.step(x, msg, update=False, want_input_grad=True)
    gd = tr.grad_dict()
    print(f"== {name}: B2={x.shape[0]} T={x.shape[1]}   key: rel err wmb200 | torch fp32")
    for k in OT.PARAM_KEYS:
        if k.endswith(T.DEAD):
            continue
        print(f"  {k:28s} {T.rel(gd[k], o64['grads'][k]):.2e} | {T.rel(o32['grads'][k], o64['grads'][k]):.2e}")
    print(f"  {'d_input':28s} {T.rel(r['d_input'], o64['d_input']):.2e} | {T.rel(o32['d_input'], o64['d_input']):.2e}")
    print(f"  losses {float(r['loc']):.7f} {float(r['bce']):.7f} | {float(o32['loc']):.7f} {float(o32['bce']):.7f} | {float(o64['loc']):.7f} {float(o64['bce']):.7f}")
