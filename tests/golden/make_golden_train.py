"""Golden vectors of the Detector's training step FROM THE REFERENCE's own classes (build container only):

    python tests/golden/make_golden_train.py

The reference's Detector / ResBlock (py/main16.py:112-125,170-186) are AST-extracted as in make_golden.py, put in
train mode and stepped twice exactly as train_one_epoch does for the detector (py/main16.py:249-264,275-278, Adam
of :504): logits = detector(cat(s_w, s)); loss = LAMBDA_LOC * loc + LAMBDA_DEC * bce; backward; Adam(lr=LR).
Stored: the initial state dict, the inputs, per-step losses, the step-1 gradients (parameters and input) and the
state dict after two steps -> tests/golden/train_step.npz.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

MESSAGE_BITS = 16
B, T = 3, 800


def main():
    ref = MG.extract(os.path.join(MG.REF, "py", "main16.py"), MG.WANT, MG.CONSTS | {"LR"})
    torch.manual_seed(7)
    det = ref.Detector(message_bits=MESSAGE_BITS)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for m in det.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.running_mean.copy_(0.2 * torch.randn(64, generator=g))
                m.running_var.copy_(0.5 + torch.rand(64, generator=g))
                m.weight.copy_(0.8 + 0.4 * torch.rand(64, generator=g))
                m.bias.copy_(0.1 * torch.randn(64, generator=g))
    det.train()
    out = {"init." + k: v.detach().numpy().copy() for k, v in det.state_dict().items()}
    opt = torch.optim.Adam(det.parameters(), lr=ref.LR)
    for step in range(2):
        s = 0.1 * torch.randn(B, 1, T, generator=g)
        delta = 0.01 * torch.randn(B, 1, T, generator=g)
        message = torch.randint(0, 2 ** MESSAGE_BITS, (B,), generator=g)
        s_w = s + delta
        combined = torch.cat([s_w, s], dim=0).requires_grad_(True)
        opt.zero_grad()
        # ---- py/main16.py:250-264, 275-278 (detector terms) ----
        logits = det(combined)
        detection_logits = logits[:, :, 0]
        decode_logits = logits[:B, :, 1:]
        target_detection = torch.cat([torch.ones(B, T), torch.zeros(B, T)], dim=0)
        loc_loss = F.binary_cross_entropy_with_logits(detection_logits, target_detection)
        bitmask = (1 << torch.arange(MESSAGE_BITS))
        target_bits = ((message.unsqueeze(1) & bitmask) > 0).float()
        target_bits = target_bits.unsqueeze(1).expand(-1, T, -1)
        bce = F.binary_cross_entropy_with_logits(decode_logits, target_bits)
        loss = ref.LAMBDA_LOC * loc_loss + ref.LAMBDA_DEC * bce
        loss.backward()
        if step == 0:
            for k, p in det.named_parameters():
                out["grad1." + k] = p.grad.detach().numpy().copy()
            out["grad1.input"] = combined.grad[:, 0].numpy().copy()
        opt.step()
        out[f"x{step}"] = combined.detach()[:, 0].numpy().copy()
        out[f"message{step}"] = message.numpy().copy()
        out[f"losses{step}"] = np.array([loc_loss.item(), bce.item()], dtype=np.float64)
    out.update({"final." + k: v.detach().numpy().copy() for k, v in det.state_dict().items()})
    out["hyper"] = np.array([ref.LR, ref.LAMBDA_LOC, ref.LAMBDA_DEC], dtype=np.float64)
    path = os.path.join(HERE, "train_step.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", {k: out[k] for k in ("losses0", "losses1")})


if __name__ == "__main__":
    main()
