"""The eval-mode twins of the reference's training loop, forward only (eval half):
`validate_one_epoch` (py/main16.py:297-364) and `evaluate_model` (py/main16.py:378-421), same signatures, same
returned dictionaries, same RNG consumption (one `torch.randint(0, 2**MESSAGE_BITS, (B,), device=device)` per batch).

Per batch the reference runs G, fir/clamp/rms, D on cat([s_w, s]) and a dozen reductions with a host sync each;
here the batch goes through the fused embed+detect kernels, the staged-FFT loss kernels and one detector call on
the clean clips, and only the final scalars leave the device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import functional as Fn
from . import ops
from .losses import MultiScaleMelLoss, TFLoudnessLoss, step_losses


def _progress(it, desc):
    try:
        from tqdm import tqdm
        return tqdm(it, desc=desc)
    except ImportError:
        return it


@torch.no_grad()
def validate_one_epoch(generator, detector, val_loader, losses, device):
    """py/main16.py:297-364: mean over batches of total / raw_total / l1 / mel / loud / loc / bce."""
    generator.eval()
    detector.eval()
    losses = losses or {"mel": MultiScaleMelLoss(), "loud": TFLoudnessLoss()}
    keys = ("total", "raw_total", "l1", "mel", "loud", "loc", "bce")
    acc = None
    n = 0
    for s in _progress(val_loader, "Validation Epoch"):
        s = s.to(device)
        message = torch.randint(0, 2 ** Fn.MESSAGE_BITS, (s.size(0),), device=device)
        r = step_losses(generator, detector, s, message, losses)
        vec = torch.stack([r[k].reshape(()) for k in keys])
        acc = vec if acc is None else acc + vec              # stays on the device: one read at the end
        n += 1
    if n == 0:
        raise ZeroDivisionError("validate_one_epoch: empty loader")          # the reference divides by num_batches
    out = (acc / n).cpu().tolist()
    return dict(zip(keys, out))


@torch.no_grad()
def evaluate_model(generator, detector, dataloader, device, threshold=0.5):
    """py/main16.py:378-421: mean clip probability of watermarked and clean clips, majority-vote bit accuracy and
    watermark RMS over a loader."""
    generator.eval()
    detector.eval()
    probs_wm, probs_clean, bit_accs, rms_all = [], [], [], []
    for s in _progress(dataloader, "Evaluating"):
        s = s.to(device)
        B = s.size(0)
        message = torch.randint(0, 2 ** Fn.MESSAGE_BITS, (B,), device=device)
        r = Fn.embed_detect(generator, detector, s, message, want_delta=False, want_probs=False, want_votes=True,
                            want_rms=True)
        clean = detector.detect(s, want_probs=False, want_votes=False)
        probs_wm.append(r["clip_prob"])
        probs_clean.append(clean["clip_prob"])
        if detector.message_bits > 0:
            decoded = r["vote_frac"] > 0.5                                    # per-sample majority (:398)
            target = Fn.bit_targets(message, detector.message_bits) > 0.5
            bit_accs.append((decoded == target).float().mean(dim=1))
        rms_all.append(r["delta_rms"])
    cat = lambda xs: torch.cat(xs).cpu().numpy() if xs else np.zeros(0, dtype=np.float32)
    avg_real, avg_wm = float(np.mean(cat(probs_clean))), float(np.mean(cat(probs_wm)))
    avg_bit, avg_rms = float(np.mean(cat(bit_accs))) if bit_accs else float("nan"), float(np.mean(cat(rms_all)))
    print("\nEvaluation Results:")
    print(f"  Avg Detection Prob - Watermarked: {avg_wm:.4f}")
    print(f"  Avg Detection Prob - Clean:       {avg_real:.4f}")
    print(f"  Avg Bit Attribution Accuracy:     {avg_bit:.4f}")
    print(f"  Avg Watermark RMS:                {avg_rms:.6f}")
    return {"watermarked_prob": avg_wm, "clean_prob": avg_real, "bit_accuracy": avg_bit, "delta_rms": avg_rms}
