""""training path additionally under autograd".  The body of the reference's train_one_epoch
(py/main16.py:238-278) runs AS WRITTEN on wmb200's modules, helper functions and loss objects — train-mode forward,
`loss.backward()`, `torch.optim.Adam.step()` — and must agree with `wmb200.Trainer.step` (the one-call fast path that
is pinned on the reference's golden training vectors in test_train_full.py) on the same batch."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import wmb200
from wmb200 import (HF_PENALTY_W, LAMBDA_DEC, LAMBDA_L1, LAMBDA_LOC, LAMBDA_LOUD, LAMBDA_MSSPEC, MESSAGE_BITS,
                    MultiScaleMelLoss, TFLoudnessLoss, clamp_peak, fir_lowpass, high_freq_penalty, limit_rms)

pytestmark = pytest.mark.gpu
device = "cuda"


def reference_loop_body(generator, detector, s, message, optimizer, losses):
    """py/main16.py:238-278, the statements between `for s in train_loader` and the loss bookkeeping, unmodified
    (the random message is passed in so that both paths see the same one)."""
    B = s.size(0)
    optimizer.zero_grad()

    delta = generator(s, message)
    delta = fir_lowpass(delta)
    delta = clamp_peak(delta)
    delta = limit_rms(delta)
    s_w = s + delta
    combined = torch.cat([s_w, s], dim=0)
    logits = detector(combined)

    detection_logits = logits[:, :, 0]
    decode_logits = logits[:B, :, 1:]

    target_detection = torch.cat([
        torch.ones(B, s.shape[-1], device=device),
        torch.zeros(B, s.shape[-1], device=device)
    ], dim=0)

    loc_loss = F.binary_cross_entropy_with_logits(detection_logits, target_detection)
    bitmask = (1 << torch.arange(MESSAGE_BITS, device=device))
    target_bits = ((message.unsqueeze(1) & bitmask) > 0).float()
    target_bits = target_bits.unsqueeze(1).expand(-1, s.shape[-1], -1)
    bce = F.binary_cross_entropy_with_logits(decode_logits, target_bits)

    l1 = F.l1_loss(delta, torch.zeros_like(delta))
    mel = losses["mel"](s, s_w)
    loud = losses["loud"](s, s_w)

    hf_penalty = high_freq_penalty(delta)

    raw_loss = l1 + mel + loud + loc_loss + bce

    loss = (LAMBDA_L1 * l1 + LAMBDA_MSSPEC * mel + LAMBDA_LOUD * loud +
            LAMBDA_LOC * loc_loss + LAMBDA_DEC * bce + HF_PENALTY_W * hf_penalty)
    loss.backward()
    optimizer.step()
    return {"total": loss, "raw_total": raw_loss, "l1": l1, "mel": mel, "loud": loud, "loc": loc_loss, "bce": bce,
            "hf": hf_penalty}


@pytest.mark.parametrize("B,T", [(3, 4000), (2, 16000)])
def test_reference_loop_body_runs_under_autograd_and_matches_trainer(B, T):
    torch.manual_seed(B * 100 + T)
    generator = wmb200.Generator(message_bits=MESSAGE_BITS).to(device)
    detector = wmb200.Detector(message_bits=MESSAGE_BITS).to(device)
    g0, d0 = copy.deepcopy(generator.state_dict()), copy.deepcopy(detector.state_dict())
    s = (0.1 * torch.randn(B, 1, T, device=device)).clamp(-0.99, 0.99)
    message = torch.randint(0, 2 ** MESSAGE_BITS, (B,), device=device)

    # --- the fast path: one C-ABI call per iteration ---
    tr = wmb200.Trainer(generator, detector)
    ref = tr.forward_backward(s, message)
    gg_ref, dg_ref = tr.grad_dicts()
    tr.apply()
    gsd_ref, dsd_ref = tr.state_dicts()

    # --- the reference's loop body on the modules themselves ---
    generator.train()
    detector.train()
    optimizer = torch.optim.Adam(list(generator.parameters()) + list(detector.parameters()), lr=1e-3)   # py/main16.py:504
    out = reference_loop_body(generator, detector, s, message, optimizer,
                              {"mel": MultiScaleMelLoss(), "loud": TFLoudnessLoss()})
    for k in ("l1", "mel", "loud", "loc", "bce", "hf", "total", "raw_total"):
        a, b = float(out[k].detach()), float(ref[k])
        assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (k, a, b)

    # gradients: the same kernels in the same order -> agreement far below the fp32-vs-fp64 distance of either path.
    # (convolution biases in front of a BatchNorm have an analytically zero gradient: both paths hold round-off there)
    named = dict(list(("g." + k, v) for k, v in generator.named_parameters()) +
                 list(("d." + k, v) for k, v in detector.named_parameters()))
    refs = dict(list(("g." + k, v) for k, v in gg_ref.items()) + list(("d." + k, v) for k, v in dg_ref.items()))
    checked = 0
    for name, p in named.items():
        if name.endswith(("block.0.bias", "block.3.bias")):
            continue
        g, r = p.grad.float(), refs[name].to(device)
        scale = float(r.abs().max())
        assert float((g - r).abs().max()) <= 1e-4 * max(scale, 1e-6), (name, float((g - r).abs().max()), scale)
        checked += 1
    assert checked >= 30

    # one Adam step from the same start.  Adam's first step moves every weight by lr * sign(g): the two paths add the
    # residual-branch gradient in a different order (conv epilogue vs autograd's accumulation), so an element whose
    # gradient is round-off noise can take the opposite sign (2 * lr apart); everything else lands on the same value.
    # BatchNorm running statistics come from the forward pass, which is the same kernels in the same order.
    for sd_new, sd_ref in ((generator.state_dict(), gsd_ref), (detector.state_dict(), dsd_ref)):
        for k, v in sd_ref.items():
            if k.endswith(("block.0.bias", "block.3.bias", "num_batches_tracked")):
                continue
            dlt = (sd_new[k].float() - v.to(device)).abs()
            if "running" in k:
                assert float(dlt.max()) <= 1e-6 * max(1.0, float(v.abs().max())), k
            else:
                assert float(dlt.max()) <= 2.1e-3, k
                assert float(dlt.median()) <= 1e-6, k
    assert int(generator.encoder[1].block[1].num_batches_tracked) == 1

    # back to eval: the fused inference path picks up the updated parameters
    generator.eval()
    detector.eval()
    with torch.no_grad():
        r = wmb200.embed_detect(generator, detector, s, message)
    assert torch.isfinite(r["probs"]).all()


def test_eval_entry_points_refuse_train_mode_modules():
    gen = wmb200.Generator(16).to(device)
    det = wmb200.Detector(16).to(device)
    gen.train(); det.train()
    s = torch.zeros(1, 1, 2000, device=device)
    with pytest.raises((RuntimeError, NotImplementedError)):
        wmb200.embed_detect(gen, det, s, torch.zeros(1, dtype=torch.int64, device=device))
    with pytest.raises(RuntimeError):
        det.detect(s)
