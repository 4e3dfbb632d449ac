"""Golden vectors of ONE FULL training iteration from the reference's own classes and loss objects (build container):

    python tests/golden/make_golden_train_full.py

Generator, Detector, MultiScaleMelLoss, TFLoudnessLoss, fir_lowpass / clamp_peak / limit_rms / high_freq_penalty are
AST-extracted from py/main16.py as in make_golden.py; the loop body below is py/main16.py:240-278 with the batch and
message fixed.  To keep the fixture small the 65536 x 64 embedding is zero except the rows the batch uses (rows
without gradient do not move in Adam's first step, so nothing is lost).  Stored: initial parameters, inputs, the
seven losses, every gradient (embedding: used rows) and the parameters after the Adam step -> train_full.npz.
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

B, T, MESSAGE_BITS = 2, 2400, 16


def main():
    ref = MG.extract(os.path.join(MG.REF, "py", "main16.py"), MG.WANT, MG.CONSTS | {"LR"})
    torch.manual_seed(21)
    generator, detector = ref.Generator(message_bits=MESSAGE_BITS), ref.Detector(message_bits=MESSAGE_BITS)
    g = torch.Generator().manual_seed(22)
    message = torch.tensor([40000, 5])
    with torch.no_grad():
        for m in list(generator.modules()) + list(detector.modules()):
            if isinstance(m, nn.BatchNorm1d):
                m.weight.copy_(0.8 + 0.4 * torch.rand(64, generator=g))
                m.bias.copy_(0.1 * torch.randn(64, generator=g))
        rows = generator.embedding.weight[message].clone()
        generator.embedding.weight.zero_()
        generator.embedding.weight[message] = rows
        # a head scale that puts delta around the RMS cap so clamp and limiter are both active
        generator.train()
        d = generator(0.1 * torch.randn(B, 1, T, generator=g), message)
        k = 0.008 / d.pow(2).mean().sqrt().item()
        generator.decoder[2].weight.mul_(k)
        generator.decoder[2].bias.mul_(k)
        for m in generator.modules():               # undo the running-stat update of the probe pass
            if isinstance(m, nn.BatchNorm1d):
                m.reset_running_stats()
    generator.train(); detector.train()
    out = {}
    for tag, mod in (("g", generator), ("d", detector)):
        for k_, v in mod.state_dict().items():
            if k_ == "embedding.weight":
                out["init.g.embedding.rows"] = v[message].numpy().copy()
            else:
                out[f"init.{tag}.{k_}"] = v.detach().numpy().copy()
    losses = {"mel": ref.MultiScaleMelLoss(), "loud": ref.TFLoudnessLoss()}
    optimizer = torch.optim.Adam(list(generator.parameters()) + list(detector.parameters()), lr=ref.LR)
    t = torch.arange(T) / 16000.0
    s = (0.1 * torch.randn(B, 1, T, generator=g) + 0.2 * torch.sin(2 * np.pi * 300.0 * t)).float()
    device = torch.device("cpu")
    # ---- py/main16.py:242-278 ----
    optimizer.zero_grad()
    delta = generator(s, message)
    delta = ref.fir_lowpass(delta)
    delta = ref.clamp_peak(delta)
    delta = ref.limit_rms(delta)
    s_w = s + delta
    combined = torch.cat([s_w, s], dim=0)
    logits = detector(combined)
    detection_logits = logits[:, :, 0]
    decode_logits = logits[:B, :, 1:]
    target_detection = torch.cat([torch.ones(B, s.shape[-1], device=device), torch.zeros(B, s.shape[-1], device=device)], dim=0)
    loc_loss = F.binary_cross_entropy_with_logits(detection_logits, target_detection)
    bitmask = (1 << torch.arange(MESSAGE_BITS, device=device))
    target_bits = ((message.unsqueeze(1) & bitmask) > 0).float()
    target_bits = target_bits.unsqueeze(1).expand(-1, s.shape[-1], -1)
    bce = F.binary_cross_entropy_with_logits(decode_logits, target_bits)
    l1 = F.l1_loss(delta, torch.zeros_like(delta))
    mel = losses["mel"](s, s_w)
    loud = losses["loud"](s, s_w)
    hf_penalty = ref.high_freq_penalty(delta)
    raw_loss = l1 + mel + loud + loc_loss + bce
    loss = (ref.LAMBDA_L1 * l1 + ref.LAMBDA_MSSPEC * mel + ref.LAMBDA_LOUD * loud +
            ref.LAMBDA_LOC * loc_loss + ref.LAMBDA_DEC * bce + ref.HF_PENALTY_W * hf_penalty)
    loss.backward()
    for tag, mod in (("g", generator), ("d", detector)):
        for k_, p in mod.named_parameters():
            if k_ == "embedding.weight":
                out["grad.g.embedding.rows"] = p.grad[message].numpy().copy()
                rest = p.grad.clone(); rest[message] = 0
                assert float(rest.abs().max()) == 0.0
            else:
                out[f"grad.{tag}.{k_}"] = p.grad.detach().numpy().copy()
    optimizer.step()
    for tag, mod in (("g", generator), ("d", detector)):
        for k_, v in mod.state_dict().items():
            if k_ == "embedding.weight":
                out["final.g.embedding.rows"] = v[message].detach().numpy().copy()
            else:
                out[f"final.{tag}.{k_}"] = v.detach().numpy().copy()
    out["s"] = s[:, 0].numpy().copy()
    out["message"] = message.numpy().copy()
    out["s_w"] = s_w.detach()[:, 0].numpy().copy()
    out["losses"] = np.array([l1.item(), mel.item(), loud.item(), loc_loss.item(), bce.item(), hf_penalty.item(),
                              loss.item(), raw_loss.item()], dtype=np.float64)
    path = os.path.join(HERE, "train_full.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", out["losses"], "delta rms", delta.pow(2).mean(dim=(1, 2)).sqrt())


if __name__ == "__main__":
    main()
