"""CPU oracle for main14b_2's Generator / Detector (py/main14b_2.py:83-224).  TEST INFRASTRUCTURE ONLY.

Functional torch-CPU-fp32 restatement over a state dict (reference key names).  Pinned against outputs of the
reference's own class definitions executed in the build container (tests/golden/make_golden_14b2.py ->
tests/golden/main14b2_io.npz); the reference ships no weights or golden vectors for this model, so the
fixtures use seeded constructions (torch.manual_seed) whose parameters the drop-in modules reproduce draw by
draw.  Only tests/ may import this file.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]
STRIDES = (2, 4, 5, 8)          # py/main14b_2.py:44


def residual_block(x: Tensor, sd: SD, p: str, stride: int) -> Tensor:
    """py/main14b_2.py:97-105."""
    out = F.elu(F.conv1d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=stride, padding=1))
    out = F.conv1d(out, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    res = F.conv1d(x, sd[p + ".skip_conv.weight"], sd[p + ".skip_conv.bias"], stride=stride) \
        if (p + ".skip_conv.weight") in sd else x
    return F.elu(out + res)


def _fit(y: Tensor, T: int) -> Tensor:
    if y.shape[-1] > T:
        return y[:, :, :T]
    return F.pad(y, (0, T - y.shape[-1])) if y.shape[-1] < T else y


def _up(x: Tensor, sd: SD, p: str, strides: Sequence[int]) -> Tensor:
    for i, st in enumerate(reversed(list(strides))):
        x = F.conv_transpose1d(x, sd[f"{p}.{2 * i}.weight"], sd[f"{p}.{2 * i}.bias"], stride=st, padding=st // 2)
        x = residual_block(x, sd, f"{p}.{2 * i + 1}", 1)
    return x


def generator_forward(sd: SD, s: Tensor, message: Optional[Tensor] = None, strides: Sequence[int] = STRIDES) -> Tensor:
    """py/main14b_2.py:155-182."""
    T = s.shape[-1]
    x = F.conv1d(s, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
    for i, st in enumerate(strides):
        x = residual_block(x, sd, f"encoder_blocks.{i}", st)
    x = F.linear(x.transpose(1, 2), sd["proj.weight"], sd["proj.bias"])
    if message is not None:
        x = x + sd["E.weight"][message].unsqueeze(1)
    B, H = x.shape[0], x.shape[-1]
    w = [sd[f"lstm.{k}_l{l}"] for l in range(2) for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    z = x.new_zeros(2, B, H)
    x, _, _ = torch._VF.lstm(x, (z, z), w, True, 2, 0.0, False, False, True)
    x = F.conv1d(x.transpose(1, 2), sd["final_conv_enc.weight"], sd["final_conv_enc.bias"], padding=3)
    x = _up(x, sd, "decoder_blocks", strides)
    return _fit(F.conv1d(x, sd["final_conv_dec.weight"], sd["final_conv_dec.bias"], padding=3), T)


def detector_forward(sd: SD, x: Tensor, strides: Sequence[int] = STRIDES) -> Tensor:
    """py/main14b_2.py:209-224: raw logits (B, 1 + bits, T)."""
    T = x.shape[-1]
    x = F.conv1d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
    for i, st in enumerate(strides):
        x = residual_block(x, sd, f"encoder_blocks.{i}", st)
    x = _up(x, sd, "upsample_blocks", strides)
    return _fit(F.conv1d(x, sd["final_conv.weight"], sd["final_conv.bias"], padding=3), T)
