"""The reference's training losses by name (py/main16.py:74-81, 192-217) and the loss terms of its
train / validation step (py/main16.py:252-276), forward only, on libwmb200's staged-FFT kernels.

Every function returns 0-dim fp32 tensors on the input's device; the scalars are produced by
fixed-order reductions, so repeated calls are bit-identical.  When autograd is recording on the
watermark-side input (delta / watermarked) the call goes through the matching autograd.Function
(backward = the staged-FFT adjoint kernels); the clean signal is data and receives no gradient,
as in the reference's loop.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import autograd as AG
from . import functional as Fn
from . import ops, packing

_mel_cache = {}


def _no_clean_grad(clean):
    if torch.is_grad_enabled() and clean is not None and clean.requires_grad:
        raise NotImplementedError("the clean signal is data in the reference's loop (py/main16.py:268-269): no gradient "
                                  "w.r.t. it is implemented; detach it")


def _bt(x: torch.Tensor, name: str) -> torch.Tensor:
    if x.dim() == 3 and x.shape[1] == 1:
        return x[:, 0, :]
    if x.dim() == 2:
        return x
    raise ValueError(f"{name}: expected (B, 1, T) or (B, T), got {tuple(x.shape)}")


def high_freq_penalty(delta: torch.Tensor, cutoff: float = 3_500, n_fft: int = 512) -> torch.Tensor:
    """py/main16.py:74-81: mean over (B, n_fft/2+1, frames) of |STFT(delta)| * [rfftfreq > cutoff]."""
    freqs = torch.fft.rfftfreq(n_fft, 1 / Fn.SAMPLE_RATE)
    above = torch.nonzero(freqs > cutoff).flatten()
    first_bin = int(above[0]) if above.numel() else n_fft // 2 + 1
    if AG.needs_graph(delta):
        return AG.HighFreqPenalty.apply(ops._req(_bt(delta, "delta"), "delta"), n_fft, first_bin)
    return ops.hf_penalty(_bt(delta, "delta"), n_fft, first_bin)


class MultiScaleMelLoss(nn.Module):
    """py/main16.py:192-202: L1 between log-mel spectrograms (MelSpectrogram(16000, 1024, 256, 64) + 1e-5)."""

    def __init__(self):
        super().__init__()
        self.n_fft, self.hop_length, self.n_mels = 1024, 256, 64

    def _tables(self, device):
        key = str(device)
        if key not in _mel_cache:
            fb, band = packing.mel_filterbank(self.n_fft // 2 + 1, self.n_mels, Fn.SAMPLE_RATE)
            _mel_cache[key] = (fb.to(device), band.to(device))
        return _mel_cache[key]

    def forward(self, clean: torch.Tensor, watermarked: torch.Tensor) -> torch.Tensor:
        _no_clean_grad(clean)
        fb, band = self._tables(clean.device)
        if AG.needs_graph(watermarked):
            return AG.MelLogL1.apply(ops._req(_bt(clean, "clean"), "clean"), ops._req(_bt(watermarked, "watermarked"), "watermarked"),
                                     fb, band, self.n_fft, self.hop_length)
        return ops.mel_log_l1(_bt(clean, "clean"), _bt(watermarked, "watermarked"), fb, band, self.n_fft,
                              self.hop_length)


class TFLoudnessLoss(nn.Module):
    """py/main16.py:204-217: mean of (|S_w| - |S_c|)^2 where |S_c| > 0.01, STFT 2048/512."""

    def __init__(self):
        super().__init__()
        self.win_size = 2048
        self.hop = 512

    def forward(self, clean: torch.Tensor, watermarked: torch.Tensor) -> torch.Tensor:
        _no_clean_grad(clean)
        if AG.needs_graph(watermarked):
            return AG.Loudness.apply(ops._req(_bt(clean, "clean"), "clean"), ops._req(_bt(watermarked, "watermarked"), "watermarked"),
                                     self.win_size, self.hop, 0.01)
        return ops.loudness(_bt(clean, "clean"), _bt(watermarked, "watermarked"), self.win_size, self.hop, 0.01)


def stft_magnitude(x: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """torch.stft(x, n_fft, hop, window=hann_window(n_fft), return_complex=True).abs() for x (B,T)."""
    return ops.stft_mag(_bt(x, "x"), n_fft, hop)


@torch.no_grad()
def step_losses(generator, detector, s: torch.Tensor, message: Optional[torch.Tensor],
                losses: Optional[dict] = None) -> dict:
    """Loss terms of one step of py/main16.py:238-276 with eval-mode modules (the reference's
    validate_one_epoch): G -> fir/clamp/rms -> s_w; D on cat([s_w, s]); L1, mel, loudness, detection BCE,
    bit BCE, HF penalty and the weighted total."""
    losses = losses or {"mel": MultiScaleMelLoss(), "loud": TFLoudnessLoss()}
    B = s.shape[0]
    delta = generator(s, message)
    delta, s_w = Fn.postprocess_delta(delta, s)
    logits = detector(torch.cat([s_w, s], dim=0))
    loc, bce = ops.bce_heads(logits, message, B)
    out = {"l1": ops.abs_mean(delta), "mel": losses["mel"](s, s_w), "loud": losses["loud"](s, s_w), "loc": loc,
           "bce": bce if bce is not None else torch.zeros((), device=s.device), "hf": high_freq_penalty(delta)}
    out["raw_total"] = out["l1"] + out["mel"] + out["loud"] + out["loc"] + out["bce"]
    out["total"] = (Fn.LAMBDA_L1 * out["l1"] + Fn.LAMBDA_MSSPEC * out["mel"] + Fn.LAMBDA_LOUD * out["loud"] +
                    Fn.LAMBDA_LOC * out["loc"] + Fn.LAMBDA_DEC * out["bce"] + Fn.HF_PENALTY_W * out["hf"])
    return out
