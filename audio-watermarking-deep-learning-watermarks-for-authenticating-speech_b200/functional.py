"""The reference's helper functions by name (py/main16.py:53-72) plus the batched
embed+detect unit (forward of evaluate_model, py/main16.py:378-403), on libwmb200."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from . import autograd as AG
from . import ops, packing

SAMPLE_RATE = 16000      # py/main16.py:30
AUDIO_LEN = 16000        # py/main16.py:31
MESSAGE_BITS = 16        # py/main16.py:34
MAX_RMS = 0.005          # py/main16.py:29
LAMBDA_L1 = 1.0          # py/main16.py:38
LAMBDA_MSSPEC = 4.0      # py/main16.py:39
LAMBDA_LOUD = 20.0       # py/main16.py:40
LAMBDA_LOC = 10.0        # py/main16.py:41
LAMBDA_DEC = 1.0         # py/main16.py:42
HF_PENALTY_W = 5.0       # py/main16.py:43

_fir_cache = {}


def fir_taps_on(device, cutoff: float = 4000.0, taps: int = 101) -> torch.Tensor:
    key = (str(device), float(cutoff), int(taps))
    if key not in _fir_cache:
        _fir_cache[key] = packing.fir_taps(cutoff, taps, SAMPLE_RATE).to(device)
    return _fir_cache[key]


def _b1t(d: torch.Tensor) -> torch.Tensor:
    if d.dim() != 3 or d.shape[1] != 1:
        raise ValueError(f"expected (B, 1, T), got {tuple(d.shape)}")
    return d[:, 0, :]


def fir_lowpass(delta: torch.Tensor, cutoff: float = 4000, taps: int = 101) -> torch.Tensor:
    """py/main16.py:53-64."""
    if taps != 101:
        raise ValueError("wmb200 implements the reference's 101-tap filter")
    if AG.needs_graph(delta):
        return AG.Postprocess.apply(ops._req(_b1t(delta), "delta"), fir_taps_on(delta.device, cutoff, taps), L.POST_FIR,
                                    ops.PEAK, ops.MAX_RMS, ops.RMS_EPS).unsqueeze(1)
    d, _, _ = ops.postprocess(_b1t(delta), None, fir_taps_on(delta.device, cutoff, taps), L.POST_FIR, True, False)
    return d.unsqueeze(1)


def clamp_peak(d: torch.Tensor, thr: float = 0.02) -> torch.Tensor:
    """py/main16.py:66-67."""
    if AG.needs_graph(d):
        return AG.Postprocess.apply(ops._req(_b1t(d), "d"), None, L.POST_CLAMP, thr, ops.MAX_RMS, ops.RMS_EPS).unsqueeze(1)
    out, _, _ = ops.postprocess(_b1t(d), None, None, L.POST_CLAMP, True, False, peak=thr)
    return out.unsqueeze(1)


def limit_rms(delta: torch.Tensor, max_rms: float = MAX_RMS, eps: float = 1e-8) -> torch.Tensor:
    """py/main16.py:69-72."""
    if AG.needs_graph(delta):
        return AG.Postprocess.apply(ops._req(_b1t(delta), "delta"), None, L.POST_RMS, ops.PEAK, max_rms, eps).unsqueeze(1)
    out, _, _ = ops.postprocess(_b1t(delta), None, None, L.POST_RMS, True, False, max_rms=max_rms, eps=eps)
    return out.unsqueeze(1)


def postprocess_delta(delta: torch.Tensor, s: Optional[torch.Tensor] = None):
    """limit_rms(clamp_peak(fir_lowpass(delta))) and, when `s` is given, s + delta in one
    kernel (py/main16.py:245-248).  Returns (delta, s_w or None)."""
    d, sw, _ = ops.postprocess(_b1t(delta), _b1t(s) if s is not None else None, fir_taps_on(delta.device),
                               L.POST_ALL, True, s is not None)
    return d.unsqueeze(1), (sw.unsqueeze(1) if sw is not None else None)


@torch.no_grad()
@ops.nvtx("embed_detect")
def embed_detect(generator, detector, s: torch.Tensor, message: Optional[torch.Tensor],
                 postprocess: bool = True, want_delta: bool = True, want_probs: bool = True,
                 want_votes: bool = True, want_rms: bool = False) -> dict:
    """One batch of clips through G -> fir/clamp/rms -> s + delta -> D -> heads.

    s (B,1,T) fp32 on the GPU, message (B,) int64.  Returns device tensors:
    delta (B,1,T), s_w (B,1,T), probs (B,T), clip_prob (B,), msg_logits (B,bits),
    bits_mean (sign of the mean logit, py/main16.py:1146,1185) and bits_vote
    (per-sample majority, py/main16.py:398)."""
    if generator.training or detector.training:
        raise NotImplementedError("embed_detect is the eval-mode path; call .eval() on both modules")
    x = _b1t(s)
    use_msg = generator.message_bits > 0 and message is not None
    r = ops.embed_detect_fwd(generator.packed(), generator.embedding_table() if use_msg else None,
                             detector.packed(), fir_taps_on(s.device), message if use_msg else None, x,
                             detector.nout, L.POST_ALL if postprocess else 0, want_delta, want_probs,
                             want_votes, want_rms)
    out = {"s_w": r["s_w"].unsqueeze(1), "probs": r["probs"], "clip_prob": r["clip_prob"],
           "msg_logits": r["msg_logits"], "bits_mean": r["msg_logits"] > 0, "delta_rms": r["delta_rms"]}
    out["delta"] = r["delta"].unsqueeze(1) if r["delta"] is not None else None
    if r["vote_frac"] is not None:
        out["vote_frac"] = r["vote_frac"]
        out["bits_vote"] = r["vote_frac"] > 0.5
    return out


def bit_targets(message: torch.Tensor, bits: int = MESSAGE_BITS) -> torch.Tensor:
    """py/main16.py:261-262 — bit j of the id <-> logits channel 1+j (LSB first)."""
    return ((message.unsqueeze(1) & (1 << torch.arange(bits, device=message.device))) > 0).float()
