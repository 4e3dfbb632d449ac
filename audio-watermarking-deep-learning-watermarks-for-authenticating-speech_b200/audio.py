"""The data formats either side of the embed+detect path, on the GPU:
`Resample` / `resample` (torchaudio.transforms.Resample as the reference uses it, py/main16.py:985,1121),
`to_pcm16` / `from_pcm16` (py/main15.py:859-860), `lowpass_biquad` / `save_audio_pcm16` (py/main15.py:850-867,
main15c.ipynb cell 4) and `file_metrics` (py/main16.py:1030-1049)."""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib as L
from . import packing
from .ops import _req, _stream

_kern_cache = {}


def resample(x: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """x (..., T) fp32 on the GPU -> (..., ceil(new * T / orig)), torchaudio's default sinc_interp_hann resampler."""
    if int(orig_freq) == int(new_freq):
        return x
    lib = L.load()
    lead, T = x.shape[:-1], x.shape[-1]
    x2 = _req(x.reshape(-1, T), "x")
    key = (int(orig_freq), int(new_freq), str(x2.device))
    if key not in _kern_cache:
        kern, down, up, width = packing.resample_kernel(orig_freq, new_freq)
        _kern_cache[key] = (kern.to(x2.device), down, up, width)
    kern, down, up, width = _kern_cache[key]
    Tout = int(math.ceil(up * T / down))
    y = torch.empty(x2.shape[0], Tout, device=x2.device)
    L.check(lib.wm_resample_fwd(L.ptr(x2), L.ptr(kern), L.ptr(y), x2.shape[0], T, Tout, down, up, kern.shape[0], width,
                                _stream()), "wm_resample_fwd")
    return y.reshape(*lead, Tout)


class Resample(torch.nn.Module):
    """torchaudio.transforms.Resample(orig_freq, new_freq) for CUDA tensors."""

    def __init__(self, orig_freq: int = 16000, new_freq: int = 16000):
        super().__init__()
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        return resample(waveform, self.orig_freq, self.new_freq)


def to_pcm16(x: torch.Tensor) -> torch.Tensor:
    """(x.clamp(-1, 1) * 32767).to(torch.int16)  — py/main15.py:859-860."""
    lib = L.load()
    xc = _req(x, "x")
    q = torch.empty(xc.shape, dtype=torch.int16, device=xc.device)
    L.check(lib.wm_pcm16_quantize_fwd(L.ptr(xc), L.ptr(q), xc.numel(), _stream()), "wm_pcm16_quantize_fwd")
    return q


def from_pcm16(q: torch.Tensor, scale: float = 1.0 / 32768.0) -> torch.Tensor:
    """int16 PCM -> fp32 in [-1, 1) (what a 16-bit WAV loader returns)."""
    lib = L.load()
    qc = _req(q, "q", torch.int16)
    x = torch.empty(qc.shape, dtype=torch.float32, device=qc.device)
    L.check(lib.wm_pcm16_dequantize_fwd(L.ptr(qc), L.ptr(x), qc.numel(), scale, _stream()), "wm_pcm16_dequantize_fwd")
    return x


def file_metrics(original: torch.Tensor, watermarked: torch.Tensor, valid_len: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Rows of (B,T) waveforms -> (B,3) = watermark RMS, SI-SNR [dB], power ratio [dB]
    (py/main16.py:1030-1049, compute_si_snr :764-773), reduced on the device in double precision."""
    lib = L.load()
    s, w = _req(original, "original"), _req(watermarked, "watermarked")
    if s.shape != w.shape or s.dim() != 2:
        raise ValueError(f"file_metrics: expected two (B,T) tensors, got {tuple(s.shape)} and {tuple(w.shape)}")
    v = _req(valid_len, "valid_len", torch.int32) if valid_len is not None else None
    out = torch.empty(s.shape[0], 3, device=s.device)
    L.check(lib.wm_file_metrics_fwd(L.ptr(s), L.ptr(w), L.ptr(v), L.ptr(out), s.shape[0], s.shape[1], _stream()),
            "wm_file_metrics_fwd")
    return out


def compute_si_snr(s: torch.Tensor, s_hat: torch.Tensor, eps: float = 1e-8) -> float:
    """Scale-invariant SNR in dB, mean over the rows — py/main16.py:764-773.  s, s_hat: (B,T) (or (B,1,T) / (1,1,T) as
    the later cells of the reference pass them), on the GPU; one block per row, sums in double precision.  `eps` is
    the reference's default (the kernel's constant)."""
    if eps != 1e-8:
        raise ValueError("compute_si_snr: only the reference's eps = 1e-8 is built into the kernel")
    a = s.reshape(-1, s.shape[-1])
    b = s_hat.reshape(-1, s_hat.shape[-1])
    return float(file_metrics(a, b)[:, 1].mean())


def biquad(waveform: torch.Tensor, b0: float, b1: float, b2: float, a0: float, a1: float, a2: float,
           clamp: bool = True, want_pcm16: bool = False):
    """torchaudio.functional.biquad / lfilter for CUDA tensors: (..., T) fp32 -> same shape (and, with want_pcm16, the
    int16 codes `(y.clamp(-1, 1) * 32767).to(int16)` from the same pass).  Zero initial state, clamp as lfilter's
    default.  A chunked scan in double precision (wm_eval.cu)."""
    import ctypes as C
    lib = L.load()
    lead, T = waveform.shape[:-1], waveform.shape[-1]
    x2 = _req(waveform.reshape(-1, T), "waveform")
    rows = x2.shape[0]
    y = torch.empty_like(x2)
    q = torch.empty(x2.shape, dtype=torch.int16, device=x2.device) if want_pcm16 else None
    n = lib.wm_biquad_workspace_bytes(rows, T)
    ws = torch.empty(max(n, 256), dtype=torch.uint8, device=x2.device)
    b3, a3 = (C.c_double * 3)(b0, b1, b2), (C.c_double * 3)(a0, a1, a2)
    L.check(lib.wm_biquad_fwd(L.ptr(x2), L.ptr(y), L.ptr(q), rows, T, C.addressof(b3), C.addressof(a3), int(clamp),
                              L.ptr(ws), n, _stream()), "wm_biquad_fwd")
    y = y.reshape(*lead, T)
    return (y, q.reshape(*lead, T)) if want_pcm16 else y


def lowpass_biquad_coeffs(sample_rate: int, cutoff_freq: float, Q: float = 0.707):
    """The six coefficients torchaudio.functional.lowpass_biquad builds, evaluated in fp32 like torchaudio does for an
    fp32 waveform (so the filter is the same filter, not a double-precision cousin of it)."""
    w0 = 2 * math.pi * torch.tensor(float(cutoff_freq), dtype=torch.float32) / sample_rate
    alpha = torch.sin(w0) / 2 / torch.tensor(float(Q), dtype=torch.float32)
    b0 = (1 - torch.cos(w0)) / 2
    b1 = 1 - torch.cos(w0)
    b2 = b0
    a0 = 1 + alpha
    a1 = -2 * torch.cos(w0)
    a2 = 1 - alpha
    return tuple(float(v) for v in (b0, b1, b2, a0, a1, a2))


def lowpass_biquad(waveform: torch.Tensor, sample_rate: int, cutoff_freq: float, Q: float = 0.707) -> torch.Tensor:
    """torchaudio.functional.lowpass_biquad(waveform, sample_rate, cutoff_freq, Q) on the GPU (py/main15.py:855)."""
    return biquad(waveform, *lowpass_biquad_coeffs(sample_rate, cutoff_freq, Q))


def perceptual_postprocess(waveform: torch.Tensor, sample_rate: int = 16000, cutoff_freq: float = 7000.0):
    """7 kHz low-pass + 16-bit quantisation of py/main15.py:850-860: returns (filtered fp32, int16 codes)."""
    return biquad(waveform, *lowpass_biquad_coeffs(sample_rate, cutoff_freq), want_pcm16=True)


def save_audio_pcm16(waveform: torch.Tensor, output_path: str, sample_rate: int = 16000) -> None:
    """py/main15.py:850-867 `save_audio`: low-pass at 7 kHz, quantise to signed 16-bit PCM, write a PCM_S WAV.  Filter
    and quantiser run on the GPU for CUDA input; only the int16 codes cross to the host."""
    import wave
    w = waveform if waveform.dim() == 2 else waveform.reshape(1, -1)
    if not w.is_cuda:
        w = w.cuda()
    _, q = perceptual_postprocess(w.float(), sample_rate)
    codes = q.t().contiguous().cpu().numpy()          # (N, C) interleaved
    with wave.open(output_path, "wb") as f:
        f.setnchannels(codes.shape[1])
        f.setsampwidth(2)
        f.setframerate(sample_rate)
        f.writeframes(codes.astype("<i2").tobytes())
