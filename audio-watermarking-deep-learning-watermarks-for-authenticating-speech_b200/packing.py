"""State dict -> packed fp32 weight blobs (layout: include/wmb200.h).

Host-side, device-independent: eval-mode BatchNorm (py/main16.py:117,120) is folded
into the preceding Conv1d in float64, conv weights are re-laid out tap-major
``w[j][ci][co]``, the ConvTranspose1d (py/main16.py:144) becomes an ordinary
convolution with flipped taps, and the two LSTM biases are summed.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import _lib as L

BN_EPS = 1e-5
PREFIX = "_orig_mod."


def strip_prefix(sd: Dict[str, torch.Tensor], prefix: str = PREFIX) -> Dict[str, torch.Tensor]:
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in sd.items()}


def _f64(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to("cpu", torch.float64)


def fold_conv_bn(sd, conv: str, bn: str):
    """Conv1d (co,ci,k) followed by eval BatchNorm1d -> (w[k][ci][co], b[co]) in float64."""
    w, b = _f64(sd[conv + ".weight"]), _f64(sd[conv + ".bias"])
    sc = _f64(sd[bn + ".weight"]) / torch.sqrt(_f64(sd[bn + ".running_var"]) + BN_EPS)
    w = w * sc[:, None, None]
    b = (b - _f64(sd[bn + ".running_mean"])) * sc + _f64(sd[bn + ".bias"])
    return w.permute(2, 1, 0).contiguous(), b


def _put(blob: torch.Tensor, off: int, t: torch.Tensor) -> None:
    flat = t.reshape(-1)
    blob[off:off + flat.numel()] = flat


def _pack_resblock(blob, off, sd, p):
    w1, b1 = fold_conv_bn(sd, p + ".block.0", p + ".block.1")
    w2, b2 = fold_conv_bn(sd, p + ".block.3", p + ".block.4")
    _put(blob, off + L.RB_W1, w1)
    _put(blob, off + L.RB_B1, b1)
    _put(blob, off + L.RB_W2, w2)
    _put(blob, off + L.RB_B2, b2)


def _pack_fused_input(blob, off, w_in, b_in, w1, b1):
    """Input Conv1d(1,64,7) composed with the first 64->64 k3 convolution behind it (no non-linearity between
    them, py/main16.py:134-135 / :177-178), in float64.  w_in [7][64] (tap, channel), b_in [64],
    w1 [3][ci][co] and b1 [co] with BatchNorm folded.  Layout: WM_FIN_* of include/wmb200.h."""
    wk = torch.einsum("kio,ji->kjo", w1, w_in)            # [3][7][co]
    bk = torch.einsum("kio,i->ko", w1, b_in)              # [3][co]
    w9 = torch.zeros(9, 64, dtype=torch.float64)
    for k in range(3):
        for j in range(7):
            w9[k + j] += wk[k, j]
    _put(blob, off + L.FIN_W9, w9)
    _put(blob, off + L.FIN_B9, b1 + bk.sum(0))
    _put(blob, off + L.FIN_WK, wk)
    _put(blob, off + L.FIN_BK, bk)


def pack_generator(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Everything of Generator (py/main16.py:128-162) except the embedding table."""
    sd = strip_prefix(sd)
    blob = torch.zeros(L.G_SIZE, dtype=torch.float64)
    _put(blob, L.G_IN_W, _f64(sd["encoder.0.weight"])[:, 0, :].t())          # (64,1,7) -> [7][64]
    _put(blob, L.G_IN_B, _f64(sd["encoder.0.bias"]))
    _pack_resblock(blob, L.G_RB0, sd, "encoder.1")
    _pack_resblock(blob, L.G_RB1, sd, "encoder.2")
    _put(blob, L.G_LSTM_WIH, _f64(sd["lstm.weight_ih_l0"]))
    _put(blob, L.G_LSTM_WHH, _f64(sd["lstm.weight_hh_l0"]))
    _put(blob, L.G_LSTM_B, _f64(sd["lstm.bias_ih_l0"]) + _f64(sd["lstm.bias_hh_l0"]))
    # ConvTranspose1d weight is (ci,co,k): y[t] = sum_k x[t+3-k] w[:, :, k]  ==  conv with taps flipped
    wt = _f64(sd["decoder.0.weight"]).flip(-1).permute(2, 0, 1)              # -> [j][ci][co]
    _put(blob, L.G_CT_W, wt)
    _put(blob, L.G_CT_B, _f64(sd["decoder.0.bias"]))
    _pack_resblock(blob, L.G_RB2, sd, "decoder.1")
    _put(blob, L.G_HEAD_W, _f64(sd["decoder.2.weight"])[0, :, 0])
    _put(blob, L.G_HEAD_B, _f64(sd["decoder.2.bias"]))
    w1, b1 = fold_conv_bn(sd, "encoder.1.block.0", "encoder.1.block.1")
    _pack_fused_input(blob, L.G_FIN, _f64(sd["encoder.0.weight"])[:, 0, :].t(), _f64(sd["encoder.0.bias"]), w1, b1)
    return blob.to(torch.float32)


def pack_detector(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Detector (py/main16.py:170-186); head rows beyond 1+message_bits stay zero."""
    sd = strip_prefix(sd)
    blob = torch.zeros(L.D_SIZE, dtype=torch.float64)
    _put(blob, L.D_IN_W, _f64(sd["model.0.weight"])[:, 0, :].t())
    _put(blob, L.D_IN_B, _f64(sd["model.0.bias"]))
    _pack_resblock(blob, L.D_RB0, sd, "model.1")
    _pack_resblock(blob, L.D_RB1, sd, "model.2")
    hw = _f64(sd["model.3.weight"])[:, :, 0]                                 # (nout,64)
    if hw.shape[0] > L.MAX_HEAD:
        raise ValueError(f"detector head has {hw.shape[0]} outputs; this build supports <= {L.MAX_HEAD}")
    _put(blob, L.D_HEAD_W, hw)
    _put(blob, L.D_HEAD_B, _f64(sd["model.3.bias"]))
    w1, b1 = fold_conv_bn(sd, "model.1.block.0", "model.1.block.1")
    _pack_fused_input(blob, L.D_FIN, _f64(sd["model.0.weight"])[:, 0, :].t(), _f64(sd["model.0.bias"]), w1, b1)
    return blob.to(torch.float32)


def fir_taps(cutoff: float = 4000.0, taps: int = 101, sample_rate: int = 16000) -> torch.Tensor:
    """The taps fir_lowpass builds (py/main16.py:53-62), fp32 like the reference."""
    import math
    fc = cutoff / (sample_rate / 2)
    n = torch.arange(taps) - (taps - 1) / 2
    sinc = torch.where(n == 0, torch.tensor(2 * fc), torch.sin(2 * math.pi * fc * n) / (math.pi * n))
    window = 0.54 - 0.46 * torch.cos(2 * math.pi * (n + (taps - 1) / 2) / (taps - 1))
    k = sinc * window
    return (k / k.sum()).to(torch.float32)


def mel_filterbank(n_freqs: int = 513, n_mels: int = 64, sample_rate: int = 16000):
    """The matrix torchaudio.transforms.MelSpectrogram(16000, 1024, 256, 64) multiplies the power spectrum by
    (py/main16.py:195-197): HTK mel scale, triangular filters, norm=None, f in [0, sr/2], fp32 throughout.
    Returns (fb [n_freqs][n_mels] fp32, band [n_mels][2] int32 = half-open non-zero bin range per filter)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_max = 2595.0 * math.log10(1.0 + (sample_rate / 2) / 700.0)
    m_pts = torch.linspace(0.0, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.min(down, up), min=0.0).contiguous()
    band = torch.zeros(n_mels, 2, dtype=torch.int32)
    for m in range(n_mels):
        nz = torch.nonzero(fb[:, m]).flatten()
        if nz.numel():
            band[m, 0], band[m, 1] = int(nz[0]), int(nz[-1]) + 1
    return fb, band


def resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """The windowed-sinc table torchaudio.transforms.Resample(orig, new) convolves with (sinc_interp_hann, its defaults:
    torchaudio.functional._get_sinc_resample_kernel), built in float64 and rounded to fp32 like torchaudio does.
    Returns (kern [K][up] fp32, down, up, width): y[m*up + j] = sum_k kern[k][j] x[m*down + k - width]."""
    g = math.gcd(int(orig_freq), int(new_freq))
    down, up = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(down, up) * rolloff
    width = math.ceil(lowpass_filter_width * down / base_freq)
    idx = torch.arange(-width, width + down, dtype=torch.float64)[None, None] / down
    t = torch.arange(0, -up, -1, dtype=torch.float64)[:, None, None] / up + idx
    t = (t * base_freq).clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base_freq / down
    kern = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t) * window * scale
    return kern[:, 0, :].to(torch.float32).t().contiguous(), down, up, width
