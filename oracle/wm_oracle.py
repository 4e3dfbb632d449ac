"""CPU oracle for the main16 embed+detect hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement, in torch-CPU fp32 functional form, of
the arithmetic the reference performs on the path named by BASELINE.json.  It
is NOT product code: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The shipped package (``wmb200``) never does; it fails loudly when the CUDA
library is missing.

Parity status: the reference has no golden vectors or tests, so
this oracle is pinned against outputs of the reference's *own definitions*
executed in the build container: ``tests/golden/make_golden.py`` AST-extracts
``ResBlock/Generator/Detector`` and the helper functions from
``/root/reference/py/main16.py`` and stores their outputs in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
here against those files.  The arithmetic itself lives in a third-party,
un-vendored dependency (PyTorch; README pins "1.9+", this container has
2.11.0); ``lstm_steps`` restates the one operator whose semantics are not obvious
from its call (gate order, zero state) step by step and is checked against the library call.

All functions take a *state dict* (name -> tensor, reference key names with or
without the ``_orig_mod.`` prefix) instead of nn.Modules.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

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
BN_EPS = 1e-5            # nn.BatchNorm1d default used at py/main16.py:117,120

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def strip_prefix(sd: SD, prefix: str = "_orig_mod.") -> SD:
    """py/main16.py:707-712 (key rewrite only)."""
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in sd.items()}


# --------------------------------------------------------------------------
# a1  ResBlock  (py/main16.py:112-125), eval-mode BatchNorm
# --------------------------------------------------------------------------
def bn_eval(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def resblock(x: Tensor, sd: SD, p: str) -> Tensor:
    y = F.conv1d(x, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], padding=1)
    y = F.relu(bn_eval(y, sd, p + ".block.1"))
    y = F.conv1d(y, sd[p + ".block.3.weight"], sd[p + ".block.3.bias"], padding=1)
    y = bn_eval(y, sd, p + ".block.4")
    return F.relu(x + y)


# --------------------------------------------------------------------------
# a3  LSTM(64,64), batch_first, zero initial state (py/main16.py:138,153)
# --------------------------------------------------------------------------
def lstm(x: Tensor, sd: SD, p: str = "lstm") -> Tensor:
    """x (B,T,64) -> all hidden states (B,T,64).  Gate row order i,f,g,o."""
    w = [sd[p + ".weight_ih_l0"], sd[p + ".weight_hh_l0"],
         sd[p + ".bias_ih_l0"], sd[p + ".bias_hh_l0"]]
    B, H = x.shape[0], w[1].shape[1]
    z = x.new_zeros(1, B, H)
    out, _, _ = torch._VF.lstm(x, (z, z), w, True, 1, 0.0, False, False, True)
    return out


def lstm_steps(x: Tensor, sd: SD, p: str = "lstm") -> Tensor:
    """Explicit per-step restatement of `lstm` (slow; used to pin the library call)."""
    w_ih, w_hh = sd[p + ".weight_ih_l0"], sd[p + ".weight_hh_l0"]
    b = sd[p + ".bias_ih_l0"] + sd[p + ".bias_hh_l0"]
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = x.new_empty(B, T, H)
    xp = x @ w_ih.t() + b
    for t in range(T):
        g = xp[:, t] + h @ w_hh.t()
        i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


# --------------------------------------------------------------------------
# a2  Generator.forward (py/main16.py:149-162)
# --------------------------------------------------------------------------
def generator_encoder(sd: SD, s: Tensor) -> Tensor:
    x = F.conv1d(s, sd["encoder.0.weight"], sd["encoder.0.bias"], padding=3)
    x = resblock(x, sd, "encoder.1")
    return resblock(x, sd, "encoder.2")


def generator_decoder(sd: SD, x: Tensor) -> Tensor:
    x = F.conv_transpose1d(x, sd["decoder.0.weight"], sd["decoder.0.bias"], padding=3)
    x = resblock(x, sd, "decoder.1")
    return F.conv1d(x, sd["decoder.2.weight"], sd["decoder.2.bias"])


def generator_forward(sd: SD, s: Tensor, message: Optional[Tensor] = None,
                      emb_rows: Optional[Tensor] = None) -> Tensor:
    """s (B,1,T) fp32, message (B,) int64 -> delta (B,1,T).

    `emb_rows` (B,64) may be given instead of a full `embedding.weight` table
    (the golden fixtures store only the rows they use)."""
    sd = strip_prefix(sd)
    x = generator_encoder(sd, s)
    x = lstm(x.permute(0, 2, 1), sd).permute(0, 2, 1)
    if message is not None and ("embedding.weight" in sd or emb_rows is not None):
        emb = emb_rows if emb_rows is not None else sd["embedding.weight"][message]
        x = x + emb.unsqueeze(-1)
    return generator_decoder(sd, x)


# --------------------------------------------------------------------------
# a4  Detector.forward (py/main16.py:183-186)
# --------------------------------------------------------------------------
def detector_forward(sd: SD, x: Tensor) -> Tensor:
    """x (B,1,T) -> logits (B,T,1+bits); channel 0 detection, 1.. bits LSB first."""
    sd = strip_prefix(sd)
    y = F.conv1d(x, sd["model.0.weight"], sd["model.0.bias"], padding=3)
    y = resblock(y, sd, "model.1")
    y = resblock(y, sd, "model.2")
    y = F.conv1d(y, sd["model.3.weight"], sd["model.3.bias"])
    return y.permute(0, 2, 1)


# --------------------------------------------------------------------------
# a5-a7  delta post-processing (py/main16.py:53-72)
# --------------------------------------------------------------------------
def fir_taps(cutoff: float = 4000, taps: int = 101) -> Tensor:
    """The 101 fp32 taps py/main16.py:58-62 builds (numerically ~identity)."""
    fc = cutoff / (SAMPLE_RATE / 2)
    n = torch.arange(taps) - (taps - 1) / 2
    sinc = torch.where(n == 0, 2 * fc, torch.sin(2 * math.pi * fc * n) / (math.pi * n))
    window = 0.54 - 0.46 * torch.cos(2 * math.pi * (n + (taps - 1) / 2) / (taps - 1))
    k = sinc * window
    return k / k.sum()


def fir_lowpass(delta: Tensor, cutoff: float = 4000, taps: int = 101) -> Tensor:
    # the reference builds the taps on delta.device (py/main16.py:58)
    return F.conv1d(delta, fir_taps(cutoff, taps).to(delta).view(1, 1, -1), padding=(taps - 1) // 2)


def clamp_peak(d: Tensor, thr: float = 0.02) -> Tensor:
    return d.clamp(-thr, thr)


def limit_rms(delta: Tensor, max_rms: float = MAX_RMS, eps: float = 1e-8) -> Tensor:
    cur = torch.sqrt((delta ** 2).mean(dim=[1, 2], keepdim=True) + eps)
    return delta * torch.clamp(max_rms / cur, max=1.0)


def postprocess(delta: Tensor) -> Tensor:
    """py/main16.py:245-247."""
    return limit_rms(clamp_peak(fir_lowpass(delta)))


# --------------------------------------------------------------------------
# a8-a10  training losses (py/main16.py:74-81, 192-217)
# --------------------------------------------------------------------------
def stft_mag(x: Tensor, n_fft: int, hop: int) -> Tensor:
    """|torch.stft| with periodic Hann, centre + reflect pad: (B,T)->(B,n_fft/2+1,1+T//hop)."""
    return torch.stft(x, n_fft, hop, window=torch.hann_window(n_fft).to(x),
                      return_complex=True).abs()


def high_freq_penalty(delta: Tensor, cutoff: float = 3500, n_fft: int = 512) -> Tensor:
    spec = stft_mag(delta.squeeze(1), n_fft, n_fft // 4)
    freqs = torch.fft.rfftfreq(n_fft, 1 / SAMPLE_RATE)
    return (spec * (freqs > cutoff).to(spec).view(1, -1, 1)).mean()


def mel_filterbank(n_freqs: int = 513, n_mels: int = 64, sr: int = 16000) -> Tensor:
    """HTK triangular filterbank, norm=None, f in [0, sr/2] (torchaudio melscale_fbanks)."""
    all_freqs = torch.linspace(0, sr // 2, n_freqs)
    m_max = 2595.0 * math.log10(1.0 + (sr / 2) / 700.0)
    m_pts = torch.linspace(0.0, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)           # (n_freqs, n_mels)


def mel_spectrogram(x: Tensor) -> Tensor:
    """MelSpectrogram(16k, n_fft 1024, hop 256, 64 mels, power 2): (B,1,T)->(B,1,64,63)."""
    p = stft_mag(x.reshape(-1, x.shape[-1]), 1024, 256) ** 2           # (B,513,F)
    m = torch.matmul(p.transpose(-1, -2), mel_filterbank().to(p)).transpose(-1, -2)
    return m.reshape(x.shape[:-1] + m.shape[-2:])


def mel_loss(clean: Tensor, wm: Tensor) -> Tensor:
    return F.l1_loss(torch.log(mel_spectrogram(clean) + 1e-5),
                     torch.log(mel_spectrogram(wm) + 1e-5))


def loudness_loss(clean: Tensor, wm: Tensor) -> Tensor:
    sc = stft_mag(clean.squeeze(1), 2048, 512)
    sw = stft_mag(wm.squeeze(1), 2048, 512)
    return (((sw - sc) ** 2) * (sc > 0.01).to(sw)).mean()


def bit_targets(message: Tensor, bits: int = MESSAGE_BITS) -> Tensor:
    """py/main16.py:261-262 — bit j of the id <-> logits channel 1+j (LSB first)."""
    return ((message.unsqueeze(1) & (1 << torch.arange(bits))) > 0).float()


# --------------------------------------------------------------------------
# a11  batched embed+detect step (forward of py/main16.py:238-276 / 378-403)
# --------------------------------------------------------------------------
def embed_detect(gsd: SD, dsd: SD, s: Tensor, message: Tensor,
                 emb_rows: Optional[Tensor] = None, detect_clean: bool = False) -> dict:
    """The benchmark unit: G -> fir/clamp/rms -> s+delta -> D -> heads.

    Returns per-sample probabilities, clip-mean probability, mean message
    logits and both bit-decoding rules (mean-logit sign, py/main16.py:1146,1185;
    majority vote, py/main16.py:398)."""
    delta_raw = generator_forward(gsd, s, message, emb_rows)
    delta = postprocess(delta_raw)
    s_w = s + delta
    x = torch.cat([s_w, s], 0) if detect_clean else s_w
    logits = detector_forward(dsd, x)
    probs = torch.sigmoid(logits[:, :, 0])
    mlog = logits[:, :, 1:].mean(dim=1)
    vote = (torch.sigmoid(logits[:, :, 1:]) > 0.5).float().mean(dim=1) > 0.5
    return {"delta_raw": delta_raw, "delta": delta, "s_w": s_w, "logits": logits,
            "probs": probs, "clip_prob": probs.mean(dim=1), "msg_logits": mlog,
            "bits_mean": (mlog > 0), "bits_vote": vote}


def losses(gsd: SD, dsd: SD, s: Tensor, message: Tensor,
           emb_rows: Optional[Tensor] = None) -> dict:
    """Loss scalars of py/main16.py:252-276 with eval-mode modules (validate_one_epoch)."""
    r = embed_detect(gsd, dsd, s, message, emb_rows, detect_clean=True)
    B, T = s.shape[0], s.shape[-1]
    det = r["logits"][:, :, 0]
    tgt = torch.cat([torch.ones(B, T), torch.zeros(B, T)], 0)
    loc = F.binary_cross_entropy_with_logits(det, tgt)
    tb = bit_targets(message).unsqueeze(1).expand(-1, T, -1)
    bce = F.binary_cross_entropy_with_logits(r["logits"][:B, :, 1:], tb)
    l1 = r["delta"].abs().mean()
    mel = mel_loss(s, r["s_w"])
    loud = loudness_loss(s, r["s_w"])
    hf = high_freq_penalty(r["delta"])
    total = (LAMBDA_L1 * l1 + LAMBDA_MSSPEC * mel + LAMBDA_LOUD * loud +
             LAMBDA_LOC * loc + LAMBDA_DEC * bce + HF_PENALTY_W * hf)
    return {"l1": l1, "mel": mel, "loud": loud, "loc": loc, "bce": bce, "hf": hf,
            "total": total}
