"""Tensor-level wrappers over the C ABI (one function per entry point of
include/wmb200.h).  torch is used for device memory and the current stream only;
all arithmetic happens inside libwmb200.so.  Nothing here falls back to PyTorch ops.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib as L

PEAK = 0.02        # clamp_peak default, py/main16.py:66
MAX_RMS = 0.005    # py/main16.py:29
RMS_EPS = 1e-8     # limit_rms default, py/main16.py:69


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name}: wmb200 runs on a B200 (sm_100a) only — got a {t.device.type} tensor; "
            "there is no CPU fallback for this path")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def set_math_mode(mode: int) -> int:
    return L.load().wm_set_math_mode(mode)


def get_math_mode() -> int:
    return L.load().wm_get_math_mode()


def launch_count() -> int:
    return int(L.load().wm_launch_count())


_NVTX = os.environ.get("WMB200_NVTX", "0") == "1"


def nvtx(name: str):
    """Decorator: an NVTX range around the call when WMB200_NVTX=1 (ncu --nvtx --nvtx-include "wmb200.<name>/" selects the
    kernels of one API call; SURVEY.md section 5's tracing row).  A no-op otherwise."""
    def wrap(fn):
        if not _NVTX:
            return fn
        import functools

        @functools.wraps(fn)
        def inner(*a, **k):
            torch.cuda.nvtx.range_push("wmb200." + name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        return inner
    return wrap


def max_chunk() -> int:
    """Clips per device pass (workspace ~12.4 MB per clip in the fp32 layout)."""
    return int(os.environ.get("WMB200_MAX_CLIPS", "4736"))


# ---- single operators ------------------------------------------------------
def conv_in_k7(s: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """s (B,T) -> (B,T,64) channels-last.  py/main16.py:134,177."""
    s = _req(s, "s")
    B, T = s.shape
    y = torch.empty(B, T, 64, device=s.device, dtype=torch.float32)
    L.check(L.load().wm_conv_in_k7_fwd(L.ptr(s), L.ptr(_req(w, "w")), L.ptr(_req(b, "b")), L.ptr(y),
                                       B, T, _stream()), "wm_conv_in_k7_fwd")
    return y


def conv64(x, w, bias, residual=None, chan_add=None, taps: int = 3, relu: bool = False):
    """x (B,T,64) channels-last, w [taps][64][64] -> (B,T,64).  py/main16.py:116-121,144."""
    x = _req(x, "x")
    B, T, Cc = x.shape
    assert Cc == 64
    y = torch.empty_like(x)
    res = _req(residual, "residual") if residual is not None else None
    ca = _req(chan_add, "chan_add") if chan_add is not None else None
    L.check(L.load().wm_conv64_fwd(L.ptr(x), L.ptr(_req(w, "w")), L.ptr(_req(bias, "bias")), L.ptr(res),
                                   L.ptr(ca), L.ptr(y), B, T, taps, int(relu), _stream()), "wm_conv64_fwd")
    return y


def to_planar(x, chan_add=None):
    """fp32 channels-last (B,T,64) -> planar bf16 hi/lo planes (opaque uint8 buffer)."""
    lib = L.load()
    x = _req(x, "x")
    B, T, _ = x.shape
    y = torch.empty(max(lib.wm_planar_bytes(B, T), 16), dtype=torch.uint8, device=x.device)
    ca = _req(chan_add, "chan_add") if chan_add is not None else None
    L.check(lib.wm_to_planar(L.ptr(x), L.ptr(ca), L.ptr(y), B, T, _stream()), "wm_to_planar")
    return y


def from_planar(p, B: int, T: int):
    y = torch.empty(B, T, 64, device=p.device, dtype=torch.float32)
    L.check(L.load().wm_from_planar(L.ptr(p), L.ptr(y), B, T, _stream()), "wm_from_planar")
    return y


def pack_conv64_tc(w):
    """fp32 w [taps][64][64] -> bf16 tcgen05 weight image."""
    lib = L.load()
    w = _req(w, "w")
    taps = w.shape[0]
    img = torch.empty(lib.wm_conv64_tc_weight_bytes(taps), dtype=torch.uint8, device=w.device)
    L.check(lib.wm_pack_conv64_tc(L.ptr(w), L.ptr(img), taps, _stream()), "wm_pack_conv64_tc")
    return img


def conv64_tc(xp, w_img, bias, B: int, T: int, taps: int, residual=None, relu=False, want_planar=True,
              want_fp32=False):
    """The tensor-core convolution on planar tensors -> (planar or None, fp32 (B,T,64) or None)."""
    lib = L.load()
    y = torch.empty(max(lib.wm_planar_bytes(B, T), 16), dtype=torch.uint8, device=xp.device) if want_planar else None
    y32 = torch.empty(B, T, 64, device=xp.device, dtype=torch.float32) if want_fp32 else None
    L.check(lib.wm_conv64_tc_fwd(L.ptr(xp), L.ptr(w_img), L.ptr(_req(bias, "bias")), L.ptr(residual), L.ptr(y),
                                 L.ptr(y32), B, T, taps, int(relu), _stream()), "wm_conv64_tc_fwd")
    return y, y32


def resblock_tc(xp, w1, b1, w2, b2, B: int, T: int, want_planar=True, want_fp32=False, host_bias=False):
    """Fused ResBlock on planar x; w1/w2 fp32 [3][64][64] tap-major (BN folded).  host_bias: biases read on the
    host and passed to the kernel by value (the variant the module-level drivers run)."""
    lib = L.load()
    dev = xp.device
    img = torch.cat([pack_conv64_tc(w1), pack_conv64_tc(w2)])
    y = torch.empty(max(lib.wm_planar_bytes(B, T), 16), dtype=torch.uint8, device=dev) if want_planar else None
    y32 = torch.empty(B, T, 64, device=dev, dtype=torch.float32) if want_fp32 else None
    if host_bias:
        hb1, hb2 = b1.detach().float().cpu().contiguous(), b2.detach().float().cpu().contiguous()
        L.check(lib.wm_resblock_tc_hostbias_fwd(L.ptr(xp), L.ptr(img), L.ptr(hb1), L.ptr(hb2), L.ptr(y), L.ptr(y32), B, T,
                                                _stream()), "wm_resblock_tc_hostbias_fwd")
    else:
        L.check(lib.wm_resblock_tc_fwd(L.ptr(xp), L.ptr(img), L.ptr(_req(b1, "b1")), L.ptr(_req(b2, "b2")), L.ptr(y),
                                       L.ptr(y32), B, T, _stream()), "wm_resblock_tc_fwd")
    return y, y32


def lstm_tc(xp, w_ih, w_hh, bias, B: int, T: int, chan_add=None):
    """Tensor-core LSTM on planar tensors -> planar h (+ chan_add on the output)."""
    lib = L.load()
    dev = xp.device
    wpk = torch.empty(4 * 256 * 64 * 2, dtype=torch.uint8, device=dev)
    bp = torch.empty(256, dtype=torch.float32, device=dev)
    L.check(lib.wm_pack_lstm_tc(L.ptr(_req(w_ih, "w_ih")), L.ptr(_req(w_hh, "w_hh")), L.ptr(_req(bias, "bias")),
                                L.ptr(wpk), L.ptr(bp), _stream()), "wm_pack_lstm_tc")
    y = torch.empty(max(lib.wm_planar_bytes(B, T), 16), dtype=torch.uint8, device=dev)
    ca = _req(chan_add, "chan_add") if chan_add is not None else None
    L.check(lib.wm_lstm_tc_fwd(L.ptr(xp), L.ptr(wpk), L.ptr(bp), L.ptr(ca), L.ptr(y), B, T, _stream()),
            "wm_lstm_tc_fwd")
    return y


def lstm(x, w_ih, w_hh, bias):
    """x (B,T,64) -> all hidden states (B,T,64).  py/main16.py:138,153."""
    x = _req(x, "x")
    B, T, _ = x.shape
    h = torch.empty_like(x)
    L.check(L.load().wm_lstm_fwd(L.ptr(x), L.ptr(_req(w_ih, "w_ih")), L.ptr(_req(w_hh, "w_hh")),
                                 L.ptr(_req(bias, "bias")), L.ptr(h), B, T, _stream()), "wm_lstm_fwd")
    return h


def head(x, w, b):
    """x (B,T,64), w (nout,64) -> (B,T,nout).  py/main16.py:146,180."""
    x = _req(x, "x")
    B, T, _ = x.shape
    nout = w.shape[0]
    y = torch.empty(B, T, nout, device=x.device, dtype=torch.float32)
    L.check(L.load().wm_head_fwd(L.ptr(x), L.ptr(_req(w, "w")), L.ptr(_req(b, "b")), L.ptr(y), B, T, nout,
                                 _stream()), "wm_head_fwd")
    return y


def postprocess(delta_raw, s=None, fir=None, mode: int = L.POST_ALL, want_delta=True, want_sw=True,
                want_rms=False, peak=PEAK, max_rms=MAX_RMS, eps=RMS_EPS):
    """delta_raw, s: (B,T) -> (delta, s_w, rms) (None where not requested).  py/main16.py:245-248."""
    d = _req(delta_raw, "delta_raw")
    B, T = d.shape
    sc = _req(s, "s") if s is not None else None
    want_sw = want_sw and sc is not None
    delta = torch.empty_like(d) if want_delta else None
    s_w = torch.empty_like(d) if want_sw else None
    rms = torch.empty(B, device=d.device, dtype=torch.float32) if want_rms else None
    f = _req(fir, "fir") if fir is not None else None
    L.check(L.load().wm_postprocess_fwd(L.ptr(d), L.ptr(sc), L.ptr(f), L.ptr(delta), L.ptr(s_w), L.ptr(rms),
                                        B, T, mode, peak, max_rms, eps, _stream()), "wm_postprocess_fwd")
    return delta, s_w, rms


def detect_heads(logits, valid_len=None, want_probs=True, want_votes=True):
    """logits (B,T,nout) contiguous -> dict(probs, clip_prob, msg_logits, vote_frac)."""
    lg = _req(logits, "logits")
    B, T, nout = lg.shape
    dev = lg.device
    probs = torch.empty(B, T, device=dev) if want_probs else None
    clip = torch.empty(B, device=dev)
    ml = torch.empty(B, max(nout - 1, 0), device=dev)
    vf = torch.empty(B, max(nout - 1, 0), device=dev) if want_votes else None
    vl = _req(valid_len, "valid_len", torch.int32) if valid_len is not None else None
    L.check(L.load().wm_detect_heads_fwd(L.ptr(lg), L.ptr(vl), L.ptr(probs), L.ptr(clip), L.ptr(ml), L.ptr(vf),
                                         B, T, nout, _stream()), "wm_detect_heads_fwd")
    return {"probs": probs, "clip_prob": clip, "msg_logits": ml, "vote_frac": vf}


# ---- module-level drivers ---------------------------------------------------
def generator_fwd(blob, embedding, message, s):
    """Generator.forward on s (B,T) -> delta_raw (B,T).  py/main16.py:149-162."""
    lib = L.load()
    s = _req(s, "s")
    B, T = s.shape
    out = torch.empty_like(s)
    msg = _req(message, "message", torch.int64) if message is not None else None
    emb = _req(embedding, "embedding") if embedding is not None else None
    step = max_chunk()
    for b0 in range(0, B, step):
        nb = min(step, B - b0)
        nbytes = lib.wm_generator_workspace_bytes(nb, T)
        ws = _ws(nbytes, s.device)
        L.check(lib.wm_generator_fwd(L.ptr(blob), L.ptr(emb), emb.shape[0] if emb is not None else 0,
                                     L.ptr(msg[b0:b0 + nb]) if msg is not None else None,
                                     L.ptr(s[b0:b0 + nb]), L.ptr(out[b0:b0 + nb]), L.ptr(ws), nbytes, nb, T,
                                     _stream()), "wm_generator_fwd")
    return out


def detector_fwd(blob, x, nout: int):
    """Detector.forward on x (B,T) -> logits (B,T,nout).  py/main16.py:183-186."""
    lib = L.load()
    x = _req(x, "x")
    B, T = x.shape
    out = torch.empty(B, T, nout, device=x.device, dtype=torch.float32)
    step = max_chunk()
    for b0 in range(0, B, step):
        nb = min(step, B - b0)
        nbytes = lib.wm_detector_workspace_bytes(nb, T)
        ws = _ws(nbytes, x.device)
        L.check(lib.wm_detector_fwd(L.ptr(blob), L.ptr(x[b0:b0 + nb]), L.ptr(out[b0:b0 + nb]), L.ptr(ws),
                                    nbytes, nb, T, nout, _stream()), "wm_detector_fwd")
    return out


def detect_fwd(blob, x, nout: int, valid_len=None, want_probs=True, want_votes=True):
    """Detector + heads, no logits tensor.  py/main16.py:1140-1146, :392-398."""
    lib = L.load()
    x = _req(x, "x")
    B, T = x.shape
    dev = x.device
    nbits = max(nout - 1, 0)
    probs = torch.empty(B, T, device=dev) if want_probs else None
    clip = torch.empty(B, device=dev)
    ml = torch.empty(B, nbits, device=dev)
    vf = torch.empty(B, nbits, device=dev) if want_votes else None
    vl = _req(valid_len, "valid_len", torch.int32) if valid_len is not None else None
    step = max_chunk()
    for b0 in range(0, B, step):
        nb = min(step, B - b0)
        nbytes = lib.wm_detector_workspace_bytes(nb, T)
        ws = _ws(nbytes, dev)
        sl = slice(b0, b0 + nb)
        L.check(lib.wm_detect_fwd(L.ptr(blob), L.ptr(x[sl]), L.ptr(vl[sl]) if vl is not None else None,
                                  L.ptr(probs[sl]) if probs is not None else None, L.ptr(clip[sl]),
                                  L.ptr(ml[sl]) if nbits else None, L.ptr(vf[sl]) if vf is not None and nbits else None,
                                  L.ptr(ws), nbytes, nb, T, nout, _stream()), "wm_detect_fwd")
    return {"probs": probs, "clip_prob": clip, "msg_logits": ml, "vote_frac": vf}


def embed_detect_fwd(g_blob, embedding, d_blob, fir, message, s, nout: int, post_mode: int = L.POST_ALL,
                     want_delta=True, want_probs=True, want_votes=False, want_rms=False):
    """The benchmark unit on device tensors: s (B,T), message (B,) -> dict."""
    lib = L.load()
    s = _req(s, "s")
    B, T = s.shape
    dev = s.device
    nbits = max(nout - 1, 0)
    msg = _req(message, "message", torch.int64) if message is not None else None
    emb = _req(embedding, "embedding") if embedding is not None else None
    f = _req(fir, "fir") if fir is not None else None
    delta = torch.empty_like(s) if want_delta else None
    s_w = torch.empty_like(s)
    rms = torch.empty(B, device=dev) if want_rms else None
    probs = torch.empty(B, T, device=dev) if want_probs else None
    clip = torch.empty(B, device=dev)
    ml = torch.empty(B, nbits, device=dev)
    vf = torch.empty(B, nbits, device=dev) if want_votes else None
    step = max_chunk()
    for b0 in range(0, B, step):
        nb = min(step, B - b0)
        sl = slice(b0, b0 + nb)
        nbytes = lib.wm_embed_detect_workspace_bytes(nb, T)
        ws = _ws(nbytes, dev)
        opt = lambda t: L.ptr(t[sl]) if t is not None else None
        L.check(lib.wm_embed_detect_fwd(L.ptr(g_blob), L.ptr(emb), emb.shape[0] if emb is not None else 0,
                                        L.ptr(d_blob), L.ptr(f), opt(msg), L.ptr(s[sl]), opt(delta), L.ptr(s_w[sl]),
                                        opt(rms), opt(probs), L.ptr(clip[sl]), opt(ml) if nbits else None,
                                        opt(vf) if nbits else None, L.ptr(ws), nbytes, nb, T, nout, post_mode,
                                        _stream()), "wm_embed_detect_fwd")
    return {"delta": delta, "s_w": s_w, "delta_rms": rms, "probs": probs, "clip_prob": clip,
            "msg_logits": ml, "vote_frac": vf}


class HostPipeline:
    """embed+detect with HOST (pinned) buffers through wm_embed_detect_host: H2D, the device
    pipeline in micro-batches of `chunk` clips and D2H all on the current stream."""

    def __init__(self, g_blob, embedding, d_blob, fir, nout: int, T: int = 16000, chunk: Optional[int] = None,
                 post_mode: int = L.POST_ALL, device=None):
        self.lib = L.load()
        self.g_blob, self.embedding, self.d_blob, self.fir = g_blob, embedding, d_blob, fir
        self.nout, self.T, self.post_mode = nout, T, post_mode
        self.chunk = int(chunk or max_chunk())
        self.device = device or g_blob.device
        self.nbytes = self.lib.wm_embed_detect_host_workspace_bytes(self.chunk, T, nout)
        self.ws = _ws(self.nbytes, self.device)

    def __call__(self, host_s, host_message, host_s_w, host_probs=None, host_clip_prob=None,
                 host_msg_logits=None):
        """host_s: (B, T), or a flat tensor of fewer than B * T samples (B taken from host_s_w): the missing tail of
        the last clip is zero on the device (wm_embed_detect_host_ragged)."""
        for name, t in (("host_s", host_s), ("host_s_w", host_s_w), ("host_probs", host_probs)):
            if t is not None and (t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32):
                raise ValueError(f"{name}: expected a contiguous fp32 host tensor")
        B, T = host_s_w.shape
        assert T == self.T
        n = host_s.numel()
        if n > B * T or n <= (B - 1) * T:
            raise ValueError(f"host_s holds {n} samples, expected more than {(B - 1) * T} and at most {B * T}")
        emb = self.embedding
        L.check(self.lib.wm_embed_detect_host_ragged(
            L.ptr(self.g_blob), L.ptr(emb), emb.shape[0] if emb is not None else 0, L.ptr(self.d_blob),
            L.ptr(self.fir), L.ptr(host_message), L.ptr(host_s), n, L.ptr(host_s_w), L.ptr(host_probs),
            L.ptr(host_clip_prob), L.ptr(host_msg_logits), L.ptr(self.ws), self.nbytes, B, T, self.nout,
            self.chunk, self.post_mode, _stream()), "wm_embed_detect_host_ragged")


# ---- training-loss forward (py/main16.py:74-81, 192-217, 255-266) ----------------------------
def _loss_ws(lib, B: int, T: int, dev):
    n = lib.wm_loss_workspace_bytes(B, T)
    return _ws(n, dev), n


def stft_mag(x: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """|torch.stft(x, n_fft, hop, window=hann_window(n_fft), return_complex=True)|: (B,T) -> (B, n_fft/2+1, frames)."""
    lib = L.load()
    x = _req(x, "x")
    B, T = x.shape
    out = torch.empty(B, n_fft // 2 + 1, lib.wm_stft_frames(T, hop), device=x.device)
    L.check(lib.wm_stft_mag_fwd(L.ptr(x), L.ptr(out), B, T, n_fft, hop, _stream()), "wm_stft_mag_fwd")
    return out


def hf_penalty(delta: torch.Tensor, n_fft: int, first_bin: int) -> torch.Tensor:
    lib = L.load()
    d = _req(delta, "delta")
    B, T = d.shape
    out = torch.empty(1, device=d.device)
    ws, n = _loss_ws(lib, B, T, d.device)
    L.check(lib.wm_hf_penalty_fwd(L.ptr(d), L.ptr(out), L.ptr(ws), n, B, T, n_fft, first_bin, _stream()),
            "wm_hf_penalty_fwd")
    return out[0]


def loudness(clean: torch.Tensor, wm: torch.Tensor, n_fft: int = 2048, hop: int = 512, thresh: float = 0.01):
    lib = L.load()
    c, w = _req(clean, "clean"), _req(wm, "watermarked")
    if c.shape != w.shape:
        raise ValueError(f"clean {tuple(c.shape)} and watermarked {tuple(w.shape)} differ in shape")
    B, T = c.shape
    out = torch.empty(1, device=c.device)
    ws, n = _loss_ws(lib, B, T, c.device)
    L.check(lib.wm_loud_fwd(L.ptr(c), L.ptr(w), L.ptr(out), L.ptr(ws), n, B, T, n_fft, hop, thresh, _stream()),
            "wm_loud_fwd")
    return out[0]


def mel_log_l1(clean: torch.Tensor, wm: torch.Tensor, fb: torch.Tensor, band: torch.Tensor, n_fft: int = 1024,
               hop: int = 256):
    lib = L.load()
    c, w = _req(clean, "clean"), _req(wm, "watermarked")
    if c.shape != w.shape:
        raise ValueError(f"clean {tuple(c.shape)} and watermarked {tuple(w.shape)} differ in shape")
    fb = _req(fb, "fb")
    band = _req(band, "band", torch.int32)
    B, T = c.shape
    out = torch.empty(1, device=c.device)
    ws, n = _loss_ws(lib, B, T, c.device)
    L.check(lib.wm_mel_log_l1_fwd(L.ptr(c), L.ptr(w), L.ptr(fb), L.ptr(band), fb.shape[1], L.ptr(out), L.ptr(ws), n,
                                  B, T, n_fft, hop, _stream()), "wm_mel_log_l1_fwd")
    return out[0]


def bce_heads(logits: torch.Tensor, message: Optional[torch.Tensor], n_watermarked: int):
    """logits (B2,T,nout) -> (loc, bce): py/main16.py:255-264."""
    lib = L.load()
    lg = _req(logits, "logits")
    B2, T, nout = lg.shape
    msg = _req(message, "message", torch.int64) if message is not None else None
    out = torch.empty(2, device=lg.device)
    ws, n = _loss_ws(lib, B2, T, lg.device)
    want_bits = nout > 1 and n_watermarked > 0
    L.check(lib.wm_bce_heads_fwd(L.ptr(lg), L.ptr(msg), L.ptr(out[0:1]), L.ptr(out[1:2]) if want_bits else None,
                                 L.ptr(ws), n, n_watermarked, B2, T, nout, _stream()), "wm_bce_heads_fwd")
    return out[0], (out[1] if want_bits else None)


def abs_mean(x: torch.Tensor) -> torch.Tensor:
    lib = L.load()
    x = _req(x, "x")
    x2 = x.reshape(x.shape[0], -1)
    B, T = x2.shape
    out = torch.empty(1, device=x.device)
    ws, n = _loss_ws(lib, B, T, x.device)
    L.check(lib.wm_abs_mean_fwd(L.ptr(x2), L.ptr(out), L.ptr(ws), n, B, T, _stream()), "wm_abs_mean_fwd")
    return out[0]
