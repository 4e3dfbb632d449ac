"""Training-mode pieces of the main16 path (py/main16.py:223-294, BASELINE config 4).

Built so far: the Detector's half of `train_one_epoch` — forward with batch-statistics BatchNorm, the detection and
message BCE losses, backward through the whole network and Adam, all inside one C-ABI call
(`wm_detector_train_step`) on one flat parameter buffer — plus the operators it is made of (`bn_train_fwd/bwd`,
`conv64_bwd`, `adam_step`).  The Generator's backward (the LSTM through 16 000 steps and the spectral losses) is not
built: `DetectorTrainer.step` returns the gradient w.r.t. the detector input, which is where it would attach.
Everything runs in fp32 on the CUDA cores and is deterministic (fixed-order partial sums).
"""
from __future__ import annotations

from typing import Dict, Optional

import ctypes as C

import torch

from . import _lib as L
from .ops import _req, _stream, _ws
from .ops import nvtx as ops_nvtx

LR = 1e-3            # py/main16.py:33
LAMBDA_L1 = 1.0      # py/main16.py:38
LAMBDA_MSSPEC = 4.0  # py/main16.py:39
LAMBDA_LOUD = 20.0   # py/main16.py:40
LAMBDA_LOC = 10.0    # py/main16.py:41
LAMBDA_DEC = 1.0     # py/main16.py:42
HF_PENALTY_W = 5.0   # py/main16.py:43


# ---- operators ------------------------------------------------------------------------------------------------
def bn_train_fwd(z: torch.Tensor, gamma, beta, residual: Optional[torch.Tensor] = None, relu: bool = True,
                 running_mean: Optional[torch.Tensor] = None, running_var: Optional[torch.Tensor] = None):
    """relu?(BatchNorm1d(64)(z) + residual) with batch statistics on channels-last z (..., 64); running stats
    (momentum 0.1, unbiased variance) are updated in place when given.  Returns (out, mean, rstd)."""
    lib = L.load()
    z = _req(z, "z")
    rows = z.numel() // 64
    out = torch.empty_like(z)
    mean, rstd = torch.empty(64, device=z.device), torch.empty(64, device=z.device)
    res = _req(residual, "residual") if residual is not None else None
    n = lib.wm_bn_train_workspace_bytes(rows)
    ws = _ws(n, z.device)
    L.check(lib.wm_bn_train_fwd(L.ptr(z), L.ptr(_req(gamma, "gamma")), L.ptr(_req(beta, "beta")), L.ptr(res), L.ptr(out),
                                L.ptr(mean), L.ptr(rstd), L.ptr(running_mean), L.ptr(running_var), rows, int(relu),
                                L.ptr(ws), n, _stream()), "wm_bn_train_fwd")
    return out, mean, rstd


def bn_train_bwd(dout: torch.Tensor, act: Optional[torch.Tensor], z: torch.Tensor, mean, rstd, gamma,
                 want_residual_grad: bool = False):
    """Backward of bn_train_fwd: returns (dz, dres or None, dgamma, dbeta)."""
    lib = L.load()
    dout, z = _req(dout, "dout"), _req(z, "z")
    rows = z.numel() // 64
    dz = torch.empty_like(z)
    dres = torch.empty_like(z) if want_residual_grad else None
    dg, db = torch.empty(64, device=z.device), torch.empty(64, device=z.device)
    n = lib.wm_bn_train_workspace_bytes(rows)
    ws = _ws(n, z.device)
    L.check(lib.wm_bn_train_bwd(L.ptr(dout), L.ptr(_req(act, "act")) if act is not None else None, L.ptr(z),
                                L.ptr(_req(mean, "mean")), L.ptr(_req(rstd, "rstd")), L.ptr(_req(gamma, "gamma")),
                                L.ptr(dz), L.ptr(dres), L.ptr(dg), L.ptr(db), rows, L.ptr(ws), n, _stream()),
            "wm_bn_train_bwd")
    return dz, dres, dg, db


def conv64_bwd(x: torch.Tensor, dy: torch.Tensor, weight: torch.Tensor, want_dx: bool = True):
    """Gradients of Conv1d(64,64,K,padding=K//2) on channels-last x, dy (B,T,64); `weight` in the reference's
    (co,ci,K) layout.  Returns (dweight (co,ci,K), dbias (64), dx or None)."""
    lib = L.load()
    x, dy = _req(x, "x"), _req(dy, "dy")
    B, T, _ = x.shape
    K = weight.shape[-1]
    w_t = _req(weight, "weight").permute(2, 1, 0).contiguous()          # [k][ci][co]
    dw = torch.empty(K, 64, 64, device=x.device)
    db = torch.empty(64, device=x.device)
    dx = torch.empty_like(x) if want_dx else None
    n = lib.wm_conv64_bwd_workspace_bytes(B, T, K)
    ws = _ws(n, x.device)
    L.check(lib.wm_conv64_bwd(L.ptr(x), L.ptr(dy), L.ptr(w_t), L.ptr(dw), L.ptr(db), L.ptr(dx), B, T, K, L.ptr(ws), n,
                              _stream()), "wm_conv64_bwd")
    return dw.permute(2, 1, 0).contiguous(), db, dx


def conv64_train_fwd(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                     residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Forward of Conv1d(64,64,K,padding=K//2) as the training step runs it (tensor cores in the default math mode,
    fp32 FMA under WM_MATH_FP32): channels-last x (B,T,64), `weight` in the reference's (co,ci,K) layout."""
    lib = L.load()
    x = _req(x, "x")
    B, T, _ = x.shape
    K = weight.shape[-1]
    w_t = _req(weight, "weight").permute(2, 1, 0).contiguous()          # [k][ci][co]
    y = torch.empty_like(x)
    res = _req(residual, "residual") if residual is not None else None
    n = lib.wm_conv64_bwd_workspace_bytes(B, T, K)
    ws = _ws(n, x.device)
    L.check(lib.wm_conv64_train_fwd(L.ptr(x), L.ptr(w_t), L.ptr(_req(bias, "bias")), L.ptr(res), L.ptr(y), B, T, K,
                                    L.ptr(ws), n, _stream()), "wm_conv64_train_fwd")
    return y


def _gate_t(w: torch.Tensor) -> torch.Tensor:
    """(256,64) PyTorch LSTM weight -> per-gate transposed wT[q][k][r] = W[q*64 + r][k]; its own inverse."""
    return w.reshape(4, 64, 64).permute(0, 2, 1).contiguous()


def lstm_train_fwd(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh):
    """nn.LSTM(64,64,batch_first)(x)[0] on x (B,T,64) with PyTorch-layout weights; returns (h, saved) where `saved`
    feeds lstm_train_bwd."""
    lib = L.load()
    x = _req(x, "x")
    B, T, _ = x.shape
    wti, wth = _gate_t(_req(w_ih, "w_ih")), _gate_t(_req(w_hh, "w_hh"))
    h = torch.empty_like(x)
    gates = torch.empty(B, T, 256, device=x.device)
    cell = torch.empty(B, T, 64, device=x.device)
    L.check(lib.wm_lstm_train_fwd(L.ptr(x), L.ptr(wti), L.ptr(wth), L.ptr(_req(b_ih, "b_ih")), L.ptr(_req(b_hh, "b_hh")),
                                  L.ptr(h), L.ptr(gates), L.ptr(cell), B, T, _stream()), "wm_lstm_train_fwd")
    return h, (x, h, wti, wth, gates, cell)


def lstm_train_bwd(dy: torch.Tensor, saved):
    """-> (dx, dw_ih (256,64), dw_hh (256,64), db (256,)); db is the gradient of b_ih and of b_hh."""
    lib = L.load()
    x, h, wti, wth, gates, cell = saved
    dy = _req(dy, "dy")
    B, T, _ = x.shape
    dx = torch.empty_like(x)
    dwi, dwh = torch.empty(4, 64, 64, device=x.device), torch.empty(4, 64, 64, device=x.device)
    db = torch.empty(256, device=x.device)
    n = lib.wm_lstm_train_bwd_workspace_bytes(B, T)
    ws = _ws(n, x.device)
    L.check(lib.wm_lstm_train_bwd(L.ptr(dy), L.ptr(x), L.ptr(h), L.ptr(wti), L.ptr(wth), L.ptr(gates), L.ptr(cell),
                                  L.ptr(dx), L.ptr(dwi), L.ptr(dwh), L.ptr(db), B, T, L.ptr(ws), n, _stream()),
            "wm_lstm_train_bwd")
    return dx, _gate_t(dwi).reshape(256, 64), _gate_t(dwh).reshape(256, 64), db


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float = LR,
              betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    """torch.optim.Adam's update of the flat fp32 buffer p in place (py/main16.py:504)."""
    lib = L.load()
    for t, name in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError(f"adam_step: {name} must be a contiguous fp32 CUDA tensor")
    L.check(lib.wm_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), lr, betas[0], betas[1], eps, step,
                             _stream()), "wm_adam_step")


# ---- backward of the losses and of the post-processing ----------------------------------------------------------
def _stft_bwd_ws(lib, B, T, n_fft, hop, dev):
    n = lib.wm_stft_bwd_workspace_bytes(B, T, n_fft, hop)
    return _ws(n, dev), n


def _grad_target(like: torch.Tensor, into: Optional[torch.Tensor]):
    if into is None:
        return torch.empty_like(like), 0
    if not (into.is_cuda and into.dtype == torch.float32 and into.is_contiguous() and into.shape == like.shape):
        raise ValueError("the gradient to accumulate into must be a contiguous fp32 CUDA tensor of the input's shape")
    return into, 1


def hf_penalty_bwd(delta: torch.Tensor, n_fft: int = 512, first_bin: int = 113, weight: float = 1.0,
                   into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """weight * d high_freq_penalty(delta) / d delta for delta (B,T); added to `into` when given."""
    lib = L.load()
    d = _req(delta, "delta")
    B, T = d.shape
    g, acc = _grad_target(d, into)
    ws, n = _stft_bwd_ws(lib, B, T, n_fft, n_fft // 4, d.device)
    L.check(lib.wm_hf_penalty_bwd(L.ptr(d), L.ptr(g), L.ptr(ws), n, B, T, n_fft, first_bin, weight, acc, _stream()),
            "wm_hf_penalty_bwd")
    return g


def loudness_bwd(clean, wm, n_fft: int = 2048, hop: int = 512, thresh: float = 0.01, weight: float = 1.0,
                 into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """weight * d TFLoudnessLoss(clean, wm) / d wm."""
    lib = L.load()
    c, w = _req(clean, "clean"), _req(wm, "watermarked")
    B, T = c.shape
    g, acc = _grad_target(w, into)
    ws, n = _stft_bwd_ws(lib, B, T, n_fft, hop, c.device)
    L.check(lib.wm_loud_bwd(L.ptr(c), L.ptr(w), L.ptr(g), L.ptr(ws), n, B, T, n_fft, hop, thresh, weight, acc,
                            _stream()), "wm_loud_bwd")
    return g


def mel_log_l1_bwd(clean, wm, fb, band, n_fft: int = 1024, hop: int = 256, weight: float = 1.0,
                   into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """weight * d MultiScaleMelLoss(clean, wm) / d wm."""
    lib = L.load()
    c, w = _req(clean, "clean"), _req(wm, "watermarked")
    fb, band = _req(fb, "fb"), _req(band, "band", torch.int32)
    B, T = c.shape
    g, acc = _grad_target(w, into)
    ws, n = _stft_bwd_ws(lib, B, T, n_fft, hop, c.device)
    L.check(lib.wm_mel_log_l1_bwd(L.ptr(c), L.ptr(w), L.ptr(fb), L.ptr(band), fb.shape[1], L.ptr(g), L.ptr(ws), n, B, T,
                                  n_fft, hop, weight, acc, _stream()), "wm_mel_log_l1_bwd")
    return g


def abs_mean_bwd(x: torch.Tensor, weight: float = 1.0, into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """weight * d mean|x| / dx (F.l1_loss(delta, 0), py/main16.py:266)."""
    x = _req(x, "x")
    B, T = x.shape
    g, acc = _grad_target(x, into)
    L.check(L.load().wm_abs_mean_bwd(L.ptr(x), L.ptr(g), B, T, weight, acc, _stream()), "wm_abs_mean_bwd")
    return g


def postprocess_bwd(g: torch.Tensor, delta_raw: torch.Tensor, fir: Optional[torch.Tensor], mode: int = L.POST_ALL,
                    peak: float = 0.02, max_rms: float = 0.005, eps: float = 1e-8) -> torch.Tensor:
    """Gradient w.r.t. delta_raw of limit_rms(clamp_peak(fir_lowpass(delta_raw))) given g = dL/d delta."""
    from . import ops
    lib = L.load()
    g, dr = _req(g, "g"), _req(delta_raw, "delta_raw")
    B, T = dr.shape
    f = _req(fir, "fir") if fir is not None else None
    d1 = ops.postprocess(dr, None, f, mode & L.POST_FIR, want_sw=False)[0] if mode & L.POST_FIR else dr
    out = torch.empty_like(dr)
    n = B * T * 4
    ws = _ws(n, dr.device)
    L.check(lib.wm_postprocess_bwd(L.ptr(g), L.ptr(d1), L.ptr(f), L.ptr(out), L.ptr(ws), n, B, T, mode, peak, max_rms,
                                   eps, _stream()), "wm_postprocess_bwd")
    return out


def bce_heads_bwd(logits: torch.Tensor, message: Optional[torch.Tensor], n_watermarked: int, lam_loc: float = LAMBDA_LOC,
                  lam_dec: float = LAMBDA_DEC) -> torch.Tensor:
    """d(lam_loc * loc + lam_dec * bce) / d logits for logits (B_total, T, nout) (py/main16.py:255-264)."""
    lg = _req(logits, "logits")
    B2, T, nout = lg.shape
    msg = _req(message, "message", torch.int64) if message is not None else None
    out = torch.empty_like(lg)
    L.check(L.load().wm_bce_heads_bwd(L.ptr(lg), L.ptr(msg), L.ptr(out), n_watermarked, B2, T, nout, lam_loc, lam_dec,
                                      _stream()), "wm_bce_heads_bwd")
    return out


def head_bwd(dlogits: torch.Tensor, y: torch.Tensor, weight: torch.Tensor):
    """Backward of Conv1d(64,nout,1) on channels-last y (..., 64): -> (dy, dweight (nout,64,1), dbias (nout,))."""
    lib = L.load()
    dl, y = _req(dlogits, "dlogits"), _req(y, "y")
    nout = dl.shape[-1]
    rows = y.numel() // 64
    w = _req(weight, "weight").reshape(nout, 64).contiguous()
    dy, dw, db = torch.empty_like(y), torch.empty(nout, 64, device=y.device), torch.empty(nout, device=y.device)
    n = lib.wm_head_bwd_workspace_bytes(rows, nout)
    ws = _ws(n, y.device)
    L.check(lib.wm_head_bwd(L.ptr(dl), L.ptr(y), L.ptr(w), L.ptr(dy), L.ptr(dw), L.ptr(db), rows, nout, L.ptr(ws), n,
                            _stream()), "wm_head_bwd")
    return dy, dw.reshape(nout, 64, 1), db


def conv_in_k7_bwd(s: torch.Tensor, dx: torch.Tensor, weight: torch.Tensor, want_ds: bool = True):
    """Backward of Conv1d(1,64,7,padding=3): s (B,T), dx (B,T,64) channels-last, weight (64,1,7) ->
    (dweight (64,1,7), dbias (64,), ds (B,T) or None)."""
    lib = L.load()
    s, dx = _req(s, "s"), _req(dx, "dx")
    B, T = s.shape
    w = _req(weight, "weight").permute(2, 1, 0).reshape(7, 64).contiguous()
    dw, db = torch.empty(7, 64, device=s.device), torch.empty(64, device=s.device)
    ds = torch.empty_like(s) if want_ds else None
    n = lib.wm_conv_in_k7_bwd_workspace_bytes(B, T)
    ws = _ws(n, s.device)
    L.check(lib.wm_conv_in_k7_bwd(L.ptr(s), L.ptr(dx), L.ptr(w), L.ptr(dw), L.ptr(db), L.ptr(ds), B, T, L.ptr(ws), n,
                                  _stream()), "wm_conv_in_k7_bwd")
    return dw.reshape(7, 1, 64).permute(2, 1, 0).contiguous(), db, ds


# ---- flat parameter buffer <-> state dict -----------------------------------------------------------------------
def _dt_slices(nout: int):
    """name -> (offset, shape in the flat buffer, whether the state-dict layout is its (2,1,0) transpose)."""
    out = {"model.0.weight": (L.DT_IN_W, (7, 1, 64), True), "model.0.bias": (L.DT_IN_B, (64,), False)}
    for k in range(2):
        base, pre = L.DT_RB0 + k * L.DT_RB_SIZE, f"model.{1 + k}.block."
        out[pre + "0.weight"] = (base + L.DT_RB_W1, (3, 64, 64), True)
        out[pre + "0.bias"] = (base + L.DT_RB_B1, (64,), False)
        out[pre + "1.weight"] = (base + L.DT_RB_G1, (64,), False)
        out[pre + "1.bias"] = (base + L.DT_RB_BE1, (64,), False)
        out[pre + "3.weight"] = (base + L.DT_RB_W2, (3, 64, 64), True)
        out[pre + "3.bias"] = (base + L.DT_RB_B2, (64,), False)
        out[pre + "4.weight"] = (base + L.DT_RB_G2, (64,), False)
        out[pre + "4.bias"] = (base + L.DT_RB_BE2, (64,), False)
    out["model.3.weight"] = (L.DT_HEAD_W, (nout, 64, 1), False)
    out["model.3.bias"] = (L.DT_HEAD_B, (nout,), False)
    return out


def _dt_stat_slices():
    out = {}
    for k in range(2):
        for i, bn in enumerate(("1", "4")):
            pre = f"model.{1 + k}.block.{bn}."
            out[pre + "running_mean"] = k * 256 + i * 128
            out[pre + "running_var"] = k * 256 + i * 128 + 64
    return out


def flatten_detector(sd: Dict[str, torch.Tensor], nout: int, device) -> torch.Tensor:
    flat = torch.zeros(L.DT_SIZE, dtype=torch.float32, device=device)
    for name, (off, shape, perm) in _dt_slices(nout).items():
        t = sd[name].detach().to(device=device, dtype=torch.float32)
        if perm:
            t = t.permute(2, 1, 0)
        flat[off:off + t.numel()] = t.reshape(-1)
    return flat


def unflatten_detector(flat: torch.Tensor, nout: int) -> Dict[str, torch.Tensor]:
    out = {}
    for name, (off, shape, perm) in _dt_slices(nout).items():
        n = 1
        for d in shape:
            n *= d
        t = flat[off:off + n].reshape(shape)
        if perm:
            t = t.permute(2, 1, 0)
        out[name] = t.contiguous().clone()
    return out


class DetectorTrainer:
    """Holds a Detector's training state on the device (parameters, gradients, Adam moments, BatchNorm running stats)
    and advances it one `train_one_epoch` iteration at a time (py/main16.py:249-264,277-278, detector part).

        tr = DetectorTrainer(detector)          # wmb200.Detector or the reference's Detector (same state dict)
        for s_w, s, message in batches:
            out = tr.step(torch.cat([s_w, s]), message)     # {"loc": ..., "bce": ...}
        tr.write_back(detector)                 # parameters + running stats into the module
    """

    def __init__(self, detector: torch.nn.Module, lr: float = LR, betas=(0.9, 0.999), eps: float = 1e-8,
                 lambda_loc: float = LAMBDA_LOC, lambda_dec: float = LAMBDA_DEC, device=None):
        sd = detector.state_dict()
        self.nout = int(sd["model.3.weight"].shape[0])
        if self.nout > L.MAX_HEAD:
            raise ValueError(f"detector head has {self.nout} outputs; at most {L.MAX_HEAD} are supported")
        if device is None:
            device = sd["model.3.weight"].device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("DetectorTrainer runs on a B200 (sm_100a) only; there is no CPU fallback — move the "
                               "detector to cuda first")
        L.load()
        self.device = device
        self.params = flatten_detector(sd, self.nout, device)
        self.grads = torch.zeros_like(self.params)
        self.adam_m = torch.zeros_like(self.params)
        self.adam_v = torch.zeros_like(self.params)
        self.run_stats = torch.zeros(L.DT_STATS, dtype=torch.float32, device=device)
        for name, off in _dt_stat_slices().items():
            self.run_stats[off:off + 64] = sd[name].to(device=device, dtype=torch.float32)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.lambda_loc, self.lambda_dec = float(lambda_loc), float(lambda_dec)
        self.steps = 0
        self._steps0 = 0
        self._nbt0 = {k: int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")}
        self._ws = None
        self._ws_key = None

    def _workspace(self, B2: int, T: int):
        key = (B2, T)
        if self._ws_key != key:
            n = L.load().wm_detector_train_workspace_bytes(B2, T, self.nout)
            self._ws, self._ws_key = (_ws(n, self.device), n), key
        return self._ws

    def step(self, x: torch.Tensor, message: Optional[torch.Tensor], n_watermarked: Optional[int] = None,
             update: bool = True, want_input_grad: bool = False) -> Dict[str, torch.Tensor]:
        """x (B_total, T) or (B_total, 1, T): the first n_watermarked clips (default: len(message)) carry message[b].
        Returns {"loc", "bce"} (0-d device tensors, the unweighted losses) and "d_input" when asked.
        update=False computes the gradients (self.grads) and BatchNorm statistics without touching the parameters."""
        lib = L.load()
        x = _req(x, "x")
        if x.dim() == 3 and x.shape[1] == 1:
            x = x[:, 0]
        if x.dim() != 2:
            raise ValueError(f"x must be (B,T) or (B,1,T), got {tuple(x.shape)}")
        B2, T = x.shape
        x = x.contiguous()
        msg = _req(message, "message", torch.int64) if message is not None else None
        if n_watermarked is None:
            n_watermarked = 0 if msg is None else int(msg.numel())
        if msg is not None and msg.numel() < n_watermarked:
            raise ValueError("message must hold one value per watermarked clip")
        losses = torch.zeros(2, device=self.device)
        d_in = torch.empty_like(x) if want_input_grad else None
        ws, n = self._workspace(B2, T)
        step_no = self.steps + 1 if update else 0
        L.check(lib.wm_detector_train_step(L.ptr(self.params), L.ptr(self.grads), L.ptr(self.adam_m), L.ptr(self.adam_v),
                                           L.ptr(self.run_stats), L.ptr(x), L.ptr(msg), n_watermarked, B2, T, self.nout,
                                           self.lambda_loc, self.lambda_dec, self.lr, self.betas[0], self.betas[1],
                                           self.eps, step_no, L.ptr(losses), L.ptr(d_in), L.ptr(ws), n, _stream()),
                "wm_detector_train_step")
        if update:
            self.steps += 1
        out = {"loc": losses[0], "bce": losses[1]}
        if want_input_grad:
            out["d_input"] = d_in
        return out

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """The detector's state dict (reference key names and layouts) at the current step."""
        sd = unflatten_detector(self.params, self.nout)
        for name, off in _dt_stat_slices().items():
            sd[name] = self.run_stats[off:off + 64].clone()
        return sd

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        """Gradients of the last step, keyed and laid out like the parameters of the state dict."""
        return unflatten_detector(self.grads, self.nout)

    @torch.no_grad()
    def write_back(self, detector: torch.nn.Module) -> None:
        own = detector.state_dict()
        for k, v in self.state_dict().items():
            own[k].copy_(v.to(own[k].device))
        for k, n0 in self._nbt0.items():
            own[k].fill_(n0 + self.steps - self._steps0)


# ---- the whole training step (py/main16.py:238-278) ---------------------------------------------------------------
def _rb_slices(out, base, pre):
    out[pre + "0.weight"] = (base + L.DT_RB_W1, (3, 64, 64), "t210")
    out[pre + "0.bias"] = (base + L.DT_RB_B1, (64,), None)
    out[pre + "1.weight"] = (base + L.DT_RB_G1, (64,), None)
    out[pre + "1.bias"] = (base + L.DT_RB_BE1, (64,), None)
    out[pre + "3.weight"] = (base + L.DT_RB_W2, (3, 64, 64), "t210")
    out[pre + "3.bias"] = (base + L.DT_RB_B2, (64,), None)
    out[pre + "4.weight"] = (base + L.DT_RB_G2, (64,), None)
    out[pre + "4.bias"] = (base + L.DT_RB_BE2, (64,), None)


def _gt_slices():
    """name -> (offset, flat shape, layout transform); transforms are involutions or come with an inverse below."""
    out = {"encoder.0.weight": (L.GT_IN_W, (7, 1, 64), "t210"), "encoder.0.bias": (L.GT_IN_B, (64,), None)}
    _rb_slices(out, L.GT_RB0, "encoder.1.block.")
    _rb_slices(out, L.GT_RB1, "encoder.2.block.")
    out["lstm.weight_ih_l0"] = (L.GT_LSTM_WIH, (4, 64, 64), "gate")
    out["lstm.weight_hh_l0"] = (L.GT_LSTM_WHH, (4, 64, 64), "gate")
    out["lstm.bias_ih_l0"] = (L.GT_LSTM_BIH, (256,), None)
    out["lstm.bias_hh_l0"] = (L.GT_LSTM_BHH, (256,), None)
    out["decoder.0.weight"] = (L.GT_CT_W, (7, 64, 64), "convT")
    out["decoder.0.bias"] = (L.GT_CT_B, (64,), None)
    _rb_slices(out, L.GT_RB2, "decoder.1.block.")
    out["decoder.2.weight"] = (L.GT_HEAD_W, (1, 64, 1), None)
    out["decoder.2.bias"] = (L.GT_HEAD_B, (1,), None)
    out["embedding.weight"] = (L.GT_EMB, (65536, 64), None)
    return out


def _gt_stat_slices():
    out = {}
    for k, mod in enumerate(("encoder.1", "encoder.2", "decoder.1")):
        for i, bn in enumerate(("1", "4")):
            out[f"{mod}.block.{bn}.running_mean"] = k * 256 + i * 128
            out[f"{mod}.block.{bn}.running_var"] = k * 256 + i * 128 + 64
    return out


def _to_flat(t: torch.Tensor, how) -> torch.Tensor:
    if how == "t210":
        return t.permute(2, 1, 0)
    if how == "gate":
        return _gate_t(t)
    if how == "convT":            # ConvTranspose1d (ci,co,k) -> convolution taps [j][ci][co] = w[ci][co][6-j]
        return t.flip(-1).permute(2, 0, 1)
    return t


def _from_flat(t: torch.Tensor, how) -> torch.Tensor:
    if how == "t210":
        return t.permute(2, 1, 0)
    if how == "gate":
        return _gate_t(t).reshape(256, 64)
    if how == "convT":
        return t.permute(1, 2, 0).flip(-1)
    return t


def flatten_generator(sd: Dict[str, torch.Tensor], device) -> torch.Tensor:
    flat = torch.zeros(L.GT_SIZE, dtype=torch.float32, device=device)
    for name, (off, shape, how) in _gt_slices().items():
        t = _to_flat(sd[name].detach().to(device=device, dtype=torch.float32), how)
        flat[off:off + t.numel()] = t.reshape(-1)
    return flat


def unflatten_generator(flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    out = {}
    for name, (off, shape, how) in _gt_slices().items():
        n = 1
        for d in shape:
            n *= d
        out[name] = _from_flat(flat[off:off + n].reshape(shape), how).contiguous().clone()
    return out


class Trainer:
    """`train_one_epoch`'s loop body (py/main16.py:238-278) for a Generator / Detector pair, on the device:

        tr = Trainer(generator, detector)                  # modules with the reference's state-dict layout
        for s in loader:                                    # s (B,1,T) or (B,T) on cuda
            message = torch.randint(0, 2 ** 16, (s.shape[0],), device=s.device)
            losses = tr.step(s, message)                    # {"total","raw_total","l1","mel","loud","loc","bce","hf"}
        tr.write_back(generator, detector)

    Both networks run in train mode (batch-statistics BatchNorm), every gradient comes from hand-written backward
    kernels, and one Adam (lr 1e-3) updates both flat parameter buffers.  With torch.distributed initialised
    (one process per GPU, NCCL), `step` averages the two gradient buffers across ranks between backward and Adam —
    data-parallel training with per-replica BatchNorm statistics, as DistributedDataParallel would do."""

    LOSS_KEYS = ("l1", "mel", "loud", "loc", "bce", "hf", "total", "raw_total")

    def __init__(self, generator: torch.nn.Module, detector: torch.nn.Module, lr: float = LR, betas=(0.9, 0.999),
                 eps: float = 1e-8, lambdas=None, device=None):
        from . import functional as Fn
        from . import packing
        gsd, dsd = generator.state_dict(), detector.state_dict()
        self.nout = int(dsd["model.3.weight"].shape[0])
        if device is None:
            device = dsd["model.3.weight"].device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("Trainer runs on a B200 (sm_100a) only; there is no CPU fallback — move the modules to "
                               "cuda first")
        if tuple(gsd["embedding.weight"].shape) != (65536, 64):
            raise ValueError("Trainer supports the reference's 16-bit message embedding (65536 x 64) only")
        L.load()
        self.device = device
        self.g_params, self.d_params = flatten_generator(gsd, device), flatten_detector(dsd, self.nout, device)
        self.g_grads, self.d_grads = torch.zeros_like(self.g_params), torch.zeros_like(self.d_params)
        self.g_m, self.g_v = torch.zeros_like(self.g_params), torch.zeros_like(self.g_params)
        self.d_m, self.d_v = torch.zeros_like(self.d_params), torch.zeros_like(self.d_params)
        self.g_stats = torch.zeros(L.GT_STATS, dtype=torch.float32, device=device)
        self.d_stats = torch.zeros(L.DT_STATS, dtype=torch.float32, device=device)
        for name, off in _gt_stat_slices().items():
            self.g_stats[off:off + 64] = gsd[name].to(device=device, dtype=torch.float32)
        for name, off in _dt_stat_slices().items():
            self.d_stats[off:off + 64] = dsd[name].to(device=device, dtype=torch.float32)
        lam = dict(l1=LAMBDA_L1, msspec=LAMBDA_MSSPEC, loud=LAMBDA_LOUD, loc=LAMBDA_LOC, dec=LAMBDA_DEC, hf=HF_PENALTY_W)
        lam.update(lambdas or {})
        self.lambdas = lam
        self._lam = (C.c_float * 6)(lam["l1"], lam["msspec"], lam["loud"], lam["loc"], lam["dec"], lam["hf"])
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.fir = Fn.fir_taps_on(device)
        fb, band = packing.mel_filterbank(513, 64, 16000)
        self.mel_fb, self.mel_band = fb.to(device).contiguous(), band.to(device=device, dtype=torch.int32).contiguous()
        self.steps = 0
        self._steps0 = 0          # optimiser steps already behind the state this trainer was built from (resume)
        self._nbt0 = [{k: int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")} for sd in (gsd, dsd)]
        self._bad_message = torch.zeros((), dtype=torch.bool, device=device)
        self._ws, self._ws_key = None, None

    def _workspace(self, B: int, T: int):
        if self._ws_key != (B, T):
            n = L.load().wm_train_step_workspace_bytes(B, T, self.nout)
            self._ws, self._ws_key = (_ws(n, self.device), n), (B, T)
        return self._ws

    def forward_backward(self, s: torch.Tensor, message: torch.Tensor, want_s_w: bool = False):
        """Losses and gradients (self.g_grads / self.d_grads) of one batch; parameters untouched."""
        lib = L.load()
        s = _req(s, "s")
        if s.dim() == 3 and s.shape[1] == 1:
            s = s[:, 0]
        if s.dim() != 2:
            raise ValueError(f"s must be (B,T) or (B,1,T), got {tuple(s.shape)}")
        s = s.contiguous()
        B, T = s.shape
        msg = _req(message, "message", torch.int64)
        if msg.numel() != B:
            raise ValueError("message must hold one value per clip")
        # range check without a host sync (the loop must be free to run ahead of the GPU): out-of-range ids are
        # clamped for the kernels and remembered on the device; the error surfaces at the next host read
        # (state_dicts / write_back / check_messages)
        self._bad_message |= ((msg < 0) | (msg > 65535)).any()
        msg = msg.clamp(0, 65535)
        losses = torch.zeros(8, device=self.device)
        s_w = torch.empty_like(s) if want_s_w else None
        ws, n = self._workspace(B, T)
        L.check(lib.wm_train_forward_backward(
            L.ptr(self.g_params), L.ptr(self.g_grads), L.ptr(self.g_stats), L.ptr(self.d_params), L.ptr(self.d_grads),
            L.ptr(self.d_stats), L.ptr(s), L.ptr(msg), L.ptr(self.fir), L.ptr(self.mel_fb), L.ptr(self.mel_band),
            self.mel_fb.shape[1], C.cast(self._lam, C.c_void_p), B, T, self.nout, L.ptr(losses), L.ptr(s_w), L.ptr(ws),
            n, _stream()), "wm_train_forward_backward")
        out = {k: losses[i] for i, k in enumerate(self.LOSS_KEYS)}
        if want_s_w:
            out["s_w"] = s_w
        return out

    def all_reduce_gradients(self) -> None:
        """Average both gradient buffers over the ranks of the default process group (NCCL over NVLink on a B200 box)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        average_gradients((self.g_grads, self.d_grads))

    def apply(self) -> None:
        """torch.optim.Adam step on both parameter buffers (py/main16.py:278,504)."""
        self.steps += 1
        adam_step(self.g_params, self.g_grads, self.g_m, self.g_v, self.steps, self.lr, self.betas, self.eps)
        adam_step(self.d_params, self.d_grads, self.d_m, self.d_v, self.steps, self.lr, self.betas, self.eps)

    @ops_nvtx("train_step")
    def step(self, s: torch.Tensor, message: torch.Tensor) -> Dict[str, torch.Tensor]:
        out = self.forward_backward(s, message)
        self.all_reduce_gradients()
        self.apply()
        return out

    def check_messages(self) -> None:
        """Raises if any step so far was given a message outside [0, 65536) (what nn.Embedding would have raised at
        once; here the check is deferred so that training steps never wait for the GPU)."""
        if bool(self._bad_message):
            raise IndexError("message out of range for the 16-bit embedding in an earlier training step")

    def state_dicts(self):
        self.check_messages()
        g = unflatten_generator(self.g_params)
        for name, off in _gt_stat_slices().items():
            g[name] = self.g_stats[off:off + 64].clone()
        d = unflatten_detector(self.d_params, self.nout)
        for name, off in _dt_stat_slices().items():
            d[name] = self.d_stats[off:off + 64].clone()
        return g, d

    def grad_dicts(self):
        return unflatten_generator(self.g_grads), unflatten_detector(self.d_grads, self.nout)

    @torch.no_grad()
    def write_back(self, generator: torch.nn.Module, detector: torch.nn.Module) -> None:
        for mod, sd, nbt0 in zip((generator, detector), self.state_dicts(), self._nbt0):
            own = mod.state_dict()
            for k, v in sd.items():
                own[k].copy_(v.to(own[k].device))
            for k, n0 in nbt0.items():
                own[k].fill_(n0 + self.steps - self._steps0)


def average_gradients(buffers) -> None:
    """all_reduce(SUM) / world_size of flat gradient buffers over the default process group — the one exchange of the
    data-parallel training step.  Works on whatever backend the group was created with (NCCL on the
    GPU box; gloo in the CPU tests of this host-side logic)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    for b in buffers:
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        b.div_(world)
