"""`torch.autograd.Function`s over the library's forward / backward entry points, so that the reference's training
loop body (py/main16.py:238-278) runs UNMODIFIED on these modules: `generator.train()`, `delta = generator(s, message)`,
`fir_lowpass` / `clamp_peak` / `limit_rms`, `detector(torch.cat([s_w, s]))`, the loss objects, `loss.backward()`,
`torch.optim.Adam.step()` ("training path additionally under autograd").

Every Function's forward and backward is one C-ABI call (the same kernels `wmb200.Trainer` strings together inside
`wm_train_forward_backward`); autograd only carries the graph.  `Trainer.step` remains the fast path (one call per
iteration, no per-operator allocation); this module is the drop-in path and is tested against it.

Internal activations are channels-last (B, T, 64), as in the kernels.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from . import train as TR


class ConvIn(torch.autograd.Function):
    """Conv1d(1, 64, 7, padding=3) on s (B,T) -> (B,T,64).  weight (64,1,7), bias (64,)  (py/main16.py:134,177)."""

    @staticmethod
    def forward(ctx, s, weight, bias):
        ctx.save_for_backward(s, weight)
        return ops.conv_in_k7(s, weight.permute(2, 1, 0).reshape(7, 64).contiguous(), bias)

    @staticmethod
    def backward(ctx, dy):
        s, weight = ctx.saved_tensors
        dw, db, ds = TR.conv_in_k7_bwd(s, dy.contiguous(), weight, want_ds=ctx.needs_input_grad[0])
        return ds, dw, db


class Conv64(torch.autograd.Function):
    """Conv1d(64, 64, K, padding=K//2), K in {1,3,7}, on channels-last x.  weight (co,ci,K)  (py/main16.py:116,119)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return TR.conv64_train_fwd(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dw, db, dx = TR.conv64_bwd(x, dy.contiguous(), weight, want_dx=ctx.needs_input_grad[0])
        return dx, dw, db


class BatchNormTrain(torch.autograd.Function):
    """relu?(BatchNorm1d(64) with batch statistics (z) + residual) on channels-last z; running statistics (momentum
    0.1, unbiased variance) are updated in place  (py/main16.py:117,120,124-125 in train mode)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, residual, relu, running_mean, running_var):
        out, mean, rstd = TR.bn_train_fwd(z, gamma, beta, residual, relu, running_mean, running_var)
        ctx.save_for_backward(out, z, mean, rstd, gamma)
        ctx.relu, ctx.has_res = relu, residual is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        out, z, mean, rstd, gamma = ctx.saved_tensors
        dz, dres, dg, db = TR.bn_train_bwd(dout.contiguous(), out if ctx.relu else None, z, mean, rstd, gamma,
                                           want_residual_grad=ctx.has_res and ctx.needs_input_grad[3])
        return dz, dg, db, dres, None, None, None


class LSTM(torch.autograd.Function):
    """nn.LSTM(64, 64, batch_first=True)(x)[0], zero initial state  (py/main16.py:138,153)."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh):
        h, saved = TR.lstm_train_fwd(x, w_ih, w_hh, b_ih, b_hh)
        ctx.saved = saved
        return h

    @staticmethod
    def backward(ctx, dy):
        dx, dwi, dwh, db = TR.lstm_train_bwd(dy.contiguous(), ctx.saved)
        ctx.saved = None
        return dx, dwi, dwh, db, db.clone()


class Head(torch.autograd.Function):
    """Conv1d(64, nout, 1) on channels-last y -> (B,T,nout).  weight (nout,64,1)  (py/main16.py:146,180)."""

    @staticmethod
    def forward(ctx, y, weight, bias):
        ctx.save_for_backward(y, weight)
        return ops.head(y, weight.reshape(weight.shape[0], 64).contiguous(), bias)

    @staticmethod
    def backward(ctx, dlogits):
        y, weight = ctx.saved_tensors
        dy, dw, db = TR.head_bwd(dlogits.contiguous(), y, weight)
        return dy, dw, db


class Postprocess(torch.autograd.Function):
    """fir_lowpass / clamp_peak / limit_rms (any combination `mode` of them, applied in that order) on delta (B,T)
    (py/main16.py:53-72)."""

    @staticmethod
    def forward(ctx, delta, fir, mode, peak, max_rms, eps):
        ctx.save_for_backward(delta, fir if fir is not None else delta.new_empty(0))
        ctx.args = (mode, peak, max_rms, eps, fir is not None)
        return ops.postprocess(delta, None, fir, mode, True, False, peak=peak, max_rms=max_rms, eps=eps)[0]

    @staticmethod
    def backward(ctx, g):
        delta, fir = ctx.saved_tensors
        mode, peak, max_rms, eps, has_fir = ctx.args
        return (TR.postprocess_bwd(g.contiguous(), delta, fir if has_fir else None, mode, peak, max_rms, eps), None, None,
                None, None, None)


class HighFreqPenalty(torch.autograd.Function):
    """high_freq_penalty(delta) for delta (B,T)  (py/main16.py:74-81)."""

    @staticmethod
    def forward(ctx, delta, n_fft, first_bin):
        ctx.save_for_backward(delta)
        ctx.args = (n_fft, first_bin)
        return ops.hf_penalty(delta, n_fft, first_bin)

    @staticmethod
    def backward(ctx, g):
        (delta,) = ctx.saved_tensors
        return TR.hf_penalty_bwd(delta, ctx.args[0], ctx.args[1]) * g, None, None


class MelLogL1(torch.autograd.Function):
    """MultiScaleMelLoss()(clean, watermarked); the gradient goes to `watermarked` (py/main16.py:192-202, 268)."""

    @staticmethod
    def forward(ctx, clean, wm, fb, band, n_fft, hop):
        ctx.save_for_backward(clean, wm, fb, band)
        ctx.args = (n_fft, hop)
        return ops.mel_log_l1(clean, wm, fb, band, n_fft, hop)

    @staticmethod
    def backward(ctx, g):
        clean, wm, fb, band = ctx.saved_tensors
        return None, TR.mel_log_l1_bwd(clean, wm, fb, band, ctx.args[0], ctx.args[1]) * g, None, None, None, None


class Loudness(torch.autograd.Function):
    """TFLoudnessLoss()(clean, watermarked); the gradient goes to `watermarked` (py/main16.py:204-217, 269)."""

    @staticmethod
    def forward(ctx, clean, wm, n_fft, hop, thresh):
        ctx.save_for_backward(clean, wm)
        ctx.args = (n_fft, hop, thresh)
        return ops.loudness(clean, wm, n_fft, hop, thresh)

    @staticmethod
    def backward(ctx, g):
        clean, wm = ctx.saved_tensors
        return None, TR.loudness_bwd(clean, wm, *ctx.args) * g, None, None, None


def _bump(bn: torch.nn.BatchNorm1d) -> None:
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1


def resblock_train(rb, x: torch.Tensor) -> torch.Tensor:
    """ResBlock.forward in train mode on channels-last x (py/main16.py:112-125)."""
    c1, bn1, _, c2, bn2 = rb.block
    z1 = Conv64.apply(x, c1.weight, c1.bias)
    u = BatchNormTrain.apply(z1, bn1.weight, bn1.bias, None, True, bn1.running_mean, bn1.running_var)
    z2 = Conv64.apply(u, c2.weight, c2.bias)
    y = BatchNormTrain.apply(z2, bn2.weight, bn2.bias, x, True, bn2.running_mean, bn2.running_var)
    _bump(bn1)
    _bump(bn2)
    return y


def generator_train(gen, s: torch.Tensor, message) -> torch.Tensor:
    """Generator.forward in train mode: s (B,T) -> delta_raw (B,T)  (py/main16.py:149-162)."""
    enc, dec = gen.encoder, gen.decoder
    x = ConvIn.apply(s, enc[0].weight, enc[0].bias)
    x = resblock_train(enc[1], x)
    x = resblock_train(enc[2], x)
    lstm = gen.lstm
    h = LSTM.apply(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0)
    if gen.message_bits > 0 and message is not None:
        h = h + gen.embedding(message).unsqueeze(1)                   # broadcast over time (:156-159)
    # ConvTranspose1d(64,64,7,padding=3), stride 1 == Conv1d with the (in,out) axes swapped and the taps flipped
    w_ct = dec[0].weight.permute(1, 0, 2).flip(-1)
    x = Conv64.apply(h.contiguous(), w_ct, dec[0].bias)
    x = resblock_train(dec[1], x)
    return Head.apply(x, dec[2].weight, dec[2].bias)[..., 0]


def detector_train(det, x: torch.Tensor) -> torch.Tensor:
    """Detector.forward in train mode: x (B,T) -> logits (B,T,1+bits)  (py/main16.py:183-186)."""
    m = det.model
    y = ConvIn.apply(x, m[0].weight, m[0].bias)
    y = resblock_train(m[1], y)
    y = resblock_train(m[2], y)
    return Head.apply(y, m[3].weight, m[3].bias)


def needs_graph(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)
