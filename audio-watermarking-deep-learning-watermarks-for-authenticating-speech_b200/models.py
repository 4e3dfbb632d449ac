"""`Generator` / `Detector` with the reference's constructor, attributes, state-dict
keys and call signatures (py/main16.py:112-186), running on libwmb200.

The torch sub-modules below exist to own the parameters under the reference's names
(`encoder.1.block.0.weight`, `lstm.weight_ih_l0`, `model.3.bias`, ...), so that the
reference's `.pth` files load unchanged; they are never called.  In eval mode `forward`
packs the parameters once per parameter version (eval BatchNorm folded, see packing.py)
and hands raw pointers to the fused inference kernels; in train mode it runs the
batch-statistics path operator by operator under autograd (autograd.py), so the
reference's loop body (py/main16.py:238-278) works on these modules as written.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import autograd as AG
from . import ops, packing

HIDDEN = 64


def _conv_bn_pair(ch: int):
    return [nn.Conv1d(ch, ch, 3, padding=1), nn.BatchNorm1d(ch)]


class ResBlock(nn.Module):
    """relu(x + BN(conv3(relu(BN(conv3(x))))))  — py/main16.py:112-125 (eval BatchNorm)."""

    def __init__(self, ch: int):
        super().__init__()
        if ch != HIDDEN:
            raise ValueError("wmb200 implements the reference's 64-channel blocks only")
        self.block = nn.Sequential(*_conv_bn_pair(ch), nn.ReLU(), *_conv_bn_pair(ch))
        self.relu = nn.ReLU()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x (B,64,T) -> (B,64,T) through two wm_conv64_fwd launches (train mode: batch statistics, under autograd)."""
        if self.training:
            return AG.resblock_train(self, x.permute(0, 2, 1).contiguous()).permute(0, 2, 1)
        sd = {"rb." + k: v for k, v in self.state_dict().items()}
        w1, b1 = packing.fold_conv_bn(sd, "rb.block.0", "rb.block.1")
        w2, b2 = packing.fold_conv_bn(sd, "rb.block.3", "rb.block.4")
        dev = x.device
        f = lambda t: t.to(torch.float32).to(dev)
        xl = x.permute(0, 2, 1).contiguous()
        y = ops.conv64(xl, f(w1), f(b1), taps=3, relu=True)
        y = ops.conv64(y, f(w2), f(b2), residual=xl, taps=3, relu=True)
        return y.permute(0, 2, 1)


def _no_training(m: nn.Module) -> None:
    if m.training:
        raise RuntimeError(
            "this entry point is the fused inference path (eval-mode BatchNorm, py/main16.py:979,1115); call .eval() "
            "first.  Train-mode modules run through forward() (autograd) or wmb200.Trainer.step")


class _Packed(nn.Module):
    """Caches the packed weight blob on the module's device, keyed by parameter versions."""

    def __init__(self):
        super().__init__()
        self._blob = None
        self._blob_key = None

    def _pack(self, sd):
        raise NotImplementedError

    def packed(self) -> torch.Tensor:
        tensors = list(self.parameters()) + list(self.buffers())
        dev = tensors[0].device
        key = (str(dev),) + tuple((t.data_ptr(), t._version) for t in tensors)
        if self._blob is None or self._blob_key != key:
            sd = {k: v for k, v in self.state_dict().items() if k != "embedding.weight"}
            host = self._pack(sd)                      # fp32 part, packed on the host
            blob = torch.zeros(self._blob_floats, dtype=torch.float32, device=dev)
            blob[:host.numel()] = host.to(dev)
            if dev.type == "cuda":                     # tcgen05 weight images, built on the device
                with torch.cuda.device(dev):
                    L.check(getattr(L.load(), self._finalize)(blob.data_ptr(), torch.cuda.current_stream().cuda_stream),
                            self._finalize)
            self._blob = blob
            self._blob_key = key
        return self._blob

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts the reference's checkpoints as shipped: keys saved from a
        torch.compile'd module carry `_orig_mod.` (py/main16.py:551-555, 707-712)."""
        return super().load_state_dict(packing.strip_prefix(state_dict), strict=strict, **kw)


def _as_bt(x: torch.Tensor, name: str) -> torch.Tensor:
    if x.dim() != 3 or x.shape[1] != 1:
        raise ValueError(f"{name}: expected (B, 1, T), got {tuple(x.shape)}")
    return x[:, 0, :]


class Generator(_Packed):
    """Encoder -> LSTM -> (+ message embedding) -> decoder; returns the watermark delta.
    Constructor / forward signature of py/main16.py:128-162."""

    def __init__(self, message_bits: int = 0):
        super().__init__()
        self.message_bits = message_bits
        self.encoder = nn.Sequential(nn.Conv1d(1, HIDDEN, 7, padding=3), ResBlock(HIDDEN), ResBlock(HIDDEN))
        self.lstm = nn.LSTM(HIDDEN, HIDDEN, batch_first=True)
        if message_bits > 0:
            self.embedding = nn.Embedding(2 ** message_bits, HIDDEN)
        self.decoder = nn.Sequential(nn.ConvTranspose1d(HIDDEN, HIDDEN, 7, padding=3), ResBlock(HIDDEN),
                                     nn.Conv1d(HIDDEN, 1, 1))

    _blob_floats = L.G_BLOB
    _finalize = "wm_finalize_generator_blob"

    def _pack(self, sd):
        return packing.pack_generator(sd)

    def embedding_table(self) -> Optional[torch.Tensor]:
        return self.embedding.weight.detach() if self.message_bits > 0 else None

    def forward(self, s: torch.Tensor, message: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = _as_bt(s, "s")
        use_msg = self.message_bits > 0 and message is not None
        if use_msg and message.shape != (x.shape[0],):
            raise ValueError(f"message: expected shape ({x.shape[0]},), got {tuple(message.shape)}")
        if self.training:       # batch-statistics BatchNorm, operator by operator under autograd (py/main16.py:244)
            return AG.generator_train(self, ops._req(x, "s"), message if use_msg else None).unsqueeze(1)
        with torch.no_grad():
            delta = ops.generator_fwd(self.packed(), self.embedding_table() if use_msg else None,
                                      message if use_msg else None, x)
        return delta.unsqueeze(1)


class Detector(_Packed):
    """Residual conv stack -> per-sample logits (B, T, 1 + message_bits); channel 0 is the
    detection logit, channels 1.. the message bits LSB first.  py/main16.py:170-186."""

    def __init__(self, message_bits: int = 0):
        super().__init__()
        self.message_bits = message_bits
        if 1 + message_bits > L.MAX_HEAD:
            raise ValueError(f"message_bits must be <= {L.MAX_HEAD - 1}")
        self.model = nn.Sequential(nn.Conv1d(1, HIDDEN, kernel_size=7, padding=3), ResBlock(HIDDEN),
                                   ResBlock(HIDDEN), nn.Conv1d(HIDDEN, 1 + message_bits, kernel_size=1))

    _blob_floats = L.D_BLOB
    _finalize = "wm_finalize_detector_blob"

    def _pack(self, sd):
        return packing.pack_detector(sd)

    @property
    def nout(self) -> int:
        return 1 + self.message_bits

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:       # py/main16.py:250 in the training loop
            return AG.detector_train(self, ops._req(_as_bt(x, "x"), "x"))
        with torch.no_grad():
            return ops.detector_fwd(self.packed(), _as_bt(x, "x"), self.nout)

    @torch.no_grad()
    def detect(self, x: torch.Tensor, valid_len: Optional[torch.Tensor] = None, want_probs: bool = True,
               want_votes: bool = True) -> dict:
        """Fused fast path: sigmoid(ch 0), clip means and message-logit means without the
        (B,T,1+bits) logits tensor (what detect_watermark / evaluate_model consume)."""
        _no_training(self)
        return ops.detect_fwd(self.packed(), _as_bt(x, "x"), self.nout, valid_len, want_probs, want_votes)


def load_state_dict_strip_prefix(model: nn.Module, state_dict, prefix: str = "_orig_mod."):
    """py/main16.py:707-712: drop the torch.compile prefix, load non-strictly."""
    cleaned = {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in state_dict.items()}
    return nn.Module.load_state_dict(model, cleaned, strict=False)
