"""main14b_2's configurable residual stack (py/main14b_2.py:83-224, BASELINE config 3): `ResidualBlock`, `Generator`,
`Detector` with the reference's constructor arguments, attribute and state-dict names (so its checkpoints load unchanged
and a seeded construction draws identical weights).

The torch sub-modules only own the parameters.  `Generator.forward` / `Detector.forward` run the whole model through
the tensor-core layer walk of `pconv.py` (tcgen05 implicit GEMMs over planar bf16-pair activations, `wm_pconv_fwd`) when
the layer pattern fits it and the math mode is the default; otherwise — other shapes, `WM_MATH_FP32` — layer by layer
through the fp32 CUDA operators below (`wm_conv1d_fwd`, `wm_convtranspose1d_fwd`, `wm_lstm_small_fwd`), channels-first
fp32 exactly as the reference holds its tensors.  Inference only (this model has no BatchNorm, so eval == train forward).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import pconv
from .ops import _req, _stream, get_math_mode

HIDDEN_DIM = 32            # py/main14b_2.py:40
NUM_BITS = 16              # :41
CHANNELS = 32              # :42
OUTPUT_CH = 128            # :43
STRIDES = [2, 4, 5, 8]     # :44
LSTM_LAYERS = 2            # :45


def _tensor_core_path(mod, x) -> bool:
    """The whole-model tcgen05 walk (pconv.py) in the default math mode when the layer pattern fits it; otherwise —
    and always under WM_MATH_FP32 — the layer-by-layer fp32 CUDA operators below (the exact-order cross-check)."""
    _req(x, "input")
    return get_math_mode() == L.MATH_BF16X2 and pconv.supported(mod, x.shape[-1])


def conv1d(x, conv: nn.Conv1d, act: bool = False, residual=None, chan_add=None) -> torch.Tensor:
    """y = act(conv(x) + chan_add[:, :, None] + residual) through wm_conv1d_fwd."""
    lib = L.load()
    x = _req(x, "x")
    B, Cin, Tin = x.shape
    Cout, K, s, p = conv.out_channels, conv.kernel_size[0], conv.stride[0], conv.padding[0]
    if Cin != conv.in_channels:
        raise ValueError(f"conv1d: input has {Cin} channels, the layer expects {conv.in_channels}")
    y = torch.empty(B, Cout, lib.wm_conv1d_out_len(Tin, K, s, p), device=x.device)
    res = _req(residual, "residual") if residual is not None else None
    if res is not None and res.shape != y.shape:
        raise ValueError(f"conv1d: residual {tuple(res.shape)} does not match the output {tuple(y.shape)}")
    add = _req(chan_add, "chan_add") if chan_add is not None else None
    L.check(lib.wm_conv1d_fwd(L.ptr(x), L.ptr(_req(conv.weight.detach(), "weight")), L.ptr(_req(conv.bias.detach(), "bias")),
                              L.ptr(add), L.ptr(res), L.ptr(y), B, Cin, Tin, Cout, K, s, p, 1 if act else 0, _stream()),
            "wm_conv1d_fwd")
    return y


_phase_cache = {}


def _phase_weights(ct: nn.ConvTranspose1d) -> torch.Tensor:
    """Per-phase 3-tap convolution weights of a ConvTranspose1d with kernel_size == 2 * stride, built on the device
    once per parameter version (wm_convtranspose1d_pack)."""
    lib = L.load()
    w, b = ct.weight, ct.bias
    key = (id(ct), w.data_ptr(), w._version, b.data_ptr(), b._version, str(w.device))
    hit = _phase_cache.get(id(ct))
    if hit is None or hit[0] != key:
        Cin, Cout, K = w.shape
        s, p = ct.stride[0], ct.padding[0]
        packed = torch.empty(lib.wm_convtranspose1d_phase_weight_floats(Cin, Cout, s), device=w.device)
        L.check(lib.wm_convtranspose1d_pack(L.ptr(_req(w.detach(), "weight")), L.ptr(_req(b.detach(), "bias")),
                                            L.ptr(packed), Cin, Cout, K, s, p, _stream()), "wm_convtranspose1d_pack")
        _phase_cache[id(ct)] = hit = (key, packed)
    return hit[1]


def conv_transpose1d(x, ct: nn.ConvTranspose1d, direct: bool = False) -> torch.Tensor:
    """nn.ConvTranspose1d forward.  kernel_size == 2 * stride (every layer of this model) runs as one 3-tap
    convolution over stride-many phase channels (wm_convtranspose1d_phase_fwd); `direct` forces the gather kernel."""
    lib = L.load()
    x = _req(x, "x")
    B, Cin, Tin = x.shape
    Cout, K, s, p = ct.out_channels, ct.kernel_size[0], ct.stride[0], ct.padding[0]
    y = torch.empty(B, Cout, lib.wm_convtranspose1d_out_len(Tin, K, s, p), device=x.device)
    if not direct and K == 2 * s and 0 <= p < s and ct.output_padding[0] == 0:
        L.check(lib.wm_convtranspose1d_phase_fwd(L.ptr(x), L.ptr(_phase_weights(ct)), L.ptr(y), B, Cin, Tin, Cout, K, s, p,
                                                 _stream()), "wm_convtranspose1d_phase_fwd")
        return y
    L.check(lib.wm_convtranspose1d_fwd(L.ptr(x), L.ptr(_req(ct.weight.detach(), "weight")),
                                       L.ptr(_req(ct.bias.detach(), "bias")), L.ptr(y), B, Cin, Tin, Cout, K, s, p,
                                       _stream()), "wm_convtranspose1d_fwd")
    return y


def lstm_small(x, lstm: nn.LSTM) -> torch.Tensor:
    """x (B,H,T) channels-first -> top-layer hidden states (B,H,T)."""
    lib = L.load()
    x = _req(x, "x")
    B, H, T = x.shape
    nl = lstm.num_layers
    if lstm.input_size != H or lstm.hidden_size != H or lstm.bidirectional:
        raise ValueError("lstm_small: expected a unidirectional LSTM(H, H)")
    w_ih = torch.stack([getattr(lstm, f"weight_ih_l{l}").detach() for l in range(nl)]).contiguous()
    w_hh = torch.stack([getattr(lstm, f"weight_hh_l{l}").detach() for l in range(nl)]).contiguous()
    bias = torch.stack([(getattr(lstm, f"bias_ih_l{l}") + getattr(lstm, f"bias_hh_l{l}")).detach() for l in range(nl)])
    y = torch.empty_like(x)
    L.check(lib.wm_lstm_small_fwd(L.ptr(x), L.ptr(_req(w_ih, "w_ih")), L.ptr(_req(w_hh, "w_hh")),
                                  L.ptr(_req(bias.contiguous(), "bias")), L.ptr(y), B, H, T, nl, _stream()),
            "wm_lstm_small_fwd")
    return y


def make_conv1d(in_ch, out_ch, kernel_size=3, stride=1, padding=1):
    return nn.Conv1d(in_ch, out_ch, kernel_size, stride=stride, padding=padding)


class ResidualBlock(nn.Module):
    """elu(conv2(elu(conv1(x))) + skip(x)) — py/main14b_2.py:86-105."""

    def __init__(self, in_ch, out_ch, stride=1):
        super().__init__()
        self.downsample = (stride != 1 or in_ch != out_ch)
        self.conv1 = make_conv1d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)
        self.conv2 = make_conv1d(out_ch, out_ch, kernel_size=3, stride=1, padding=1)
        self.elu = nn.ELU()
        if self.downsample:
            self.skip_conv = make_conv1d(in_ch, out_ch, kernel_size=1, stride=stride, padding=0)

    @torch.no_grad()
    def forward(self, x):
        out = conv1d(x, self.conv1, act=True)
        residual = conv1d(x, self.skip_conv) if self.downsample else x
        return conv1d(out, self.conv2, act=True, residual=residual)


def _fit_length(y: torch.Tensor, T: int) -> torch.Tensor:
    """Crop or right-zero-pad the time axis to T (py/main14b_2.py:175-181, 217-222)."""
    if y.shape[-1] > T:
        return y[:, :, :T]
    if y.shape[-1] < T:
        return torch.nn.functional.pad(y, (0, T - y.shape[-1]))
    return y


class Generator(nn.Module):
    """py/main14b_2.py:107-182."""

    def __init__(self, in_channels=1, base_channels=CHANNELS, hidden_dim=HIDDEN_DIM, message_bits=NUM_BITS,
                 output_channels=OUTPUT_CH, strides=STRIDES):
        super().__init__()
        self.message_bits = message_bits
        self.hidden_dim = hidden_dim
        self.E = nn.Embedding(num_embeddings=(2 ** message_bits), embedding_dim=hidden_dim)
        self.init_conv = nn.Conv1d(in_channels, base_channels, kernel_size=7, stride=1, padding=3)
        enc_blocks, ch = [], base_channels
        for st in strides:
            enc_blocks.append(ResidualBlock(ch, ch * 2, stride=st))
            ch *= 2
        self.encoder_blocks = nn.Sequential(*enc_blocks)
        self.proj = nn.Linear(ch, hidden_dim)
        self.lstm = nn.LSTM(input_size=hidden_dim, hidden_size=hidden_dim, num_layers=2, batch_first=True,
                            bidirectional=False)
        self.final_conv_enc = nn.Conv1d(hidden_dim, output_channels, kernel_size=7, stride=1, padding=3)
        dec_blocks, in_ch = [], output_channels
        for st in reversed(list(strides)):
            out_ch = in_ch // 2
            dec_blocks.append(nn.ConvTranspose1d(in_ch, out_ch, kernel_size=2 * st, stride=st, padding=(st // 2),
                                                 output_padding=0))
            dec_blocks.append(ResidualBlock(out_ch, out_ch, stride=1))
            in_ch = out_ch
        self.decoder_blocks = nn.Sequential(*dec_blocks)
        self.final_conv_dec = nn.Conv1d(in_ch, 1, kernel_size=7, stride=1, padding=3)

    @torch.no_grad()
    def forward(self, s, message=None):
        if s.dim() != 3:
            raise ValueError(f"s: expected (B, C, T), got {tuple(s.shape)}")
        B, _, T = s.shape
        if _tensor_core_path(self, s):
            return pconv.generator_forward(self, _req(s, "s"), message)
        x = conv1d(s, self.init_conv)
        for blk in self.encoder_blocks:
            x = blk(x)
        # proj is a Linear over channels = a 1-tap convolution; the message embedding is a per-clip channel offset
        proj = _LinearAsConv(self.proj)
        e = self.E.weight.detach()[message.to(torch.int64)] if message is not None else None
        x = conv1d(x, proj, chan_add=e)
        x = lstm_small(x, self.lstm)
        x = conv1d(x, self.final_conv_enc)
        for blk in self.decoder_blocks:
            x = conv_transpose1d(x, blk) if isinstance(blk, nn.ConvTranspose1d) else blk(x)
        return _fit_length(conv1d(x, self.final_conv_dec), T)


class _LinearAsConv:
    """View of nn.Linear(C, H) as Conv1d(C, H, 1) for conv1d()."""

    def __init__(self, lin: nn.Linear):
        self.in_channels, self.out_channels = lin.in_features, lin.out_features
        self.kernel_size, self.stride, self.padding = (1,), (1,), (0,)
        self.weight, self.bias = lin.weight, lin.bias      # (H, C) == (H, C, 1) contiguous


class Detector(nn.Module):
    """py/main14b_2.py:184-224: logits (B, 1 + bits, T), channel-first, raw (no sigmoid)."""

    def __init__(self, in_channels=1, base_channels=CHANNELS, hidden_dim=HIDDEN_DIM, message_bits=NUM_BITS,
                 strides=STRIDES):
        super().__init__()
        self.message_bits = message_bits
        self.init_conv = nn.Conv1d(in_channels, base_channels, kernel_size=7, stride=1, padding=3)
        enc_blocks, ch = [], base_channels
        for st in strides:
            enc_blocks.append(ResidualBlock(ch, ch * 2, stride=st))
            ch *= 2
        self.encoder_blocks = nn.Sequential(*enc_blocks)
        dec_blocks, in_ch = [], ch
        for st in reversed(list(strides)):
            out_ch = in_ch // 2
            dec_blocks.append(nn.ConvTranspose1d(in_ch, out_ch, kernel_size=2 * st, stride=st, padding=(st // 2),
                                                 output_padding=0))
            dec_blocks.append(ResidualBlock(out_ch, out_ch, stride=1))
            in_ch = out_ch
        self.upsample_blocks = nn.Sequential(*dec_blocks)
        self.final_conv = nn.Conv1d(base_channels, 1 + message_bits, kernel_size=7, stride=1, padding=3)

    @torch.no_grad()
    def forward(self, x):
        if x.dim() != 3:
            raise ValueError(f"x: expected (B, C, T), got {tuple(x.shape)}")
        T = x.shape[-1]
        if _tensor_core_path(self, x):
            return pconv.detector_forward(self, _req(x, "x"))
        x = conv1d(x, self.init_conv)
        for blk in self.encoder_blocks:
            x = blk(x)
        for blk in self.upsample_blocks:
            x = conv_transpose1d(x, blk) if isinstance(blk, nn.ConvTranspose1d) else blk(x)
        return _fit_length(conv1d(x, self.final_conv), T)


class GraphedEmbedDetect:
    """`logits = D(s + G(s, message))` for a FIXED batch shape as one CUDA graph: the ~46 kernel launches of the layer walk
    are captured once and replayed, so a call costs the host one launch (the eager walk costs ~2 ms of Python per pass,
    which is what bounds small batches — and large ones on a slow host).  Inputs are copied into static buffers; the
    returned tensors are static too and are OVERWRITTEN by the next call (clone what must survive).  Inference only;
    rebuild after changing parameters."""

    def __init__(self, generator: "Generator", detector: "Detector", B: int, T: int = 16000, device=None):
        dev = torch.device(device) if device is not None else next(generator.parameters()).device
        self.G, self.D = generator.eval(), detector.eval()
        self.s = torch.zeros(B, 1, T, device=dev)
        self.message = torch.zeros(B, dtype=torch.int64, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                      # weight packing, function attributes, allocator warm-up
                self._walk()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.delta, self.logits = self._walk()

    def _walk(self):
        delta = self.G(self.s, self.message)
        return delta, self.D(self.s + delta)

    def __call__(self, s: torch.Tensor, message: torch.Tensor):
        """-> (delta (B,1,T), logits (B,1+bits,T)), both static buffers"""
        self.s.copy_(s, non_blocking=True)
        self.message.copy_(message, non_blocking=True)
        self.graph.replay()
        return self.delta, self.logits
