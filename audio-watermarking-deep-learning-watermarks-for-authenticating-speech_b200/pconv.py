"""main14b_2's residual stack (py/main14b_2.py:83-224, BASELINE config 3) on the tensor cores.

Every channel-heavy layer becomes one implicit GEMM of `wm_pconv_fwd` (csrc/wm_pconv_tc.cu) over "planar" bf16-pair
activations; this module holds the host side of that: the row geometry, the translation of each reference layer into
(sources, taps, GEMM weights), weight packing per parameter version, and the layer walk of Generator / Detector.

  * Conv1d k3 stride s (ResidualBlock.conv1, :90): the producer wrote its rows split by phase, so tap 0 reads phase
    s-1 one row up, tap 1 phase 0, tap 2 phase 1 — three single-tap sources.
  * conv2 + skip_conv (:91,95,100-103): one GEMM whose K is conv2's 3 x Cout plus the 1x1 stride-s skip's Cin
    (phase 0 of the block input); the residual never goes through memory.
  * ConvTranspose1d(k = 2s, stride s, padding s//2) (:146,201): output step s q + ph gets inputs q-1, q (ph + p < s) or
    q, q+1 (otherwise): a 2-tap convolution over s * Cout phase columns (3-tap with zero weights where an N chunk
    mixes the two kinds).
  * the k7 output convolutions (:149,205) and everything with fewer than 16 input channels stay on the fp32 kernels.

The GEMM descriptions are plain torch code (no CUDA needed), so tests/test_pconv_host.py checks them on the CPU against
torch's own convolutions through an emulation of wm_pconv_fwd's contract.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

GAP = 4                 # WM_PC_GAP
OUT_PLANAR, OUT_CONVT, OUT_FP32 = 0, 2, 3


def plane_rows(B: int, T: int) -> int:
    """wm_pconv_plane_rows"""
    R = B * (T + GAP) + GAP
    return (R + 7) // 8 * 8 + 144


@dataclass
class Gemm:
    """One wm_pconv_fwd call minus its buffers: sources as (source id, channels, row offset, taps), fp32 GEMM weights
    wd[chunk][slice][16][nc] (slice = source-major, 16-channel group, tap), bias per column."""
    srcs: List[Tuple[int, int, int, int]]
    wd: torch.Tensor
    bias: torch.Tensor
    nc: int
    n_total: int
    chunk_off: List[int] = field(default_factory=list)
    cout: int = 0            # real output channels (before padding to a multiple of 16)
    img: Optional[torch.Tensor] = None     # packed on the device by the backend
    dbias: Optional[torch.Tensor] = None
    interleave: int = 1      # transposed convolutions: phases interleaved per 8-channel group (convt_columns)
    cols: Optional[list] = None


def _nc_for(n: int) -> int:
    for nc in (128, 64, 32, 16):
        if n % nc == 0 and n // nc <= 64:
            return nc
    raise ValueError(f"pconv: {n} GEMM columns cannot be cut into chunks of 16..128")


def _pad_out(w: torch.Tensor, b: torch.Tensor):
    cout = w.shape[0]
    n = (cout + 15) // 16 * 16
    if n != cout:
        w = torch.cat([w, w.new_zeros((n - cout,) + tuple(w.shape[1:]))])
        b = torch.cat([b, b.new_zeros(n - cout)])
    return w, b, cout, n


def gemm_conv_s1(w: torch.Tensor, b: torch.Tensor) -> Gemm:
    """Conv1d(Cin, Cout, K, stride 1, padding K//2); w (Cout, Cin, K)."""
    w, b, cout, n = _pad_out(w.detach().float(), b.detach().float())
    _, cin, K = w.shape
    nc = _nc_for(n)
    wd = w.reshape(n // nc, nc, cin // 16, 16, K).permute(0, 2, 4, 3, 1).reshape(n // nc, cin // 16 * K, 16, nc)
    return Gemm([(0, cin, -(K // 2), K)], wd.contiguous(), b, nc, n, [0] * (n // nc), cout)


def gemm_conv_strided(w: torch.Tensor, b: torch.Tensor, s: int) -> Gemm:
    """Conv1d(Cin, Cout, 3, stride s, padding 1) on an input split into s phases: source ids are phases."""
    w, b, cout, n = _pad_out(w.detach().float(), b.detach().float())
    _, cin, K = w.shape
    assert K == 3 and s >= 2
    nc = _nc_for(n)
    wd = w.reshape(n // nc, nc, cin // 16, 16, 3).permute(0, 4, 2, 3, 1).reshape(n // nc, 3 * (cin // 16), 16, nc)
    return Gemm([(s - 1, cin, -1, 1), (0, cin, 0, 1), (1, cin, 0, 1)], wd.contiguous(), b, nc, n, [0] * (n // nc), cout)


def gemm_conv2_skip(w2, b2, ws, bs) -> Gemm:
    """conv2 (k3, stride 1) of a down-sampling ResidualBlock plus its 1x1 stride-s skip_conv as extra K: source 0 =
    elu(conv1(x)), source 1 = phase 0 of the block input."""
    g = gemm_conv_s1(w2, b2)
    ws, bs, cout, n = _pad_out(ws.detach().float(), bs.detach().float())
    assert n == g.n_total and ws.shape[2] == 1
    cinx = ws.shape[1]
    nc = g.nc
    wsk = ws[:, :, 0].reshape(n // nc, nc, cinx // 16, 16).permute(0, 2, 3, 1)
    return Gemm(g.srcs + [(1, cinx, 0, 1)], torch.cat([g.wd, wsk], dim=1).contiguous(), g.bias + bs, nc, n,
                [0] * (n // nc), cout)


def convt_columns(s: int, cout: int, pk: int) -> List[Tuple[int, int]]:
    """(phase, channel) of every GEMM column of a transposed convolution laid out with `pk` phases interleaved per
    8-channel group: n = kb * (cout * pk) + g * (8 * pk) + phl * 8 + c  <->  phase kb * pk + phl, channel 8 g + c.
    pk = 1 is phase-major (n = phase * cout + co).  With pk >= 2 a thread of the epilogue holds consecutive phases of the
    same 8 channels, i.e. consecutive output rows: it stores them 32 bytes at a time."""
    assert s % pk == 0 and cout % 8 == 0
    return [(kb * pk + phl, g * 8 + c) for kb in range(s // pk) for g in range(cout // 8) for phl in range(pk)
            for c in range(8)]


def convt_decode(n: int, cout: int, pk: int) -> Tuple[int, int]:
    """the kernel's decoding of column n (csrc/wm_pconv_tc.cu, WM_PC_OUT_CONVT)"""
    blk = cout * pk
    kb, rem = divmod(n, blk)
    g, r2 = divmod(rem, 8 * pk)
    return kb * pk + r2 // 8, g * 8 + r2 % 8


def gemm_convT(w: torch.Tensor, b: torch.Tensor, s: int, p: int) -> Gemm:
    """ConvTranspose1d(Cin, Cout, 2s, stride s, padding p); w (Cin, Cout, 2s).  Columns = (phase, channel) pairs in the
    order of convt_columns(s, Cout, interleave)."""
    w, b = w.detach().float(), b.detach().float()
    cin, cout, K = w.shape
    assert K == 2 * s and 0 <= p < s and cout % 8 == 0
    w3 = w.new_zeros(3, cin, s, cout)
    for j in range(3):
        for ph in range(s):
            k = s * (1 - j) + ph + p
            if 0 <= k < K:
                w3[j, :, ph, :] = w[:, :, k]
    n = s * cout
    nc = _nc_for(n)
    nch = n // nc

    def kinds_of(cols):
        return [{ph + p >= s for ph, _ in cols[j * nc:(j + 1) * nc]} for j in range(nch)]

    # candidates: phases of one kind (same pair of input rows) interleaved per 8 channels, phase-major, fully interleaved
    choice = None
    for pk in ([s // 2] if s % 2 == 0 and s >= 4 else []) + [1]:
        cols = convt_columns(s, cout, pk)
        kinds = kinds_of(cols)
        if all(len(k) == 1 for k in kinds):
            choice = (pk, cols, [1 if True in k else 0 for k in kinds], 2)
            break
    if choice is None:
        pk = s if s % 2 == 0 else 1
        choice = (pk, convt_columns(s, cout, pk), [0] * nch, 3)
    pk, cols, offs, taps = choice
    ph_idx = torch.tensor([c[0] for c in cols])
    co_idx = torch.tensor([c[1] for c in cols])
    wconv = w3[:, :, ph_idx, co_idx].permute(2, 1, 0)                            # (n, cin, 3)
    wd5 = wconv.reshape(nch, nc, cin // 16, 16, 3).permute(0, 2, 4, 3, 1)       # chunk, kc, tap, 16, nc
    if taps == 2:
        wd5 = torch.stack([wd5[j, :, offs[j]:offs[j] + 2] for j in range(nch)])
    wd = wd5.reshape(nch, (cin // 16) * taps, 16, nc).contiguous()
    g = Gemm([(0, cin, -1, taps)], wd, b[co_idx], nc, n, offs, cout)
    g.interleave, g.cols = pk, cols
    return g


# ---------------------------------------------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------------------------------------------
class _Src(C.Structure):
    _fields_ = [("base", C.c_void_p), ("cin", C.c_int), ("row_off", C.c_int), ("taps", C.c_int), ("reserved", C.c_int)]


class _Desc(C.Structure):
    _fields_ = [("src", _Src * 3), ("nsrc", C.c_int), ("B", C.c_int), ("T", C.c_int), ("plane_rows", C.c_longlong),
                ("w", C.c_void_p), ("bias", C.c_void_p), ("n_total", C.c_int), ("nc", C.c_int),
                ("chunk_off", C.c_byte * 64), ("elu", C.c_int), ("mode", C.c_int), ("residual", C.c_void_p),
                ("y", C.c_void_p), ("out_plane_rows", C.c_longlong), ("out_phase_rows", C.c_longlong),
                ("out_split", C.c_int), ("ct_stride", C.c_int), ("ct_pad", C.c_int), ("ct_cout", C.c_int),
                ("out_T", C.c_int), ("fused", C.c_int), ("ct_interleave", C.c_int), ("w2", C.c_void_p), ("bias2", C.c_void_p),
                ("skip", _Src)]


class Planar:
    """A planar tensor: C channels, B clips of T rows, optionally `split` phase buffers (T = rows per phase)."""

    def __init__(self, Cn: int, B: int, T: int, split: int = 1, device=None):
        if Cn % 8:
            raise ValueError(f"planar tensors hold a multiple of 8 channels (got {Cn})")
        self.C, self.B, self.T, self.split = Cn, B, T, split
        self.RP = plane_rows(B, T)
        self.phase_rows = 2 * (Cn // 8) * self.RP
        self.store = torch.empty((split * self.phase_rows + 16) * 16, dtype=torch.uint8, device=device)

    def ptr(self, phase: int = 0) -> int:
        return self.store.data_ptr() + 128 + phase * self.phase_rows * 16


class CudaBackend:
    """The product path: every operation is a libwmb200 call on the current stream."""

    def planar(self, Cn, B, T, split, device):
        return Planar(Cn, B, T, split, device)

    def fp32(self, shape, device):
        return torch.empty(shape, device=device, dtype=torch.float32)

    def pack(self, g: Gemm, device):
        from . import _lib as L
        from .ops import _stream
        if g.img is None or g.img.device != torch.device(device):
            lib = L.load()
            wd = g.wd.to(device).contiguous()
            nsl = wd.shape[0] * wd.shape[1]
            img = torch.empty(lib.wm_pconv_weight_bytes(nsl, g.nc), dtype=torch.uint8, device=device)
            L.check(lib.wm_pconv_pack(wd.data_ptr(), img.data_ptr(), nsl, g.nc, _stream()), "wm_pconv_pack")
            g.img, g.dbias = img, g.bias.to(device).contiguous()
        return g

    def conv_in(self, s, conv: nn.Conv1d, out: Planar):
        from . import _lib as L
        from .ops import _stream
        B, _, T = s.shape
        if T % out.split or (conv.out_channels, B, T // out.split) != (out.C, out.B, out.T):
            raise ValueError("conv_in: the planar output does not match the input")
        L.check(L.load().wm_pconv_in_fwd(s.data_ptr(), conv.weight.detach().contiguous().data_ptr(),
                                         conv.bias.detach().contiguous().data_ptr(), out.ptr(), B, T, conv.out_channels,
                                         conv.kernel_size[0], out.split, out.RP, _stream()), "wm_pconv_in_fwd")

    def to_planar(self, x, out: Planar, phase: int = 0):
        from . import _lib as L
        from .ops import _req, _stream
        x = _req(x, "x")
        B, Cn, T = x.shape
        if (Cn, B, T) != (out.C, out.B, out.T):
            raise ValueError(f"to_planar: {tuple(x.shape)} does not match the planar tensor ({out.B}, {out.C}, {out.T})")
        L.check(L.load().wm_pconv_to_planar(x.data_ptr(), out.ptr(phase), B, Cn, T, out.RP, _stream()),
                "wm_pconv_to_planar")

    def from_planar(self, x: Planar, Tout: int, phase: int = 0):
        from . import _lib as L
        from .ops import _stream
        y = torch.empty(x.B, x.C, Tout, device=x.store.device, dtype=torch.float32)
        L.check(L.load().wm_pconv_from_planar(x.ptr(phase), y.data_ptr(), x.B, x.C, x.T, Tout, x.RP, _stream()),
                "wm_pconv_from_planar")
        return y

    def tail8(self, x: Planar, rb, final: nn.Conv1d, T: int):
        from . import _lib as L
        from .ops import _stream
        y = torch.empty(x.B, 1, T, device=x.store.device, dtype=torch.float32)
        p = [t.detach().contiguous() for t in (rb.conv1.weight, rb.conv1.bias, rb.conv2.weight, rb.conv2.bias,
                                               final.weight, final.bias)]
        L.check(L.load().wm_m14_tail8_fwd(x.ptr(), x.RP, x.B, x.T, *[t.data_ptr() for t in p], y.data_ptr(), T, _stream()),
                "wm_m14_tail8_fwd")
        return y

    def run(self, g: Gemm, srcs: Sequence[Tuple[Planar, int]], B, T, elu, residual, mode, out, out_split=1, ct=None,
            out_T=0, cout=0, g2: Optional[Gemm] = None, skip=None):
        """One wm_pconv_fwd call; with `g2` the fused residual block: g = conv1 (u stays in shared memory), g2 = conv2
        (+ the skip source `skip` = (planar, phase) when g2 carries its slices), residual / elu / out apply to g2."""
        from . import _lib as L
        from .ops import _stream
        dev = srcs[0][0].store.device
        self.pack(g, dev)
        d = _Desc()
        d.nsrc = len(g.srcs)
        for i, (sid, cin, off, taps) in enumerate(g.srcs):
            buf, ph = srcs[i]
            if buf.C != cin or buf.B != B or buf.T != T:
                raise ValueError(f"pconv: source {i} is ({buf.C}, {buf.B}, {buf.T}), the GEMM expects ({cin}, {B}, {T})")
            d.src[i].base, d.src[i].cin, d.src[i].row_off, d.src[i].taps = buf.ptr(ph), cin, off, taps
        d.B, d.T, d.plane_rows = B, T, srcs[0][0].RP
        d.w, d.bias, d.n_total, d.nc = g.img.data_ptr(), g.dbias.data_ptr(), g.n_total, g.nc
        for j, o in enumerate(g.chunk_off):
            d.chunk_off[j] = o
        d.elu, d.mode = int(bool(elu)), mode
        d.residual = residual.ptr() if residual is not None else None
        d.out_split = out_split
        if mode == OUT_FP32:
            d.y, d.ct_cout, d.out_T = out.data_ptr(), cout, out_T
        else:
            d.y, d.out_plane_rows, d.out_phase_rows = out.ptr(), out.RP, out.phase_rows
            if mode == OUT_CONVT:
                d.ct_stride, d.ct_pad, d.ct_cout, d.out_T, d.ct_interleave = ct[0], ct[1], ct[2], out_T, g.interleave
        if g2 is not None:
            self.pack(g2, dev)
            d.fused, d.w2, d.bias2 = 1, g2.img.data_ptr(), g2.dbias.data_ptr()
            if skip is not None:
                d.skip.base, d.skip.cin, d.skip.row_off, d.skip.taps = skip[0].ptr(skip[1]), skip[0].C, 0, 1
        L.check(L.load().wm_pconv_fwd(C.byref(d), _stream()), "wm_pconv_fwd")


_BACKEND = CudaBackend()


def fusable(g1: Gemm, g2: Gemm) -> bool:
    """conv1 + conv2 of a ResidualBlock run as ONE kernel (conv1's output stays in shared memory) when the block has at
    most 64 output channels in one chunk — the layers at T = 2000..16000, which a round trip of u makes HBM-bound."""
    return g1.n_total == g1.nc == g2.n_total == g2.nc and g1.nc <= 64 and g1.cout == g1.n_total and g2.srcs[0][1] == g1.n_total


# ---------------------------------------------------------------------------------------------------------------
# layer walk
# ---------------------------------------------------------------------------------------------------------------
def _versions(mod: nn.Module):
    return tuple((p.data_ptr(), p._version) for p in mod.parameters())


def _plan(mod: nn.Module) -> dict:
    """GEMM descriptions of a Generator / Detector, rebuilt when a parameter changes."""
    key = _versions(mod)
    plan = mod.__dict__.get("_pconv_plan")
    if plan is not None and plan["key"] == key:
        return plan
    plan = {"key": key, "enc": [], "dec": []}
    for blk in mod.encoder_blocks:
        s = blk.conv1.stride[0]
        plan["enc"].append((s, gemm_conv_strided(blk.conv1.weight, blk.conv1.bias, s),
                            gemm_conv2_skip(blk.conv2.weight, blk.conv2.bias, blk.skip_conv.weight, blk.skip_conv.bias)))
    dec = mod.decoder_blocks if hasattr(mod, "decoder_blocks") else mod.upsample_blocks
    for blk in dec:
        if isinstance(blk, nn.ConvTranspose1d):
            if blk.in_channels % 16 or blk.out_channels % 8:
                break
            plan["dec"].append(("ct", blk, gemm_convT(blk.weight, blk.bias, blk.stride[0], blk.padding[0])))
        else:
            if blk.conv1.in_channels % 16 or blk.downsample:
                break
            plan["dec"].append(("rb", blk, gemm_conv_s1(blk.conv1.weight, blk.conv1.bias),
                                gemm_conv_s1(blk.conv2.weight, blk.conv2.bias)))
    plan["dec_rest"] = list(dec)[len(plan["dec"]):]
    if hasattr(mod, "final_conv") and not plan["dec_rest"] and mod.final_conv.in_channels % 16 == 0:
        plan["final"] = gemm_conv_s1(mod.final_conv.weight, mod.final_conv.bias)
    mod.__dict__["_pconv_plan"] = plan
    return plan


def supported(mod: nn.Module, T: int) -> bool:
    """The tensor-core walk covers the reference's layer pattern (py/main14b_2.py:107-224) when every encoder block
    down-samples with a k3 convolution over a multiple of 16 channels, the strides divide T, and transposed
    convolutions have kernel_size = 2 * stride."""
    try:
        if mod.init_conv.in_channels != 1 or mod.init_conv.out_channels % 16 or mod.init_conv.kernel_size[0] > 7:
            return False
        if mod.init_conv.stride[0] != 1 or mod.init_conv.padding[0] != mod.init_conv.kernel_size[0] // 2:
            return False
        t = T
        for blk in mod.encoder_blocks:
            s = blk.conv1.stride[0]
            if not blk.downsample or s < 2 or s > 8 or t % s or blk.conv1.in_channels % 16 or blk.conv1.out_channels % 16:
                return False
            if blk.conv1.kernel_size[0] != 3 or blk.conv1.padding[0] != 1:
                return False
            t //= s
        dec = mod.decoder_blocks if hasattr(mod, "decoder_blocks") else mod.upsample_blocks
        for blk in dec:
            if isinstance(blk, nn.ConvTranspose1d):
                s, p = blk.stride[0], blk.padding[0]
                if blk.kernel_size[0] != 2 * s or not 0 <= p < s or blk.output_padding[0] != 0 or s - 2 * p > 1 or s < 2 * p:
                    return False
                t = (t - 1) * s - 2 * p + 2 * s
        return t >= T and T > 0
    except AttributeError:
        return False


def _encoder(mod, s, plan, last_fp32: bool):
    """init_conv + encoder_blocks (py/main14b_2.py:152-153, 208-209).  Returns the last block's output: planar, or
    fp32 channels-first when `last_fp32` (the Generator's projection reads it through the fp32 operators)."""
    be = _BACKEND
    B, _, T = s.shape
    dev = s.device
    strides = [e[0] for e in plan["enc"]]
    cur = be.planar(mod.init_conv.out_channels, B, T // strides[0], strides[0], dev)
    be.conv_in(s, mod.init_conv, cur)
    t = T
    for i, (st, g1, g2) in enumerate(plan["enc"]):
        to = t // st
        last = i + 1 == len(strides)
        if fusable(g1, g2) and not (last and last_fp32):
            nsp = 1 if last else strides[i + 1]
            nxt = be.planar(g2.cout, B, to // nsp, nsp, dev)
            be.run(g1, [(cur, st - 1), (cur, 0), (cur, 1)], B, to, True, None, OUT_PLANAR, nxt, out_split=nsp, g2=g2,
                   skip=(cur, 0))
            cur, t = nxt, to
            continue
        u = be.planar(g1.cout, B, to, 1, dev)
        be.run(g1, [(cur, st - 1), (cur, 0), (cur, 1)], B, to, True, None, OUT_PLANAR, u)
        if last and last_fp32:
            y = be.fp32((B, g2.cout, to), dev)
            be.run(g2, [(u, 0), (cur, 0)], B, to, True, None, OUT_FP32, y, out_T=to, cout=g2.cout)
            return y, to
        nsp = 1 if last else strides[i + 1]
        nxt = be.planar(g2.cout, B, to // nsp, nsp, dev)
        be.run(g2, [(u, 0), (cur, 0)], B, to, True, None, OUT_PLANAR, nxt, out_split=nsp)
        cur, t = nxt, to
    return cur, t


def _decoder(plan, cur, t):
    """ConvTranspose1d + ResidualBlock pairs (py/main14b_2.py:140-147, 196-203) while they fit the tensor-core path."""
    be = _BACKEND
    B, dev = cur.B, cur.store.device if hasattr(cur, "store") else None
    for item in plan["dec"]:
        if item[0] == "ct":
            _, blk, g = item
            s, p = blk.stride[0], blk.padding[0]
            to = (t - 1) * s - 2 * p + 2 * s
            y = be.planar(g.cout, B, to, 1, dev)
            be.run(g, [(cur, 0)], B, t, False, None, OUT_CONVT, y, ct=(s, p, g.cout), out_T=to)
            cur, t = y, to
        else:
            _, blk, g1, g2 = item
            y = be.planar(g2.cout, B, t, 1, dev)
            if fusable(g1, g2):
                be.run(g1, [(cur, 0)], B, t, True, cur, OUT_PLANAR, y, g2=g2)
            else:
                u = be.planar(g1.cout, B, t, 1, dev)
                be.run(g1, [(cur, 0)], B, t, True, None, OUT_PLANAR, u)
                be.run(g2, [(u, 0)], B, t, True, cur, OUT_PLANAR, y)
            cur = y
    return cur, t


def _nvtx(name):
    from .ops import nvtx
    return nvtx(name)


@_nvtx("main14b_2.detector")
def detector_forward(mod, x):
    """Detector.forward (py/main14b_2.py:207-224): logits (B, 1 + bits, T)."""
    from . import main14b_2 as M
    be = _BACKEND
    plan = _plan(mod)
    B, _, T = x.shape
    cur, t = _encoder(mod, x, plan, last_fp32=False)
    cur, t = _decoder(plan, cur, t)
    if "final" in plan:
        g = plan["final"]
        y = be.fp32((B, g.cout, T), x.device)
        be.run(g, [(cur, 0)], B, t, False, None, OUT_FP32, y, out_T=T, cout=g.cout)
        return y
    h = be.from_planar(cur, t)
    for blk in plan["dec_rest"]:
        h = M.conv_transpose1d(h, blk) if isinstance(blk, nn.ConvTranspose1d) else blk(h)
    return M._fit_length(M.conv1d(h, mod.final_conv), T)


def _tail8_fits(rest, final: nn.Conv1d, cur) -> bool:
    if len(rest) != 1 or isinstance(rest[0], nn.ConvTranspose1d) or cur.C != 8:
        return False
    rb = rest[0]
    return (not rb.downsample and rb.conv1.in_channels == 8 and rb.conv1.kernel_size[0] == 3 and rb.conv2.kernel_size[0] == 3
            and (final.in_channels, final.out_channels, final.kernel_size[0], final.padding[0], final.stride[0]) == (8, 1, 7, 3, 1))


@_nvtx("main14b_2.generator")
def generator_forward(mod, s, message=None):
    """Generator.forward (py/main14b_2.py:150-182): delta (B, 1, T)."""
    from . import main14b_2 as M
    be = _BACKEND
    plan = _plan(mod)
    B, _, T = s.shape
    x, t = _encoder(mod, s, plan, last_fp32=True)
    e = mod.E.weight.detach()[message.to(torch.int64)] if message is not None else None
    x = M.conv1d(x, M._LinearAsConv(mod.proj), chan_add=e)
    x = M.lstm_small(x, mod.lstm)
    x = M.conv1d(x, mod.final_conv_enc)
    if plan["dec"]:
        cur = be.planar(x.shape[1], B, t, 1, s.device)
        be.to_planar(x, cur)
        cur, t = _decoder(plan, cur, t)
        if _tail8_fits(plan["dec_rest"], mod.final_conv_dec, cur):
            # ResidualBlock(8, 8) + final_conv_dec + crop in one kernel on the planar tensor (py/main14b_2.py:147-149,173-177)
            return be.tail8(cur, plan["dec_rest"][0], mod.final_conv_dec, T)
        x = be.from_planar(cur, t)
    for blk in plan["dec_rest"]:
        x = M.conv_transpose1d(x, blk) if isinstance(blk, nn.ConvTranspose1d) else blk(x)
    return M._fit_length(M.conv1d(x, mod.final_conv_dec), T)
