"""Host side of the tensor-core main14b_2 path (wmb200/pconv.py) on the CPU: every reference layer's GEMM description
(sources, row offsets, taps, weights, phase-split / transposed output mapping) executed through tests/pconv_emu.py's
emulation of wm_pconv_fwd's contract and compared with torch's own convolutions, then the whole Generator / Detector
walk against the oracle (py/main14b_2.py:86-224)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import wm_oracle_14b2 as O
from tests.pconv_emu import EmuBackend, EmuPlanar
from wmb200 import main14b_2 as M
from wmb200 import pconv as PC

GAP = PC.GAP


def planar_from(x, split=1):
    """(B, C, T) -> EmuPlanar through the emulated wm_pconv_to_planar, optionally re-split by phase"""
    B, Cn, T = x.shape
    be = EmuBackend()
    if split == 1:
        p = EmuPlanar(Cn, B, T)
        be.to_planar(x, p)
        return p
    p = EmuPlanar(Cn, B, T // split, split)
    for ph in range(split):
        q = EmuPlanar(Cn, B, T // split)
        be.to_planar(x[:, :, ph::split].contiguous(), q)
        p.store[ph] = q.store[0]
    return p


def dense(p, T=None):
    return EmuBackend().from_planar(p, p.T if T is None else T)


@pytest.mark.parametrize("cin,cout,K", [(16, 16, 3), (32, 64, 3), (32, 17, 7), (64, 256, 3), (48, 8, 3)])
def test_conv_stride1(cin, cout, K):
    torch.manual_seed(cin + cout)
    B, T = 3, 37
    x, w, b = torch.randn(B, cin, T), torch.randn(cout, cin, K), torch.randn(cout)
    g = PC.gemm_conv_s1(w, b)
    res = torch.randn(B, g.n_total, T)
    want = F.elu(F.conv1d(x, w, b, padding=K // 2) + res[:, :cout])
    be = EmuBackend()
    out = EmuPlanar(g.n_total, B, T)
    be.run(g, [(planar_from(x), 0)], B, T, True, planar_from(res), PC.OUT_PLANAR, out)
    assert torch.allclose(dense(out)[:, :cout], want, atol=1e-4)
    Tp = T + GAP
    for c in range(B + 1):                                   # gap rows are rewritten with zeros
        assert (out.store[0][:, c * Tp: c * Tp + GAP] == 0).all()
    y = be.fp32((B, cout, T - 5), None)
    be.run(g, [(planar_from(x), 0)], B, T, False, None, PC.OUT_FP32, y, out_T=T - 5, cout=cout)
    assert torch.allclose(y.float(), F.conv1d(x, w, b, padding=K // 2)[:, :, :T - 5], atol=1e-4)


@pytest.mark.parametrize("s", [2, 4, 5, 8])
def test_strided_block_with_folded_skip_and_split_output(s):
    torch.manual_seed(s)
    B, cin, cout, nsp = 2, 32, 64, 5
    T = s * nsp * 6
    blk = M.ResidualBlock(cin, cout, stride=s)
    x = torch.randn(B, cin, T)
    want = O.residual_block(x, {"b." + k: v for k, v in blk.state_dict().items()}, "b", s)
    with torch.no_grad():
        u_ref = F.elu(blk.conv1(x))
        y_ref = F.elu(blk.conv2(u_ref) + blk.skip_conv(x))
    assert torch.allclose(want, y_ref, atol=1e-5)
    be = EmuBackend()
    xin = planar_from(x, s)
    g1 = PC.gemm_conv_strided(blk.conv1.weight, blk.conv1.bias, s)
    g2 = PC.gemm_conv2_skip(blk.conv2.weight, blk.conv2.bias, blk.skip_conv.weight, blk.skip_conv.bias)
    To = T // s
    u = EmuPlanar(cout, B, To)
    be.run(g1, [(xin, s - 1), (xin, 0), (xin, 1)], B, To, True, None, PC.OUT_PLANAR, u)
    assert torch.allclose(dense(u), u_ref, atol=1e-4)
    y = EmuPlanar(cout, B, To // nsp, nsp)
    be.run(g2, [(u, 0), (xin, 0)], B, To, True, None, PC.OUT_PLANAR, y, out_split=nsp)
    for ph in (0, 1, nsp - 1):                               # the phases a k3 stride-nsp convolution reads
        q = EmuPlanar(cout, B, To // nsp)
        q.store[0] = y.store[ph]
        assert torch.allclose(dense(q), y_ref[:, :, ph::nsp], atol=1e-4)
        Tp = To // nsp + GAP
        for c in range(B + 1):
            assert (y.store[ph][:, c * Tp: c * Tp + GAP] == 0).all()


@pytest.mark.parametrize("cin,cout,s", [(32, 16, 2), (16, 8, 2), (64, 32, 4), (32, 16, 4), (128, 128, 5), (64, 32, 5),
                                        (32, 256, 8), (128, 64, 8)])
def test_conv_transpose(cin, cout, s):
    torch.manual_seed(cin + s)
    B, T, p = 3, 11, s // 2
    ct = nn.ConvTranspose1d(cin, cout, 2 * s, stride=s, padding=p)
    x = torch.randn(B, cin, T)
    with torch.no_grad():
        want = ct(x)
    g = PC.gemm_convT(ct.weight, ct.bias, s, p)
    To = want.shape[-1]
    assert To == s * T + s - 2 * p
    out = EmuPlanar(cout, B, To)
    EmuBackend().run(g, [(planar_from(x), 0)], B, T, False, None, PC.OUT_CONVT, out, ct=(s, p, cout), out_T=To)
    assert torch.allclose(dense(out), want, atol=1e-4)
    Tp = To + GAP
    for c in range(B + 1):
        assert (out.store[0][:, c * Tp: c * Tp + GAP] == 0).all()
    assert not torch.isnan(out.store[0][:, :B * Tp + GAP]).any()


def _cpu_ops(monkeypatch):
    """the fp32 CUDA operators of main14b_2.py the walk still calls, replaced by torch on the CPU"""
    def conv1d(x, conv, act=False, residual=None, chan_add=None):
        w = conv.weight.detach()
        w = w[:, :, None] if w.dim() == 2 else w
        y = F.conv1d(x.float(), w, conv.bias.detach(), stride=conv.stride[0], padding=conv.padding[0])
        if chan_add is not None:
            y = y + chan_add[:, :, None]
        if residual is not None:
            y = y + residual
        return F.elu(y) if act else y

    def lstm_small(x, lstm):
        with torch.no_grad():
            return lstm(x.float().transpose(1, 2))[0].transpose(1, 2)

    def conv_transpose1d(x, ct, direct=False):
        with torch.no_grad():
            return ct(x.float())

    monkeypatch.setattr(M, "conv1d", conv1d)
    monkeypatch.setattr(M, "lstm_small", lstm_small)
    monkeypatch.setattr(M, "conv_transpose1d", conv_transpose1d)


@pytest.mark.parametrize("B,T", [(2, 640), (1, 960)])
def test_whole_models_against_the_oracle(monkeypatch, B, T):
    _cpu_ops(monkeypatch)
    be = EmuBackend()
    monkeypatch.setattr(PC, "_BACKEND", be)
    torch.manual_seed(5)
    G, D = M.Generator().eval(), M.Detector().eval()
    s = 0.3 * torch.randn(B, 1, T)
    msg = torch.randint(0, 65536, (B,))
    assert PC.supported(G, T) and PC.supported(D, T) and not PC.supported(D, T + 1)
    with torch.no_grad():
        want_d = O.generator_forward(G.state_dict(), s, msg)
        want_l = O.detector_forward(D.state_dict(), s)
        monkeypatch.setattr(M.ResidualBlock, "forward",
                            lambda self, x: F.elu(self.conv2(F.elu(self.conv1(x))) + (self.skip_conv(x) if self.downsample else x)))
        got_d = PC.generator_forward(G, s, msg)
        n_g = len(be.calls)
        got_l = PC.detector_forward(D, s)
    assert got_d.shape == want_d.shape and got_l.shape == want_l.shape
    assert float((got_d.float() - want_d).abs().max()) < 1e-5 * max(1.0, float(want_d.abs().max()))
    assert float((got_l.float() - want_l).abs().max()) < 1e-4
    # generator: 8 encoder GEMMs + 4 transposed + 3 residual blocks on the tensor-core path; detector: 8 + 4 + 8 + final
    assert n_g == 8 + 4 + 6 and len(be.calls) - n_g == 8 + 4 + 8 + 1
    # blocks of <= 64 channels are fused: encoder block 1 (both models), generator RB64/32/16, detector RB64/32
    assert sum(1 for c in be.calls if c[-1] == "fused") == 2 + 3 + 2
    assert sum(1 for c in be.calls if c[0] == PC.OUT_CONVT and len(c[3]) == 1 and c[3][0][2] == 2) >= 5   # 2-tap form used


def test_descriptor_binding_matches_the_header():
    import ctypes
    from wmb200 import _lib as L
    lib = L.load()
    assert ctypes.sizeof(PC._Desc) == lib.wm_pconv_desc_bytes()
    assert lib.wm_pconv_plane_rows(7, 123) == PC.plane_rows(7, 123)
