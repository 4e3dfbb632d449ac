"""wm_pconv_fwd (csrc/wm_pconv_tc.cu) on the B200 against torch's fp32 convolutions (TF32 off): every layer type of
py/main14b_2.py:86-224 at every channel count / stride / chunk width the model uses, clip boundaries inside tiles,
phase-split and transposed outputs, and the whole models at the reference's size."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from wmb200 import main14b_2 as M
from wmb200 import ops
from wmb200 import pconv as PC

pytestmark = pytest.mark.gpu
DEV = "cuda"
GAP = PC.GAP


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


BE = PC.CudaBackend()


def planar_from(x, split=1):
    B, Cn, T = x.shape
    p = PC.Planar(Cn, B, T // split, split, x.device)
    p.store.fill_(0xFF)                       # NaN patterns wherever nothing is written
    for ph in range(split):
        BE.to_planar(x[:, :, ph::split].contiguous(), p, ph)
    return p


def fresh(Cn, B, T, split=1):
    p = PC.Planar(Cn, B, T, split, DEV)
    p.store.fill_(0xFF)
    return p


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def gap_rows_zero(p, phases=(0,)):
    # the 128 guard bytes in front of the first plane and behind the last one still hold the fill pattern
    # (compute-sanitizer is closed on this pool: this is the out-of-bounds-write check the tests can make themselves)
    assert int(p.store[:128].min()) == 0xFF and int(p.store[-128:].min()) == 0xFF, "write outside the planar tensor"
    rows_used = p.B * (p.T + GAP) + GAP
    for ph in phases:
        raw = p.store[128 + ph * p.phase_rows * 16: 128 + (ph + 1) * p.phase_rows * 16].view(2 * (p.C // 8), p.RP, 16)
        assert int(raw[:, rows_used:].min()) == 0xFF, "write behind the last row of a plane"
    Tp = p.T + GAP
    for ph in phases:
        raw = p.store[128 + ph * p.phase_rows * 16: 128 + (ph + 1) * p.phase_rows * 16].view(2 * (p.C // 8), p.RP, 16)
        for c in range(p.B + 1):
            if int(raw[:, c * Tp: c * Tp + GAP].max()) != 0:
                return False
    return True


@pytest.mark.parametrize("cin,cout,K,B,T", [(16, 16, 3, 3, 37), (32, 32, 3, 2, 300), (64, 64, 3, 5, 129), (128, 128, 3, 2, 257),
                                            (256, 256, 3, 3, 50), (512, 512, 3, 4, 50), (32, 17, 7, 2, 500), (16, 8, 3, 2, 100),
                                            (64, 32, 3, 1, 1000)])
def test_conv_stride1(cin, cout, K, B, T):
    torch.manual_seed(cin + cout)
    x = torch.randn(B, cin, T, device=DEV)
    w = torch.randn(cout, cin, K, device=DEV) / (cin * K) ** 0.5
    b = torch.randn(cout, device=DEV)
    g = PC.gemm_conv_s1(w, b)
    res = torch.randn(B, g.n_total, T, device=DEV)
    want = F.elu(F.conv1d(x, w, b, padding=K // 2) + res[:, :cout])
    out = fresh(g.n_total, B, T)
    BE.run(g, [(planar_from(x), 0)], B, T, True, planar_from(res), PC.OUT_PLANAR, out)
    got = BE.from_planar(out, T)[:, :cout]
    assert relerr(got, want) < 2e-5, relerr(got, want)
    assert gap_rows_zero(out)
    y = torch.full((B, cout, T - 5), float("nan"), device=DEV)
    BE.run(g, [(planar_from(x), 0)], B, T, False, None, PC.OUT_FP32, y, out_T=T - 5, cout=cout)
    want2 = F.conv1d(x, w, b, padding=K // 2)[:, :, :T - 5]
    assert relerr(y, want2) < 2e-5, relerr(y, want2)


@pytest.mark.parametrize("s,cin,cout,nsp", [(2, 32, 64, 4), (4, 64, 128, 5), (5, 128, 256, 8), (8, 256, 512, 1), (2, 16, 32, 2)])
def test_strided_block_with_folded_skip_and_split_output(s, cin, cout, nsp):
    torch.manual_seed(s)
    B = 3
    T = s * nsp * 26
    blk = M.ResidualBlock(cin, cout, stride=s).to(DEV)
    x = torch.randn(B, cin, T, device=DEV)
    with torch.no_grad():
        u_ref = F.elu(blk.conv1(x))
        y_ref = F.elu(blk.conv2(u_ref) + blk.skip_conv(x))
    xin = planar_from(x, s)
    g1 = PC.gemm_conv_strided(blk.conv1.weight, blk.conv1.bias, s)
    g2 = PC.gemm_conv2_skip(blk.conv2.weight, blk.conv2.bias, blk.skip_conv.weight, blk.skip_conv.bias)
    To = T // s
    u = fresh(cout, B, To)
    BE.run(g1, [(xin, s - 1), (xin, 0), (xin, 1)], B, To, True, None, PC.OUT_PLANAR, u)
    assert relerr(BE.from_planar(u, To), u_ref) < 2e-5
    y = fresh(cout, B, To // nsp, nsp)
    BE.run(g2, [(u, 0), (xin, 0)], B, To, True, None, PC.OUT_PLANAR, y, out_split=nsp)
    phases = sorted({0, 1 % nsp, nsp - 1})
    for ph in phases:
        got = BE.from_planar(y, To // nsp, ph)
        assert relerr(got, y_ref[:, :, ph::nsp]) < 2e-5, (ph, relerr(got, y_ref[:, :, ph::nsp]))
    assert gap_rows_zero(y, phases)


@pytest.mark.parametrize("cin,cout,s,B,T", [(512, 256, 8, 3, 50), (256, 128, 5, 2, 400), (128, 64, 4, 2, 301), (64, 32, 2, 2, 500),
                                            (128, 64, 8, 3, 50), (64, 32, 5, 2, 77), (32, 16, 4, 3, 130), (16, 8, 2, 2, 260)])
def test_conv_transpose(cin, cout, s, B, T):
    torch.manual_seed(cin + s)
    p = s // 2
    ct = nn.ConvTranspose1d(cin, cout, 2 * s, stride=s, padding=p).to(DEV)
    x = torch.randn(B, cin, T, device=DEV)
    with torch.no_grad():
        want = ct(x)
    g = PC.gemm_convT(ct.weight, ct.bias, s, p)
    To = want.shape[-1]
    out = fresh(cout, B, To)
    BE.run(g, [(planar_from(x), 0)], B, T, False, None, PC.OUT_CONVT, out, ct=(s, p, cout), out_T=To)
    got = BE.from_planar(out, To)
    assert relerr(got, want) < 2e-5, relerr(got, want)
    assert gap_rows_zero(out)


@pytest.mark.parametrize("split", [1, 2])
def test_input_convolution(split):
    torch.manual_seed(1)
    B, T = 3, 640
    conv = nn.Conv1d(1, 32, 7, padding=3).to(DEV)
    s = torch.randn(B, 1, T, device=DEV)
    out = fresh(32, B, T // split, split)
    BE.conv_in(s, conv, out)
    with torch.no_grad():
        want = conv(s)
    for ph in range(split):
        assert relerr(BE.from_planar(out, T // split, ph), want[:, :, ph::split]) < 2e-5      # bf16 pair: 2^-17
    assert gap_rows_zero(out, range(split))


def test_whole_models_match_the_fp32_operators():
    """the tensor-core walk against the layer-by-layer fp32 CUDA operators (the exact-order path), 1 s clips"""
    torch.manual_seed(3)
    G, D = M.Generator().to(DEV).eval(), M.Detector().to(DEV).eval()
    s = (0.1 * torch.randn(5, 1, 16000, device=DEV)).clamp(-0.99, 0.99)
    msg = torch.randint(0, 65536, (5,), device=DEV)
    assert PC.supported(G, 16000) and PC.supported(D, 16000)
    n0 = ops.launch_count()
    d_tc, l_tc = G(s, msg), D(s)
    n_tc = ops.launch_count() - n0
    old = ops.set_math_mode(0)
    try:
        d_32, l_32 = G(s, msg), D(s)
    finally:
        ops.set_math_mode(old)
    assert d_tc.shape == d_32.shape == (5, 1, 16000) and l_tc.shape == l_32.shape == (5, 17, 16000)
    assert relerr(d_tc, d_32) < 1e-4, relerr(d_tc, d_32)
    assert float((l_tc - l_32).abs().max()) < 2e-4, float((l_tc - l_32).abs().max())
    assert n_tc < 120                       # incl. one weight pack per layer on first use


def test_generator_tail_kernel():
    """ResidualBlock(8, 8) + Conv1d(8, 1, 7) + crop (wm_m14_tail8_fwd) against torch, clip edges and a ragged last block"""
    torch.manual_seed(11)
    B, Tx, T = 3, 1000, 993
    rb, fin = M.ResidualBlock(8, 8).to(DEV), nn.Conv1d(8, 1, 7, padding=3).to(DEV)
    x = torch.randn(B, 8, Tx, device=DEV)
    with torch.no_grad():
        want = fin(F.elu(rb.conv2(F.elu(rb.conv1(x))) + x))[:, :, :T]
    got = BE.tail8(planar_from(x), rb, fin, T)
    assert got.shape == want.shape and relerr(got, want) < 2e-5, relerr(got, want)


@pytest.mark.parametrize("B,T,L", [(3, 50, 2), (2, 130, 1)])
def test_register_lstm(B, T, L):
    torch.manual_seed(T)
    lstm = nn.LSTM(32, 32, num_layers=L, batch_first=True).to(DEV)
    x = torch.randn(B, 32, T, device=DEV)
    with torch.no_grad():
        want = lstm(x.transpose(1, 2))[0].transpose(1, 2)
    assert relerr(M.lstm_small(x, lstm), want) < 5e-5      # cuDNN vs expf/tanhf round-off over 130 steps


@pytest.mark.parametrize("C,B,T", [(16, 3, 300), (32, 2, 1000), (64, 3, 517), (64, 1, 126), (32, 5, 125)])
def test_fused_residual_block(C, B, T):
    """conv1 -> shared memory -> conv2 + identity residual in one kernel, tiles of 126 rows straddling clips"""
    torch.manual_seed(C + T)
    blk = M.ResidualBlock(C, C).to(DEV)
    x = torch.randn(B, C, T, device=DEV)
    with torch.no_grad():
        want = F.elu(blk.conv2(F.elu(blk.conv1(x))) + x)
    g1, g2 = PC.gemm_conv_s1(blk.conv1.weight, blk.conv1.bias), PC.gemm_conv_s1(blk.conv2.weight, blk.conv2.bias)
    assert PC.fusable(g1, g2)
    xin, out = planar_from(x), fresh(C, B, T)
    BE.run(g1, [(xin, 0)], B, T, True, xin, PC.OUT_PLANAR, out, g2=g2)
    got = BE.from_planar(out, T)
    assert relerr(got, want) < 3e-5, relerr(got, want)
    assert gap_rows_zero(out)


@pytest.mark.parametrize("s,cin,cout,nsp", [(2, 32, 64, 4), (4, 16, 32, 1), (5, 32, 64, 5)])
def test_fused_strided_block(s, cin, cout, nsp):
    """strided conv1 from phase buffers -> shared memory -> conv2 + folded skip, phase-split output"""
    torch.manual_seed(s + cout)
    B = 3
    T = s * nsp * 53
    blk = M.ResidualBlock(cin, cout, stride=s).to(DEV)
    x = torch.randn(B, cin, T, device=DEV)
    with torch.no_grad():
        y_ref = F.elu(blk.conv2(F.elu(blk.conv1(x))) + blk.skip_conv(x))
    xin = planar_from(x, s)
    g1 = PC.gemm_conv_strided(blk.conv1.weight, blk.conv1.bias, s)
    g2 = PC.gemm_conv2_skip(blk.conv2.weight, blk.conv2.bias, blk.skip_conv.weight, blk.skip_conv.bias)
    assert PC.fusable(g1, g2)
    To = T // s
    y = fresh(cout, B, To // nsp, nsp)
    BE.run(g1, [(xin, s - 1), (xin, 0), (xin, 1)], B, To, True, None, PC.OUT_PLANAR, y, out_split=nsp, g2=g2, skip=(xin, 0))
    phases = sorted({0, 1 % nsp, nsp - 1})
    for ph in phases:
        got = BE.from_planar(y, To // nsp, ph)
        assert relerr(got, y_ref[:, :, ph::nsp]) < 3e-5, (ph, relerr(got, y_ref[:, :, ph::nsp]))
    assert gap_rows_zero(y, phases)


def test_rows_are_independent_of_batch_composition():
    """The flattened-row GEMM lets a 128-row tile straddle clips; a clip's result must not depend on its neighbours or on
    where the tile boundaries fall: the same clip alone, in a batch of 37 and at the end of a batch of 300 gives
    bit-identical logits and deltas (size-independent property, checked at a batch whose T = 50 layers span 127 tiles)."""
    torch.manual_seed(21)
    G, D = M.Generator().to(DEV).eval(), M.Detector().to(DEV).eval()
    s = (0.1 * torch.randn(300, 1, 16000, device=DEV)).clamp(-0.99, 0.99)
    msg = torch.randint(0, 65536, (300,), device=DEV)
    d_all, l_all = G(s, msg), D(s)
    for idx in ([0], [299], list(range(100, 137))):
        d_sub, l_sub = G(s[idx], msg[idx]), D(s[idx])
        assert torch.equal(d_sub, d_all[idx]), idx[0]
        assert torch.equal(l_sub, l_all[idx]), idx[0]


def test_graph_replay_matches_the_eager_walk():
    """main14b_2.GraphedEmbedDetect: the captured layer walk replayed on new inputs gives the eager walk's bits"""
    torch.manual_seed(4)
    G, D = M.Generator().to(DEV).eval(), M.Detector().to(DEV).eval()
    ge = M.GraphedEmbedDetect(G, D, 3, 640)
    for seed in (1, 2):
        g = torch.Generator(device=DEV).manual_seed(seed)
        s = 0.1 * torch.randn(3, 1, 640, device=DEV, generator=g)
        msg = torch.randint(0, 65536, (3,), device=DEV, generator=g)
        delta, logits = ge(s, msg)
        want_d = G(s, msg)
        assert torch.equal(delta, want_d) and torch.equal(logits, D(s + want_d))


def test_empty_batch():
    """B = 0 goes through the whole walk without a launch error and returns empty tensors of the right shape"""
    G, D = M.Generator().to(DEV).eval(), M.Detector().to(DEV).eval()
    s = torch.zeros(0, 1, 16000, device=DEV)
    assert G(s, torch.zeros(0, dtype=torch.int64, device=DEV)).shape == (0, 1, 16000)
    assert D(s).shape == (0, 17, 16000)
