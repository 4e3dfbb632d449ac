"""main14b_2 (py/main14b_2.py:83-224, BASELINE config 3): the oracle against the reference's own outputs, the
drop-in modules' seeded parameters against the reference's, and the CUDA path against both."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import wm_oracle_14b2 as O
from tests import helpers as H
from wmb200 import main14b_2 as M

IO = H.load_npz("main14b2_io.npz")
DEV = "cuda"


def build(kind):
    torch.manual_seed(int(IO["seed_g" if kind == "g" else "seed_d"]))
    return (M.Generator() if kind == "g" else M.Detector()).eval()


def maxerr(a, b):
    a = a.detach().float().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(a)
    b = b.detach().float().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max())


def test_seeded_construction_reproduces_the_reference_parameters():
    """Same layer types in the same construction order -> the same draws: every tensor's checksum matches the
    reference model built under the same seed, and the key set is the reference's state dict."""
    for kind, pre in (("g", "gsum/"), ("d", "dsum/")):
        sd = build(kind).state_dict()
        want = {k[len(pre):]: v for k, v in IO.items() if k.startswith(pre)}
        assert set(sd) == set(want)
        for k, v in sd.items():
            got = np.array([float(v.double().sum()), float(v.double().abs().sum())])
            assert np.allclose(got, want[k], rtol=0, atol=1e-9 * max(1.0, float(want[k][1]))), k


def test_oracle_matches_reference_outputs():
    gsd, dsd = build("g").state_dict(), build("d").state_dict()
    s, msg = torch.from_numpy(IO["s"]), torch.from_numpy(IO["messages"])
    with torch.no_grad():
        assert maxerr(O.generator_forward(gsd, s, msg), IO["delta"]) < 1e-6
        assert maxerr(O.generator_forward(gsd, s[:1]), IO["delta_nomsg"]) < 1e-6
        short = torch.from_numpy(IO["s_short"])
        assert maxerr(O.generator_forward(gsd, short, msg[:2]), IO["delta_short"]) < 1e-6
        assert maxerr(O.detector_forward(dsd, s[:2]), IO["logits"]) < 1e-5
        assert maxerr(O.detector_forward(dsd, short[:1]), IO["logits_short"]) < 1e-5


def test_modules_refuse_cpu_tensors():
    with pytest.raises(RuntimeError):
        build("d")(torch.zeros(1, 1, 16000))


@pytest.mark.gpu
def test_generator_matches_reference_goldens():
    G = build("g").to(DEV)
    s, msg = torch.from_numpy(IO["s"]).to(DEV), torch.from_numpy(IO["messages"]).to(DEV)
    d = G(s, msg)
    assert d.shape == (3, 1, 16000)
    scale = float(np.abs(IO["delta"]).max())
    assert maxerr(d, IO["delta"]) < 1e-4 * max(scale, 1.0)
    assert maxerr(G(s[:1]), IO["delta_nomsg"]) < 1e-4 * max(scale, 1.0)
    short = torch.from_numpy(IO["s_short"]).to(DEV)
    assert maxerr(G(short, msg[:2]), IO["delta_short"]) < 1e-4 * max(scale, 1.0)     # 5003 -> crop / pad branch


@pytest.mark.gpu
def test_detector_matches_reference_goldens():
    D = build("d").to(DEV)
    s = torch.from_numpy(IO["s"]).to(DEV)
    lg = D(s[:2])
    assert lg.shape == (2, 17, 16000)
    assert maxerr(lg, IO["logits"]) < 2e-4
    assert maxerr(torch.sigmoid(lg[:, 0]), torch.sigmoid(torch.from_numpy(IO["logits"][:, 0]))) < 1e-3
    assert maxerr(D(torch.from_numpy(IO["s_short"][:1]).to(DEV)), IO["logits_short"]) < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("Cin,Cout,K,stride,pad,T", [(1, 32, 7, 1, 3, 300), (32, 64, 3, 2, 1, 301), (64, 128, 3, 4, 1, 1000),
                                                     (128, 256, 3, 5, 1, 77), (256, 512, 3, 8, 1, 400), (32, 64, 1, 2, 0, 129),
                                                     (8, 1, 7, 1, 3, 500), (5, 70, 4, 3, 2, 200),
                                                     # the three tile shapes (64/32/16 channels), K > 8, output lengths
                                                     # that are / are not multiples of 4 (vector and scalar epilogues)
                                                     (32, 17, 7, 1, 3, 1000), (16, 24, 3, 1, 1, 403), (4, 33, 16, 2, 7, 500),
                                                     (3, 9, 16, 1, 8, 260), (20, 20, 9, 1, 4, 257), (64, 64, 3, 1, 1, 128)])
def test_generic_conv1d_vs_torch(Cin, Cout, K, stride, pad, T):
    g = torch.Generator().manual_seed(Cin * 131 + Cout)
    conv = torch.nn.Conv1d(Cin, Cout, K, stride=stride, padding=pad)
    x = torch.randn(3, Cin, T, generator=g)
    ref = conv(x).detach()
    res, add = torch.randn(ref.shape, generator=g), torch.randn(3, Cout, generator=g)
    convd = conv.to(DEV)
    tol = 2e-5 * (Cin * K) ** 0.5
    assert maxerr(M.conv1d(x.to(DEV), convd), ref) < tol
    assert maxerr(M.conv1d(x.to(DEV), convd, act=True, residual=res.to(DEV), chan_add=add.to(DEV)),
                  F.elu(ref + res + add[:, :, None])) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("Cin,Cout,stride,T", [(128, 64, 8, 50), (64, 32, 5, 400), (32, 16, 4, 2001), (16, 8, 2, 333),
                                               (512, 256, 8, 7), (6, 70, 3, 41)])
def test_generic_convtranspose1d_vs_torch(Cin, Cout, stride, T):
    g = torch.Generator().manual_seed(Cin + stride)
    ct = torch.nn.ConvTranspose1d(Cin, Cout, kernel_size=2 * stride, stride=stride, padding=stride // 2)
    x = torch.randn(2, Cin, T, generator=g)
    ref = ct(x).detach()
    ctd = ct.to(DEV)
    for direct in (False, True):             # per-phase convolution form and the gather kernel
        got = M.conv_transpose1d(x.to(DEV), ctd, direct=direct)
        assert got.shape == ref.shape
        assert maxerr(got, ref) < 2e-5 * (2 * Cin) ** 0.5, direct


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,T,nl", [(3, 32, 50, 2), (1, 32, 1, 2), (5, 16, 33, 1), (2, 64, 20, 3)])
def test_small_lstm_vs_torch(B, H, T, nl):
    g = torch.Generator().manual_seed(H + T)
    lstm = torch.nn.LSTM(H, H, num_layers=nl, batch_first=True)
    x = torch.randn(B, H, T, generator=g)
    ref = lstm(x.transpose(1, 2))[0].transpose(1, 2).detach()
    assert maxerr(M.lstm_small(x.to(DEV), lstm.to(DEV)), ref) < 1e-5
