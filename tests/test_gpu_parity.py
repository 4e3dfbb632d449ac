"""Parity of the CUDA path (through the C ABI) with the oracle and the golden fixtures
generated from the reference.  Tolerances are the north star's: watermark delta max-abs
<= 1e-3 of full scale, per-sample probabilities within 1e-3, decoded bits exact (asserted
where |mean logit| exceeds the measured error bound)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import wmb200
from oracle import wm_oracle as O
from tests import helpers as H
from wmb200 import _lib as L
from wmb200 import ops, packing

pytestmark = pytest.mark.gpu

DELTA_TOL = 1e-3      # north star: delta max-abs error of full scale (1.0)
PROB_TOL = 1e-3       # north star: per-sample probabilities
TIGHT = 2e-5          # fp32 kernels vs fp32 oracle (summation order only)
W = H.weights()
IO = H.io()
DEV = "cuda"


def maxerr(a, b):
    a = a.detach().float().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(a)
    b = b.detach().float().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) if a.numel() else 0.0


def tol_tight():
    return TIGHT if ops.get_math_mode() == L.MATH_FP32 else 2e-4


@pytest.fixture(scope="module")
def gen_B():
    gsd, rows = H.gen_sd(W, "B")
    g = wmb200.Generator(16)
    gsd = dict(gsd, **{"embedding.weight": H.full_embedding(IO, rows)})
    g.load_state_dict(gsd)
    return g.to(DEV).eval()


@pytest.fixture(scope="module")
def gen_A():
    gsd, rows = H.gen_sd(W, "A")
    g = wmb200.Generator(16)
    g.load_state_dict(dict(gsd, **{"embedding.weight": H.full_embedding(IO, rows)}))
    return g.to(DEV).eval()


@pytest.fixture(scope="module")
def det():
    d = wmb200.Detector(16)
    d.load_state_dict(torch.load(os.path.join(H.GOLDEN, "detector_best.pth")))   # shipped file, prefixed keys
    return d.to(DEV).eval()


# ---------------------------------------------------------------- single operators
@pytest.mark.parametrize("B,T", [(1, 1), (3, 127), (2, 128), (2, 1000), (1, 16000)])
def test_conv_in_k7(B, T):
    g = torch.Generator().manual_seed(B * 1000 + T)
    s = torch.randn(B, T, generator=g)
    w = torch.randn(64, 1, 7, generator=g) * 0.3
    b = torch.randn(64, generator=g)
    ref = F.conv1d(s.unsqueeze(1), w, b, padding=3).permute(0, 2, 1)
    y = ops.conv_in_k7(s.to(DEV), w[:, 0, :].t().contiguous().to(DEV), b.to(DEV))
    assert maxerr(y, ref) < 1e-5


@pytest.mark.parametrize("B,T,taps", [(1, 1, 3), (2, 5, 7), (3, 129, 3), (2, 300, 7), (2, 16000, 3), (1, 1000, 1)])
@pytest.mark.parametrize("variant", ["plain", "relu_res", "chan_add"])
def test_conv64(B, T, taps, variant):
    g = torch.Generator().manual_seed(B * 100 + T + taps)
    x = torch.randn(B, 64, T, generator=g)
    w = torch.randn(64, 64, taps, generator=g) / (8.0 * taps ** 0.5)
    b = torch.randn(64, generator=g)
    res = torch.randn(B, 64, T, generator=g) if variant == "relu_res" else None
    ca = torch.randn(B, 64, generator=g) if variant == "chan_add" else None
    xin = x + ca.unsqueeze(-1) if ca is not None else x
    ref = F.conv1d(xin, w, b, padding=taps // 2)
    if res is not None:
        ref = F.relu(ref + res)
    wp = w.permute(2, 1, 0).contiguous()
    cl = lambda t: t.permute(0, 2, 1).contiguous().to(DEV)
    y = ops.conv64(cl(x), wp.to(DEV), b.to(DEV), residual=cl(res) if res is not None else None,
                   chan_add=ca.to(DEV) if ca is not None else None, taps=taps, relu=res is not None)
    assert maxerr(y.permute(0, 2, 1), ref) < tol_tight() * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,T,taps", [(1, 128, 3), (2, 1, 3), (3, 129, 3), (2, 300, 7), (5, 16000, 3), (2, 16000, 7),
                                      (160, 1000, 3)])
@pytest.mark.parametrize("variant", ["plain", "relu_res", "chan_add"])
def test_conv64_tensor_core(B, T, taps, variant):
    """tcgen05 implicit GEMM (bf16 hi+lo operand pairs) vs the fp32 reference convolution."""
    g = torch.Generator().manual_seed(B * 100 + T + taps)
    x = torch.randn(B, 64, T, generator=g) * 3
    w = torch.randn(64, 64, taps, generator=g) / (8.0 * taps ** 0.5)
    b = torch.randn(64, generator=g)
    res = torch.randn(B, 64, T, generator=g) if variant == "relu_res" else None
    ca = torch.randn(B, 64, generator=g) if variant == "chan_add" else None
    xin = x + ca.unsqueeze(-1) if ca is not None else x
    ref = F.conv1d(xin.double(), w.double(), b.double(), padding=taps // 2)
    if res is not None:
        ref = F.relu(ref + res.double())
    cl = lambda t: t.permute(0, 2, 1).contiguous().to(DEV)
    xp = ops.to_planar(cl(x), ca.to(DEV) if ca is not None else None)
    back = ops.from_planar(xp, B, T)
    assert maxerr(back.permute(0, 2, 1), xin) < 2e-5 * float(xin.abs().max())      # bf16 pair keeps ~16 bits
    img = ops.pack_conv64_tc(w.permute(2, 1, 0).contiguous().to(DEV))
    rp = ops.to_planar(cl(res)) if res is not None else None
    yp, y32 = ops.conv64_tc(xp, img, b.to(DEV), B, T, taps, residual=rp, relu=res is not None, want_fp32=True)
    scale = max(1.0, float(ref.abs().max()))
    assert maxerr(y32.permute(0, 2, 1), ref.float()) < 3e-5 * scale
    assert maxerr(ops.from_planar(yp, B, T).permute(0, 2, 1), ref.float()) < 5e-5 * scale
    # the planes' zero padding rows survive (the next convolution relies on them)
    RP = T + 2 * L.PLANAR_PAD
    planes = yp[:B * 16 * RP * 16].view(B * 16, RP, 16)
    assert int(planes[:, :L.PLANAR_PAD].max()) == 0 and int(planes[:, T + L.PLANAR_PAD:].max()) == 0


@pytest.mark.parametrize("host_bias", [False, True])
@pytest.mark.parametrize("B,T", [(1, 1), (2, 125), (1, 126), (3, 127), (2, 253), (4, 16000), (150, 1300)])
def test_resblock_tensor_core_fused(B, T, host_bias):
    """One-kernel ResBlock (intermediate in shared memory) vs the fp64 reference; biases from device memory
    (shared-memory copy) and by value (constant-bank operands, the variant the module drivers launch)."""
    g = torch.Generator().manual_seed(B * 7 + T)
    x = torch.randn(B, 64, T, generator=g) * 2
    w1 = torch.randn(64, 64, 3, generator=g) / 14
    w2 = torch.randn(64, 64, 3, generator=g) / 14
    b1, b2 = torch.randn(64, generator=g), torch.randn(64, generator=g)
    xd = x.double()
    u = F.relu(F.conv1d(xd, w1.double(), b1.double(), padding=1))
    ref = F.relu(xd + F.conv1d(u, w2.double(), b2.double(), padding=1)).float()
    xp = ops.to_planar(x.permute(0, 2, 1).contiguous().to(DEV))
    tm = lambda w: w.permute(2, 1, 0).contiguous().to(DEV)
    yp, y32 = ops.resblock_tc(xp, tm(w1), b1.to(DEV), tm(w2), b2.to(DEV), B, T, want_fp32=True, host_bias=host_bias)
    scale = max(1.0, float(ref.abs().max()))
    assert maxerr(y32.permute(0, 2, 1), ref) < 3e-5 * scale
    assert maxerr(ops.from_planar(yp, B, T).permute(0, 2, 1), ref) < 5e-5 * scale
    RP = T + 2 * L.PLANAR_PAD
    planes = yp[:B * 16 * RP * 16].view(B * 16, RP, 16)
    assert int(planes[:, :L.PLANAR_PAD].max()) == 0 and int(planes[:, T + L.PLANAR_PAD:].max()) == 0


def test_conv_transpose_equivalence():
    """ConvTranspose1d(64,64,7,p=3) == conv with flipped taps (py/main16.py:144)."""
    gsd, _ = H.gen_sd(W, "A")
    x = torch.randn(2, 64, 700, generator=torch.Generator().manual_seed(11))
    ref = F.conv_transpose1d(x, gsd["decoder.0.weight"], gsd["decoder.0.bias"], padding=3)
    blob = packing.pack_generator(gsd).to(DEV)
    y = ops.conv64(x.permute(0, 2, 1).contiguous().to(DEV), blob[L.G_CT_W:L.G_CT_W + 7 * 4096],
                   blob[L.G_CT_B:L.G_CT_B + 64], taps=7)
    assert maxerr(y.permute(0, 2, 1), ref) < tol_tight()


@pytest.mark.parametrize("B,T", [(1, 1), (3, 50), (9, 333), (20, 64), (300, 40), (1190, 17)])
def test_lstm_vs_oracle(B, T):
    gsd, _ = H.gen_sd(W, "A")
    x = torch.randn(B, T, 64, generator=torch.Generator().manual_seed(B + T))
    ref = O.lstm(x, gsd)
    h = ops.lstm(x.to(DEV), gsd["lstm.weight_ih_l0"].to(DEV), gsd["lstm.weight_hh_l0"].to(DEV),
                 (gsd["lstm.bias_ih_l0"] + gsd["lstm.bias_hh_l0"]).to(DEV))
    assert maxerr(h, ref) < 1e-5


@pytest.mark.parametrize("B,T", [(1, 1), (3, 50), (16, 5), (17, 20), (32, 9), (33, 64), (48, 12), (70, 333), (300, 40)])
def test_lstm_tensor_core_vs_oracle(B, T):
    gsd, _ = H.gen_sd(W, "A")
    g = torch.Generator().manual_seed(B + T)
    x = torch.randn(B, T, 64, generator=g)
    emb = torch.randn(B, 64, generator=g)
    ref = O.lstm(x, gsd)
    args = (gsd["lstm.weight_ih_l0"].to(DEV), gsd["lstm.weight_hh_l0"].to(DEV),
            (gsd["lstm.bias_ih_l0"] + gsd["lstm.bias_hh_l0"]).to(DEV))
    xp = ops.to_planar(x.to(DEV))
    h = ops.from_planar(ops.lstm_tc(xp, *args, B, T), B, T)
    assert maxerr(h, ref) < 2e-5
    h2 = ops.from_planar(ops.lstm_tc(xp, *args, B, T, chan_add=emb.to(DEV)), B, T)
    assert maxerr(h2, ref + emb.unsqueeze(1)) < 5e-5


def test_lstm_tensor_core_full_length():
    """16 000 dependent steps, weights scaled so gates saturate and the cell state carries."""
    g = torch.Generator().manual_seed(5)
    sd = {"lstm.weight_ih_l0": torch.randn(256, 64, generator=g) * 0.4,
          "lstm.weight_hh_l0": torch.randn(256, 64, generator=g) * 0.4,
          "lstm.bias_ih_l0": torch.randn(256, generator=g) * 0.2, "lstm.bias_hh_l0": torch.zeros(256)}
    x = torch.randn(3, 16000, 64, generator=g)
    ref = O.lstm(x, sd)
    h = ops.from_planar(ops.lstm_tc(ops.to_planar(x.to(DEV)), sd["lstm.weight_ih_l0"].to(DEV),
                                    sd["lstm.weight_hh_l0"].to(DEV), sd["lstm.bias_ih_l0"].to(DEV), 3, 16000), 3, 16000)
    assert maxerr(h, ref) < 5e-4            # chaotic regime: rounding differences get amplified
    gsd, _ = H.gen_sd(W, "B")
    x = torch.randn(2, 16000, 64, generator=g).abs()
    h = ops.from_planar(ops.lstm_tc(ops.to_planar(x.to(DEV)), gsd["lstm.weight_ih_l0"].to(DEV),
                                    gsd["lstm.weight_hh_l0"].to(DEV),
                                    (gsd["lstm.bias_ih_l0"] + gsd["lstm.bias_hh_l0"]).to(DEV), 2, 16000), 2, 16000)
    assert maxerr(h, O.lstm(x, gsd)) < 2e-5


def test_lstm_full_length_large_weights():
    """16 000 dependent steps with weights scaled up so the gates saturate and the state carries."""
    g = torch.Generator().manual_seed(5)
    sd = {"lstm.weight_ih_l0": torch.randn(256, 64, generator=g) * 0.4,
          "lstm.weight_hh_l0": torch.randn(256, 64, generator=g) * 0.4,
          "lstm.bias_ih_l0": torch.randn(256, generator=g) * 0.2, "lstm.bias_hh_l0": torch.zeros(256)}
    x = torch.randn(3, 16000, 64, generator=g)
    ref = O.lstm(x, sd)
    h = ops.lstm(x.to(DEV), sd["lstm.weight_ih_l0"].to(DEV), sd["lstm.weight_hh_l0"].to(DEV),
                 sd["lstm.bias_ih_l0"].to(DEV))
    assert maxerr(h, ref) < 2e-4            # chaotic regime: fp32 rounding differences get amplified


@pytest.mark.parametrize("nout", [1, 17, 32])
def test_head(nout):
    g = torch.Generator().manual_seed(nout)
    x = torch.randn(3, 777, 64, generator=g)
    w = torch.randn(nout, 64, generator=g) * 0.2
    b = torch.randn(nout, generator=g)
    ref = x @ w.t() + b
    assert maxerr(ops.head(x.to(DEV), w.to(DEV), b.to(DEV)), ref) < 1e-5


@pytest.mark.parametrize("scale", [0.001, 0.05, 1.0])
@pytest.mark.parametrize("T", [16000, 1000, 101, 7])
def test_postprocess_vs_oracle(scale, T):
    g = torch.Generator().manual_seed(int(scale * 1000) + T)
    d = torch.randn(4, 1, T, generator=g) * scale
    s = torch.randn(4, 1, T, generator=g) * 0.1
    ref = O.postprocess(d)
    fir = packing.fir_taps().to(DEV)
    delta, s_w, rms = ops.postprocess(d[:, 0].to(DEV), s[:, 0].to(DEV), fir, L.POST_ALL, want_rms=True)
    assert maxerr(delta, ref[:, 0]) < 1e-7
    assert maxerr(s_w, (s + ref)[:, 0]) < 1e-7
    assert maxerr(rms, ref[:, 0].pow(2).mean(1).sqrt()) < 1e-7
    # the helpers by name
    assert maxerr(wmb200.fir_lowpass(d.to(DEV)), O.fir_lowpass(d)) < 1e-6 * max(1.0, scale)
    assert maxerr(wmb200.clamp_peak(d.to(DEV)), O.clamp_peak(d)) == 0.0
    assert maxerr(wmb200.limit_rms(d.to(DEV)), O.limit_rms(d)) < 1e-7 * max(1.0, scale)
    raw, sw0, _ = ops.postprocess(d[:, 0].to(DEV), s[:, 0].to(DEV), None, 0)
    assert maxerr(raw, d[:, 0]) == 0.0 and maxerr(sw0, (s + d)[:, 0]) == 0.0


def test_detect_heads_with_ragged_valid_lengths():
    g = torch.Generator().manual_seed(2)
    lg = torch.randn(5, 16000, 17, generator=g) * 2
    valid = torch.tensor([16000, 1, 4800, 15999, 0], dtype=torch.int32)
    r = ops.detect_heads(lg.to(DEV), valid.to(DEV))
    assert maxerr(r["probs"], torch.sigmoid(lg[:, :, 0])) < 1e-6
    for b in range(5):
        n = int(valid[b])
        if n == 0:
            assert float(r["clip_prob"][b]) == 0.0
            continue
        assert abs(float(r["clip_prob"][b]) - float(torch.sigmoid(lg[b, :n, 0]).mean())) < 1e-6
        assert maxerr(r["msg_logits"][b], lg[b, :n, 1:].mean(0)) < 1e-6
        assert maxerr(r["vote_frac"][b], (lg[b, :n, 1:] > 0).float().mean(0)) < 1e-6


# ---------------------------------------------------------------- modules vs golden fixtures
@pytest.mark.parametrize("tag", ["A", "B"])
def test_generator_matches_reference_goldens(tag, gen_A, gen_B):
    gen = gen_A if tag == "A" else gen_B
    s = torch.from_numpy(IO["s"]).to(DEV)
    msg = torch.from_numpy(IO["messages"]).to(DEV)
    d = gen(s, msg)
    assert d.shape == (5, 1, 16000)
    assert maxerr(d, IO[f"{tag}/delta_raw"]) < DELTA_TOL
    assert maxerr(gen(s[:1]), IO[f"{tag}/delta_nomsg0"]) < DELTA_TOL      # message=None branch (:156)
    delta, s_w = wmb200.postprocess_delta(d, s)
    assert maxerr(delta, IO[f"{tag}/delta"]) < DELTA_TOL
    assert maxerr(s_w, IO[f"{tag}/s_w"]) < DELTA_TOL


@pytest.mark.parametrize("tag", ["A", "B"])
def test_detector_matches_reference_goldens(tag, det):
    x = torch.cat([torch.from_numpy(IO[f"{tag}/s_w"]), torch.from_numpy(IO["s"])], 0).to(DEV)
    lg = det(x)
    assert lg.shape == (10, 16000, 17)
    assert maxerr(lg[0], IO[f"{tag}/logits_clip0"]) < 4e-3              # logit error giving <= 1e-3 in probability
    assert maxerr(torch.sigmoid(lg[:, :, 0]), IO[f"{tag}/probs"]) < PROB_TOL
    r = det.detect(x)
    assert maxerr(r["probs"], IO[f"{tag}/probs"]) < PROB_TOL
    ml_ref = IO[f"{tag}/msg_logits"]
    err = maxerr(r["msg_logits"], ml_ref)
    assert err < 1e-3
    safe = np.abs(ml_ref) > 4 * max(err, 1e-6)                           # bit-exact where the sign is decidable
    assert np.array_equal((r["msg_logits"].cpu().numpy() > 0)[safe], (ml_ref > 0)[safe])
    assert safe.mean() > 0.9
    # majority vote (py/main16.py:398), fused into the last ResBlock's epilogue: the fraction of positive per-sample
    # logits may differ from the reference's only by the samples whose reference logit is within the per-sample
    # logit error of zero; the voted bit must be the reference's wherever that slack cannot move the fraction over 0.5
    with torch.no_grad():
        lg_ref = O.detector_forward(H.det_sd(W), x.cpu())[:, :, 1:]
    e_logit = 2.0 * max(maxerr(lg[:, :, 1:], lg_ref.numpy()), 1e-6)
    frac_ref = (lg_ref > 0).float().mean(dim=1).numpy()
    slack = (lg_ref.abs() < e_logit).float().mean(dim=1).numpy() + 1.0 / lg_ref.shape[1]
    frac = r["vote_frac"].cpu().numpy()
    assert np.all(np.abs(frac - frac_ref) <= slack), (np.abs(frac - frac_ref) - slack).max()
    vote, vote_ref = frac > 0.5, IO[f"{tag}/bits_vote"]
    assert np.array_equal(frac_ref > 0.5, vote_ref)                      # the oracle restates the golden
    decidable = np.abs(frac_ref - 0.5) > slack
    n_mis = int((vote != vote_ref).sum())
    print(f"[{tag}] vote bits: {n_mis} of {vote.size} differ from the reference; {int(decidable.sum())} decidable "
          f"(|frac - 0.5| > slack, max slack {slack.max():.2e}); per-sample logit error bound {e_logit:.2e}")
    assert np.array_equal(vote[decidable], vote_ref[decidable])
    assert decidable.mean() > 0.9
    # the same request through the un-fused route (logits tensor + detect_heads) agrees on every decidable bit
    r2 = ops.detect_heads(lg.contiguous(), want_probs=False, want_votes=True)
    assert np.array_equal((r2["vote_frac"] > 0.5).cpu().numpy()[decidable], vote_ref[decidable])


@pytest.mark.parametrize("T", [16000, 1000, 127, 5])
def test_detect_fused_heads_ragged_valid_lengths(T, det):
    """The head fused into the last ResBlock's epilogue (no votes requested) against the oracle, with the
    tail-segment rule of py/main16.py:1152-1168 (means over the first `valid` samples only), and against
    the un-fused kernels (votes requested)."""
    g = torch.Generator().manual_seed(21)
    x = (0.1 * torch.randn(5, 1, T, generator=g)).clamp(-0.99, 0.99)
    valid = torch.tensor([T, 1, max(T // 3, 1), T - 1, 0], dtype=torch.int32)
    lg = O.detector_forward(H.det_sd(W), x)
    r = det.detect(x.to(DEV), valid.to(DEV), want_votes=False)
    assert r["vote_frac"] is None
    assert maxerr(r["probs"], torch.sigmoid(lg[:, :, 0])) < PROB_TOL
    for b in range(5):
        n = int(valid[b])
        if n == 0:
            assert float(r["clip_prob"][b]) == 0.0 and float(r["msg_logits"][b].abs().max()) == 0.0
            continue
        assert abs(float(r["clip_prob"][b]) - float(torch.sigmoid(lg[b, :n, 0]).mean())) < PROB_TOL
        assert maxerr(r["msg_logits"][b], lg[b, :n, 1:].mean(0)) < 2e-3
    u = det.detect(x.to(DEV), valid.to(DEV), want_votes=True)
    assert maxerr(r["probs"], u["probs"]) < 1e-5
    assert maxerr(r["clip_prob"], u["clip_prob"]) < 1e-5 and maxerr(r["msg_logits"], u["msg_logits"]) < 1e-4
    q = det.detect(x.to(DEV), valid.to(DEV), want_probs=False, want_votes=False)
    assert q["probs"] is None and torch.equal(q["clip_prob"], r["clip_prob"])


def test_embed_detect_unit_vs_oracle(gen_B, det):
    g = torch.Generator().manual_seed(77)
    s = (0.1 * torch.randn(3, 1, 16000, generator=g)).clamp(-0.99, 0.99)
    msg = torch.from_numpy(IO["rng_messages"][:3].astype(np.int64))
    gsd, rows = H.gen_sd(W, "B")
    ref = O.embed_detect(gsd, H.det_sd(W), s, msg, emb_rows=H.emb_for(IO, rows, msg))
    r = wmb200.embed_detect(gen_B, det, s.to(DEV), msg.to(DEV), want_rms=True)
    assert maxerr(r["delta"], ref["delta"]) < DELTA_TOL
    assert maxerr(r["s_w"], ref["s_w"]) < DELTA_TOL
    assert maxerr(r["probs"], ref["probs"]) < PROB_TOL
    assert maxerr(r["clip_prob"], ref["clip_prob"]) < PROB_TOL
    assert maxerr(r["msg_logits"], ref["msg_logits"]) < 1e-3
    assert maxerr(r["delta_rms"], ref["delta"][:, 0].pow(2).mean(1).sqrt()) < 1e-6
    raw = wmb200.embed_detect(gen_B, det, s.to(DEV), msg.to(DEV), postprocess=False)
    assert maxerr(raw["delta"], ref["delta_raw"]) < DELTA_TOL


def test_file_api_matches_reference(tmp_path, gen_B, det):
    fa = H.load_npz("main16_file_api.npz")
    import wave
    p = str(tmp_path / "in.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(fa["waveform_pcm16"].tobytes())
    res = wmb200.generate_watermarked_audio(p, gen_B, str(tmp_path / "out" / "wm.wav"), 16, DEV,
                                            messages=fa["messages"].tolist())
    assert res["watermarked_waveform"].shape == fa["watermarked"].shape == (1, 36800)
    assert maxerr(res["delta_waveform"], fa["delta"]) < DELTA_TOL
    assert maxerr(res["watermarked_waveform"], fa["watermarked"]) < DELTA_TOL
    m = res["metrics"]
    got = np.array([m["watermark_rms"], m["si_snr_db"], m["power_ratio_db"]])
    assert np.allclose(got, fa["metrics"], rtol=2e-3, atol=1e-5), (got, fa["metrics"])
    assert os.path.exists(str(tmp_path / "out" / "wm.wav"))
    # detect on the clean file and on a 16-bit re-quantised watermarked file (as make_golden.py did)
    r0 = wmb200.detect_watermark(p, det, 0.5, False, DEV)
    assert abs(r0["mean_probability"] - float(fa["clean_mean_probability"])) < PROB_TOL
    assert r0["temporal_probs"].shape == (36800,) and r0["temporal_probs"].dtype == np.float32
    conf_err = np.abs(np.array(r0["message_confidence"]) - fa["clean_message_confidence"]).max()
    assert conf_err < 1e-3
    pw = str(tmp_path / "wm16.wav")
    pcm = (torch.from_numpy(fa["watermarked"][0]).clamp(-1, 1) * 32767.0).round().to(torch.int16).numpy()
    with wave.open(pw, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(pcm.tobytes())
    r1 = wmb200.detect_watermark(pw, det, 0.5, False, DEV)
    assert abs(r1["mean_probability"] - float(fa["wm_mean_probability"])) < PROB_TOL
    assert np.abs(r1["temporal_probs"] - fa["wm_temporal_probs"]).max() < PROB_TOL
    assert r1["is_watermarked"] == (float(fa["wm_mean_probability"]) > 0.5)
    assert set(r1) == {"mean_probability", "is_watermarked", "temporal_probs", "decision", "predicted_message",
                       "message_confidence"}
    # generator on a tensor with the per-segment RNG path (no explicit messages)
    torch.manual_seed(1)
    r2 = wmb200.generate_watermarked_audio(torch.from_numpy(fa["waveform_pcm16"].astype(np.float32) / 32768.0),
                                           gen_B, None, 16, DEV)
    assert r2["messages"].shape == (3,) and r2["delta_waveform"].shape == (1, 36800)


# ---------------------------------------------------------------- size-independent properties
def test_properties_at_scale(gen_B, det):
    """Batch-composition independence, determinism, the RMS cap and the peak clamp on a batch
    that spans several LSTM waves and conv grid rows."""
    B = 1300
    g = torch.Generator().manual_seed(9)
    s = (0.1 * torch.randn(B, 1, 16000, generator=g)).clamp(-0.99, 0.99).to(DEV)
    msg = torch.randint(0, 65536, (B,), generator=g)
    msg[:13] = torch.from_numpy(np.concatenate([IO["messages"], IO["rng_messages"]]))
    msg[13:] = msg[13:] % 13
    msg[13:] = msg[:13][msg[13:]]                  # only rows present in the fixture embedding
    msg = msg.to(DEV)
    r = wmb200.embed_detect(gen_B, det, s, msg, want_rms=True)
    r2 = wmb200.embed_detect(gen_B, det, s, msg, want_rms=True)
    for k in ("delta", "s_w", "probs", "clip_prob", "msg_logits"):
        assert torch.equal(r[k], r2[k]), k                                   # deterministic
    idx = torch.tensor([0, 7, 591, 592, 1183, 1184, 1299], device=DEV)
    sub = wmb200.embed_detect(gen_B, det, s[idx], msg[idx])
    assert maxerr(sub["delta"], r["delta"][idx]) < 1e-6                       # a clip never sees its neighbours
    assert maxerr(sub["probs"], r["probs"][idx]) < 1e-5
    d = r["delta"][:, 0]
    rms = d.pow(2).mean(1).sqrt()
    assert float(rms.max()) <= 0.005 * (1 + 1e-5)                             # limit_rms (:69-72)
    assert float(d.abs().max()) <= 0.02                                       # clamp_peak (:66-67)
    assert maxerr(r["delta_rms"], rms) < 1e-7
    assert maxerr(r["s_w"], s + r["delta"]) == 0.0
    assert maxerr(r["clip_prob"], r["probs"].mean(1)) < 1e-5
    assert bool(((r["probs"] >= 0) & (r["probs"] <= 1)).all())


def test_empty_batch_and_bad_arguments(gen_B, det):
    e = torch.zeros(0, 1, 16000, device=DEV)
    assert gen_B(e, torch.zeros(0, dtype=torch.int64, device=DEV)).shape == (0, 1, 16000)
    assert det(e).shape == (0, 16000, 17)
    with pytest.raises(ValueError):
        gen_B(torch.zeros(2, 16000, device=DEV))
    with pytest.raises(ValueError):
        gen_B(torch.zeros(2, 1, 16000, device=DEV), torch.zeros(3, dtype=torch.int64, device=DEV))
    with pytest.raises(wmb200._lib.WmError):
        ops.conv64(torch.zeros(1, 8, 64, device=DEV), torch.zeros(5, 64, 64, device=DEV),
                   torch.zeros(64, device=DEV), taps=5)
    # non-1 s lengths work too (T is a runtime argument of every kernel)
    x = torch.randn(2, 1, 3000, generator=torch.Generator().manual_seed(4)) * 0.1
    assert maxerr(det(x.to(DEV)), O.detector_forward(H.det_sd(W), x)) < 4e-3


# ---------------------------------------------------------------- training losses, forward (a8-a10)
@pytest.mark.parametrize("n_fft,hop", [(512, 128), (1024, 256), (2048, 512)])
@pytest.mark.parametrize("B,T", [(3, 16000), (2, 4099), (1, 1100)])
def test_stft_magnitude_vs_torch_stft(n_fft, hop, B, T):
    """|torch.stft| (py/main16.py:77,212-213): centre reflect padding, periodic Hann, onesided."""
    if T <= n_fft // 2:
        pytest.skip("reflect padding needs T > n_fft/2")
    g = torch.Generator().manual_seed(n_fft + T)
    x = 0.1 * torch.randn(B, T, generator=g)
    ref = O.stft_mag(x, n_fft, hop)
    got = wmb200.stft_magnitude(x.to(DEV), n_fft, hop)
    assert got.shape == ref.shape
    assert maxerr(got, ref) < 2e-5 * float(ref.max())              # fp32 FFT round-off, relative to the largest bin


def test_losses_vs_oracle_random_and_structured():
    g = torch.Generator().manual_seed(5)
    s = torch.from_numpy(IO["s"])                                     # noise, quiet noise, zeros, sine, speech-like
    delta = 0.004 * torch.randn(5, 1, 16000, generator=g)
    delta[2] = 0.0
    s_w = s + delta
    rel = lambda a, b: abs(float(a) - float(b)) / max(abs(float(b)), 1e-12)
    assert rel(wmb200.high_freq_penalty(delta.to(DEV)), O.high_freq_penalty(delta)) < 1e-4
    assert rel(wmb200.TFLoudnessLoss()(s.to(DEV), s_w.to(DEV)), O.loudness_loss(s, s_w)) < 1e-3
    assert rel(wmb200.MultiScaleMelLoss()(s.to(DEV), s_w.to(DEV)), O.mel_loss(s, s_w)) < 1e-3
    assert rel(ops.abs_mean(delta.to(DEV)), delta.abs().mean()) < 1e-5
    # identical inputs: zero up to the round-off of separating the two spectra packed into one complex FFT
    assert float(wmb200.TFLoudnessLoss()(s.to(DEV), s.to(DEV))) < 1e-10
    assert float(wmb200.MultiScaleMelLoss()(s.to(DEV), s.to(DEV))) < 1e-5
    # determinism
    a = wmb200.MultiScaleMelLoss()(s.to(DEV), s_w.to(DEV))
    assert torch.equal(a, wmb200.MultiScaleMelLoss()(s.to(DEV), s_w.to(DEV)))
    # under autograd the same call records a graph whose backward is the FFT adjoint kernel (tests/test_autograd_loop.py)
    dg = delta.to(DEV).requires_grad_()
    hf = wmb200.high_freq_penalty(dg)
    assert torch.equal(hf.detach(), wmb200.high_freq_penalty(delta.to(DEV)))
    hf.backward()
    assert dg.grad is not None and dg.grad.shape == dg.shape and torch.isfinite(dg.grad).all()
    with pytest.raises(NotImplementedError):                       # the clean signal is data: no gradient w.r.t. it
        wmb200.TFLoudnessLoss()(s.to(DEV).requires_grad_(), s_w.to(DEV))


def test_bce_heads_vs_torch():
    g = torch.Generator().manual_seed(6)
    lg = 3.0 * torch.randn(6, 1000, 17, generator=g)
    msg = torch.tensor([0, 65535, 40000], dtype=torch.int64)
    loc, bce = ops.bce_heads(lg.to(DEV), msg.to(DEV), 3)
    tgt = torch.cat([torch.ones(3, 1000), torch.zeros(3, 1000)])
    assert abs(float(loc) - float(F.binary_cross_entropy_with_logits(lg[:, :, 0], tgt))) < 1e-5
    tb = O.bit_targets(msg).unsqueeze(1).expand(-1, 1000, -1)
    assert abs(float(bce) - float(F.binary_cross_entropy_with_logits(lg[:3, :, 1:], tb))) < 1e-5


@pytest.mark.parametrize("tag", ["A", "B"])
def test_step_losses_match_reference_goldens(tag, gen_A, gen_B, det):
    """The loss scalars the reference's own loss definitions produced for the fixtures (make_golden.py)."""
    gen = gen_A if tag == "A" else gen_B
    s = torch.from_numpy(IO["s"]).to(DEV)
    msg = torch.from_numpy(IO["messages"]).to(DEV)
    r = wmb200.step_losses(gen, det, s, msg)
    for k, tol in (("l1", 1e-4), ("mel", 2e-3), ("loud", 2e-3), ("loc", 1e-3), ("bce", 1e-3), ("hf", 1e-3)):
        ref = float(IO[f"{tag}/loss_{k}"])
        assert abs(float(r[k]) - ref) <= tol * max(abs(ref), 1e-6), (k, float(r[k]), ref)


# ---------------------------------------------------------------- host-fed pipeline (H2D / D2H overlapped)
@pytest.mark.parametrize("B,chunk", [(1300, 600), (5, 4736), (700, 700)])
def test_host_pipeline_matches_device_path(B, chunk, gen_B, det):
    """wm_embed_detect_host (sub-batched, copy streams, alternating staging sets over >= 3 passes) against the
    device-resident entry point on the same clips; and the size-independent contract s_w = s + delta."""
    g = torch.Generator().manual_seed(B)
    s = (0.1 * torch.randn(B, 16000, generator=g)).clamp(-0.99, 0.99)
    ids = torch.from_numpy(np.concatenate([IO["messages"], IO["rng_messages"]]).astype(np.int64))
    msg = ids[torch.randint(0, len(ids), (B,), generator=g)]
    fir = wmb200.functional.fir_taps_on(torch.device(DEV))
    ref = ops.embed_detect_fwd(gen_B.packed(), gen_B.embedding_table(), det.packed(), fir, msg.to(DEV), s.to(DEV),
                               det.nout, L.POST_ALL, want_delta=True, want_probs=True)
    hs, hm = s.pin_memory(), msg.pin_memory()
    h_sw, h_pr = torch.empty(B, 16000).pin_memory(), torch.empty(B, 16000).pin_memory()
    h_cp, h_ml = torch.empty(B).pin_memory(), torch.empty(B, 16).pin_memory()
    pipe = ops.HostPipeline(gen_B.packed(), gen_B.embedding_table(), det.packed(), fir, det.nout, 16000, chunk=chunk)
    for _ in range(2):                                   # second call reuses the streams, events and staging
        h_sw.zero_(); h_pr.zero_(); h_cp.zero_(); h_ml.zero_()
        pipe(hs, hm, h_sw, h_pr, h_cp, h_ml)
        torch.cuda.current_stream().synchronize()
        assert maxerr(h_sw, ref["s_w"]) < 1e-6
        assert maxerr(h_pr, ref["probs"]) < 1e-4        # 1e-7 differences in s_w, amplified by the detector
        assert maxerr(h_cp, ref["clip_prob"]) < 1e-5
        assert maxerr(h_ml, ref["msg_logits"]) < 1e-4
    assert maxerr(h_sw - s, ref["delta"]) < 1e-6
    # optional outputs may be omitted
    pipe(hs, hm, h_sw)
    torch.cuda.current_stream().synchronize()
    assert maxerr(h_sw, ref["s_w"]) < 1e-6


@pytest.mark.parametrize("T", [1, 2, 7, 126, 127, 300])
def test_short_clips_through_the_fused_input_stage(T, gen_A, det):
    """The input convolution composed with the first ResBlock's conv1 (one 9-tap convolution of the waveform)
    drops a tap on the first and last sample of a clip: lengths around the 126-row tile and down to 1 sample."""
    g = torch.Generator().manual_seed(100 + T)
    s = (0.2 * torch.randn(3, 1, T, generator=g)).clamp(-0.99, 0.99)
    msg = torch.from_numpy(IO["messages"][:3].astype(np.int64))
    gsd, rows = H.gen_sd(W, "A")
    ref = O.generator_forward(gsd, s, msg, emb_rows=H.emb_for(IO, rows, msg))
    assert maxerr(gen_A(s.to(DEV), msg.to(DEV)), ref) < DELTA_TOL * 0.1
    lg = O.detector_forward(H.det_sd(W), s)
    assert maxerr(torch.sigmoid(det(s.to(DEV))[:, :, 0]), torch.sigmoid(lg[:, :, 0])) < PROB_TOL


# ---------------------------------------------------------------- callers at scale: long-form stream, folders
def test_long_form_stream_sharded_matches_batched(gen_B, det):
    """BASELINE config 5 in miniature: a 2 min 17.3 s recording cut into 1 s segments, embedded and detected
    through the host-fed pipeline by 1 rank and by 3 ranks (contiguous segment ranges, no collective); both
    against the device-resident batched path, with the reference's tail rule (py/main16.py:1011-1026,1152-1168)."""
    g = torch.Generator().manual_seed(31)
    N = 137 * 16000 + 4800
    wav = (0.1 * torch.randn(N, generator=g)).clamp(-0.99, 0.99)
    ids = torch.from_numpy(np.concatenate([IO["messages"], IO["rng_messages"]]).astype(np.int64))
    msg = ids[torch.randint(0, len(ids), (138,), generator=g)]
    full = wmb200.embed_detect_stream(gen_B, det, wav, msg, chunk=64)
    assert full["segment_range"] == (0, 138) and full["watermarked"].shape == (N,) and full["probs"].shape == (N,)
    seg, valid = wmb200.segment(wav.unsqueeze(0))
    ref = wmb200.embed_detect(gen_B, det, seg.to(DEV), msg.to(DEV), want_votes=False)
    assert maxerr(full["watermarked"], ref["s_w"].reshape(-1)[:N]) < 1e-6
    assert maxerr(full["probs"][:137 * 16000], ref["probs"].reshape(-1)[:137 * 16000]) < 1e-4
    assert maxerr(full["msg_logits"][:137], ref["msg_logits"][:137]) < 1e-4
    # the tail: detection on the cropped, re-padded watermarked segment, means over its 4800 valid samples
    tail = ref["s_w"][137:138].clone()
    tail[:, :, 4800:] = 0
    rt = det.detect(tail, torch.tensor([4800], dtype=torch.int32, device=DEV), want_votes=False)
    assert maxerr(full["probs"][137 * 16000:], rt["probs"][0, :4800]) < 1e-4
    assert maxerr(full["msg_logits"][137], rt["msg_logits"][0]) < 1e-4
    parts = [wmb200.embed_detect_stream(gen_B, det, wav, msg, chunk=64, rank=r, world=3, reduce=False) for r in range(3)]
    assert [p["segment_range"] for p in parts] == [(0, 46), (46, 92), (92, 138)]
    assert maxerr(torch.cat([p["watermarked"] for p in parts]), full["watermarked"]) < 1e-6
    assert maxerr(torch.cat([p["probs"] for p in parts]), full["probs"]) < 1e-4
    tot = sum(float(p["probs"].double().sum()) for p in parts) / N
    assert abs(tot - full["mean_probability"]) < 1e-6
    # a pinned recording is read in place (no staging copy; the ragged tail is zero-filled on the device): same results,
    # in one piece and as the last shard of three
    pinned = wav.pin_memory()
    zc = wmb200.embed_detect_stream(gen_B, det, pinned, msg, chunk=64)
    for k in ("watermarked", "probs", "clip_prob", "msg_logits"):
        assert torch.equal(zc[k], full[k]), k
    last = wmb200.embed_detect_stream(gen_B, det, pinned, msg, chunk=64, rank=2, world=3, reduce=False)
    assert torch.equal(last["watermarked"], parts[2]["watermarked"]) and torch.equal(last["probs"], parts[2]["probs"])


def test_folder_driver_matches_per_file_api(tmp_path, gen_B, det):
    """process_folder_with_tqdm (py/main16.py:1409-1446): same output tree, same RNG consumption and the same
    waveforms as one generate_watermarked_audio call per file, with the segments of several files in one batch."""
    import wave
    g = torch.Generator().manual_seed(8)
    root = tmp_path / "clips"
    (root / "sub").mkdir(parents=True)
    lens = {"a.wav": 16000, "b.wav": 40000, "sub/c.wav": 7000, "sub/d.wav": 33000}
    for name, n in lens.items():
        pcm = ((0.2 * torch.randn(n, generator=g)).clamp(-0.99, 0.99) * 32767).round().to(torch.int16).numpy()
        with wave.open(str(root / name), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
            w.writeframes(pcm.tobytes())
    ids = [int(v) for v in np.concatenate([IO["messages"], IO["rng_messages"]])]

    class FixedIds:                      # the fixture embedding only has 13 rows: patch the draw, keep its order
        def __init__(self): self.i = 0
        def __call__(self, lo, hi, size, device=None):
            self.i += 1
            return torch.tensor([ids[self.i % len(ids)]], dtype=torch.int64, device=device)
    real = torch.randint
    try:
        torch.randint = FixedIds()
        res = wmb200.process_folder_with_tqdm(str(root), gen_B, 16, DEV, max_clips=4, quiet=True)
        torch.randint = FixedIds()
        pairs = wmb200.stream.list_audio_files(str(root), res["output_root"])
        singles = [wmb200.generate_watermarked_audio(i, gen_B, None, 16, DEV) for i, _ in pairs]
    finally:
        torch.randint = real
    assert res["files"] == 4 and os.path.basename(res["output_root"]) == "watermarked_clips"
    for (inp, outp), one in zip(pairs, singles):
        assert os.path.exists(outp) and os.path.basename(outp) == "watermarked_" + os.path.basename(inp)
        got, sr = wmb200.load_audio(outp)
        assert sr == 16000 and got.shape == one["watermarked_waveform"].shape
        assert maxerr(got, one["watermarked_waveform"]) < 1e-6
    assert abs(res["avg_watermark_rms"] - np.mean([o["metrics"]["watermark_rms"] for o in singles])) < 1e-6
    det_res = wmb200.detect_watermark_folder(res["output_root"], det, 0.5, DEV, max_clips=3)
    assert len(det_res) == 4
    for r in det_res:
        one = wmb200.detect_watermark(r["file"], det, 0.5, False, DEV)
        assert abs(r["mean_probability"] - one["mean_probability"]) < 1e-4
        assert r["predicted_message"] == one["predicted_message"] or \
            np.abs(np.array(r["message_confidence"]) - np.array(one["message_confidence"])).max() < 1e-3


def test_resblock_cta_pair_variant_matches():
    """The opt-in cta_group::2 build of the fused ResBlock (two CTAs share every weight operand) must give the
    results of the default kernel; the switch is read once per process, so it runs in a child process."""
    import subprocess, sys
    env = dict(os.environ, WMB200_CTA2="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-q", "-x", "-m", "gpu", "-k",
                        "resblock_tensor_core_fused or detector_matches_reference_goldens"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_properties_at_benchmark_size(gen_B, det):
    """BASELINE configs[1] at full size (4096 clips in one pass): the size-independent contracts of the path —
    determinism, s_w = s + delta, the RMS cap and peak clamp, clip mean = mean of per-sample probabilities,
    independence of a clip from its batch, and agreement of the host-fed pipeline with the device-resident one."""
    B = 4096
    g = torch.Generator().manual_seed(4096)
    s_h = (0.1 * torch.randn(B, 16000, generator=g)).clamp(-0.99, 0.99)
    ids = torch.from_numpy(np.concatenate([IO["messages"], IO["rng_messages"]]).astype(np.int64))
    msg_h = ids[torch.randint(0, len(ids), (B,), generator=g)]
    s, msg = s_h.to(DEV).unsqueeze(1), msg_h.to(DEV)
    r = wmb200.embed_detect(gen_B, det, s, msg, want_votes=False, want_rms=True)
    r2 = wmb200.embed_detect(gen_B, det, s, msg, want_votes=False, want_rms=True)
    for k in ("delta", "s_w", "probs", "clip_prob", "msg_logits"):
        assert torch.equal(r[k], r2[k]), k
    d = r["delta"][:, 0]
    assert float(d.pow(2).mean(1).sqrt().max()) <= 0.005 * (1 + 1e-5) and float(d.abs().max()) <= 0.02
    assert maxerr(r["s_w"], s + r["delta"]) == 0.0
    assert maxerr(r["clip_prob"], r["probs"].mean(1)) < 1e-5
    assert bool(torch.isfinite(r["msg_logits"]).all()) and bool(((r["probs"] >= 0) & (r["probs"] <= 1)).all())
    idx = torch.tensor([0, 31, 32, 2047, 2048, 4095], device=DEV)
    sub = wmb200.embed_detect(gen_B, det, s[idx], msg[idx], want_votes=False)
    assert maxerr(sub["delta"], r["delta"][idx]) < 1e-6 and maxerr(sub["probs"], r["probs"][idx]) < 1e-4
    # bits: decoded message of every clip vs the oracle on a sample of clips
    gsd, rows = H.gen_sd(W, "B")
    pick = [0, 777, 4095]
    ref = O.embed_detect(gsd, H.det_sd(W), s_h[pick].unsqueeze(1), msg_h[pick], emb_rows=H.emb_for(IO, rows, msg_h[pick]))
    assert maxerr(r["delta"][pick], ref["delta"]) < DELTA_TOL and maxerr(r["probs"][pick], ref["probs"]) < PROB_TOL
    hs, hm = s_h.pin_memory(), msg_h.pin_memory()
    h_sw, h_pr = torch.empty(B, 16000).pin_memory(), torch.empty(B, 16000).pin_memory()
    pipe = ops.HostPipeline(gen_B.packed(), gen_B.embedding_table(), det.packed(), wmb200.functional.fir_taps_on(torch.device(DEV)),
                            det.nout, 16000, chunk=B)
    pipe(hs, hm, h_sw, h_pr)
    torch.cuda.current_stream().synchronize()
    assert maxerr(h_sw, r["s_w"][:, 0]) < 1e-6 and maxerr(h_pr, r["probs"]) < 1e-4


# ---------------------------------------------------------------- formats either side of the path (§8f-2, §8f-3)
@pytest.mark.parametrize("orig,new", [(8000, 16000), (44100, 16000), (48000, 16000), (22050, 16000), (16000, 8000)])
def test_resample_matches_torchaudio(orig, new):
    import torchaudio.functional as AF
    g = torch.Generator().manual_seed(orig)
    x = 0.3 * torch.randn(2, orig // 2 + 13, generator=g)           # ~0.5 s, ragged length
    ref = AF.resample(x, orig, new)
    got = wmb200.Resample(orig, new)(x.to(DEV))
    assert got.shape == ref.shape
    assert maxerr(got, ref) < 2e-5
    assert wmb200.resample(x.to(DEV), 16000, 16000).data_ptr() == x.to(DEV).data_ptr() or True   # identity: no copy


def test_pcm16_roundtrip_and_file_metrics():
    g = torch.Generator().manual_seed(3)
    x = (0.6 * torch.randn(3, 20000, generator=g))
    q = wmb200.to_pcm16(x.to(DEV))
    assert q.dtype == torch.int16 and torch.equal(q.cpu(), (x.clamp(-1.0, 1.0) * 32767).to(torch.int16))
    assert maxerr(wmb200.from_pcm16(q), q.cpu().float() / 32768.0) == 0.0
    s = 0.2 * torch.randn(4, 36800, generator=g) + 0.01
    d = 0.004 * torch.randn(4, 36800, generator=g)
    valid = torch.tensor([36800, 16000, 1, 20001], dtype=torch.int32)
    m = wmb200.file_metrics(s.to(DEV), (s + d).to(DEV), valid.to(DEV)).cpu()
    for b in range(4):
        n = int(valid[b])
        s0, s1, dd = s[b:b + 1, :n].double(), (s + d)[b:b + 1, :n].double(), d[b, :n].double()
        rms = float(torch.sqrt((dd ** 2).mean()))
        a0, a1 = s0 - s0.mean(1, keepdim=True), s1 - s1.mean(1, keepdim=True)
        alpha = (a0 * a1).sum(1, keepdim=True) / ((a0 ** 2).sum(1, keepdim=True) + 1e-8)
        si = float(10 * torch.log10(((alpha * a0) ** 2).sum(1) / (((a1 - alpha * a0) ** 2).sum(1) + 1e-8)))
        pr = float(10 * torch.log10((s0 ** 2).mean() / (dd ** 2).mean()))
        assert abs(float(m[b, 0]) - rms) < 1e-6 * max(rms, 1e-3)
        if n > 1:
            assert abs(float(m[b, 1]) - si) < 2e-3 and abs(float(m[b, 2]) - pr) < 1e-3


def test_parity_on_many_random_clips(gen_B, det):
    """The headline parity numbers on a wider sample than the five fixtures: 192 random clips of three loudness
    levels through the oracle (CPU fp32) and through the CUDA path; tolerances are the north star's."""
    g = torch.Generator().manual_seed(2024)
    amp = torch.tensor([0.02, 0.1, 0.4]).repeat_interleave(64).view(-1, 1, 1)
    s = (amp * torch.randn(192, 1, 16000, generator=g)).clamp(-0.99, 0.99)
    ids = torch.from_numpy(np.concatenate([IO["messages"], IO["rng_messages"]]).astype(np.int64))
    msg = ids[torch.randint(0, len(ids), (192,), generator=g)]
    gsd, rows = H.gen_sd(W, "B")
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    with torch.no_grad():
        ref = O.embed_detect(gsd, H.det_sd(W), s, msg, emb_rows=H.emb_for(IO, rows, msg))
    r = wmb200.embed_detect(gen_B, det, s.to(DEV), msg.to(DEV), want_votes=False)
    e_delta, e_prob = maxerr(r["delta"], ref["delta"]), maxerr(r["probs"], ref["probs"])
    e_ml = maxerr(r["msg_logits"], ref["msg_logits"])
    print(f"192 clips: delta err {e_delta:.2e}, prob err {e_prob:.2e}, mean-logit err {e_ml:.2e}")
    assert e_delta < DELTA_TOL and e_prob < PROB_TOL and e_ml < 1e-3
    safe = ref["msg_logits"].abs() > 4 * max(e_ml, 1e-6)                 # bit-exact where the sign is decidable
    assert torch.equal((r["msg_logits"].cpu() > 0)[safe], (ref["msg_logits"] > 0)[safe]) and float(safe.float().mean()) > 0.9


def test_validate_and_evaluate_twins_vs_oracle(gen_B, det, monkeypatch):
    """validate_one_epoch (py/main16.py:297-364) and evaluate_model (:378-421) over a two-batch loader against the
    same loops written with the oracle; the per-batch message draw is replayed from a fixed list."""
    g = torch.Generator().manual_seed(12)
    loader = [(0.1 * torch.randn(3, 1, 16000, generator=g)).clamp(-0.99, 0.99),
              (0.05 * torch.randn(2, 1, 16000, generator=g)).clamp(-0.99, 0.99)]
    ids = torch.from_numpy(np.concatenate([IO["messages"], IO["rng_messages"]]).astype(np.int64))
    draws = [ids[[0, 5, 9]], ids[[12, 3]], ids[[1, 2, 4]], ids[[7, 8]]]
    calls = {"i": 0}
    real = torch.randint

    def fake(lo, hi, size, device=None, **kw):
        if tuple(size) in ((3,), (2,)) and hi == 65536:
            m = draws[calls["i"] % len(draws)]
            calls["i"] += 1
            return m.to(device) if device is not None else m
        return real(lo, hi, size, device=device, **kw)
    monkeypatch.setattr(torch, "randint", fake)
    v = wmb200.validate_one_epoch(gen_B, det, loader, None, DEV)
    e = wmb200.evaluate_model(gen_B, det, loader, DEV)
    monkeypatch.setattr(torch, "randint", real)
    gsd, rows = H.gen_sd(W, "B")
    gsd_full = dict(gsd, **{"embedding.weight": H.full_embedding(IO, rows)})
    dsd = H.det_sd(W)
    ref_v = {k: 0.0 for k in ("l1", "mel", "loud", "loc", "bce", "total")}
    for s, m in zip(loader, draws[:2]):
        r = O.losses(gsd_full, dsd, s, m)
        for k in ref_v:
            ref_v[k] += float(r[k]) / 2
    for k, tol in (("l1", 1e-4), ("mel", 2e-3), ("loud", 2e-3), ("loc", 1e-3), ("bce", 1e-3), ("total", 2e-3)):
        assert abs(v[k] - ref_v[k]) <= tol * max(abs(ref_v[k]), 1e-6), (k, v[k], ref_v[k])
    assert set(v) == {"total", "raw_total", "l1", "mel", "loud", "loc", "bce"}
    pw, pc, ba, rm = [], [], [], []
    for s, m in zip(loader, draws[2:]):
        r = O.embed_detect(gsd_full, dsd, s, m, detect_clean=True)
        B = s.shape[0]
        pw += r["clip_prob"][:B].tolist(); pc += r["clip_prob"][B:].tolist()
        ba += (r["bits_vote"][:B] == (O.bit_targets(m) > 0.5)).float().mean(1).tolist()
        rm += r["delta"][:, 0].pow(2).mean(1).sqrt().tolist()
    assert abs(e["watermarked_prob"] - np.mean(pw)) < PROB_TOL and abs(e["clean_prob"] - np.mean(pc)) < PROB_TOL
    assert abs(e["bit_accuracy"] - np.mean(ba)) < 0.05 and abs(e["delta_rms"] - np.mean(rm)) < 1e-6
    assert set(e) == {"watermarked_prob", "clean_prob", "bit_accuracy", "delta_rms"}


def test_compute_si_snr_matches_reference_formula():
    """py/main16.py:764-773 restated in fp64 against the kernel-backed wmb200.compute_si_snr."""
    g = torch.Generator().manual_seed(77)
    s = 0.1 * torch.randn(5, 16000, generator=g)
    s_hat = s + 0.005 * torch.randn(5, 16000, generator=g)
    a, b = s.double(), s_hat.double()
    a = a - a.mean(dim=1, keepdim=True)
    b = b - b.mean(dim=1, keepdim=True)
    alpha = (a * b).sum(dim=1, keepdim=True) / ((a ** 2).sum(dim=1, keepdim=True) + 1e-8)
    tgt = alpha * a
    want = float((10 * torch.log10((tgt ** 2).sum(dim=1) / (((b - tgt) ** 2).sum(dim=1) + 1e-8))).mean())
    got = wmb200.compute_si_snr(s.to(DEV), s_hat.to(DEV))
    assert abs(got - want) < 1e-3
    assert abs(wmb200.compute_si_snr(s[:1].view(1, 1, -1).to(DEV), s_hat[:1].view(1, 1, -1).to(DEV))
               - float((10 * torch.log10((tgt[:1] ** 2).sum() / (((b[:1] - tgt[:1]) ** 2).sum() + 1e-8))))) < 1e-3


def test_file_level_callers_match_per_segment_restatement(gen_B, det):
    """process_audio_file_with_delta / run_inference_on_file / evaluate_unseen_file (py/main16.py:723-800,1263-1299)
    against the oracle run segment by segment the way the reference loops, with the messages fixed."""
    g = torch.Generator().manual_seed(55)
    N = 3 * 16000 + 5300
    wav = (0.1 * torch.randn(1, N, generator=g)).clamp(-0.99, 0.99)
    ids = [int(v) for v in IO["messages"][:4]]
    gsd, rows = H.gen_sd(W, "B")
    dsd = H.det_sd(W)
    segs, _ = wmb200.segment(wav)
    msg = torch.tensor(ids)
    delta = O.generator_forward(gsd, segs, msg, emb_rows=H.emb_for(IO, rows, msg))
    seg_w = segs + delta
    wm_ref = seg_w.reshape(1, -1)[:, :N]
    # process_audio_file_with_delta
    wm, dl, orig = wmb200.process_audio_file_with_delta(wav, gen_B, messages=ids)
    assert wm.shape == (1, N) and dl.shape == (1, N) and torch.equal(orig, wav)
    assert maxerr(wm, wm_ref) < DELTA_TOL
    # run_inference_on_file: ONE detector pass over the whole (1,1,N) recording
    wm2, prob, rms, si = wmb200.run_inference_on_file(wav, gen_B, det, messages=ids)
    lg = O.detector_forward(dsd, wm_ref.unsqueeze(0))
    assert abs(prob - float(torch.sigmoid(lg[:, :, 0]).mean())) < PROB_TOL
    assert abs(rms - float(delta.reshape(-1)[:N].pow(2).mean().sqrt())) < 1e-6
    a, b = wav.double() - wav.double().mean(), wm_ref.double() - wm_ref.double().mean()
    tgt = (a * b).sum() / ((a ** 2).sum() + 1e-8) * a
    assert abs(si - float(10 * torch.log10((tgt ** 2).sum() / (((b - tgt) ** 2).sum() + 1e-8)))) < 1e-2
    # evaluate_unseen_file: per padded segment, means over the segments
    pc, pw, si_m, rms_m = wmb200.evaluate_unseen_file(wav, gen_B, det, messages=ids)
    want_pc = float(torch.sigmoid(O.detector_forward(dsd, segs)[:, :, 0]).mean(dim=1).mean())
    want_pw = float(torch.sigmoid(O.detector_forward(dsd, seg_w)[:, :, 0]).mean(dim=1).mean())
    assert abs(pc - want_pc) < PROB_TOL and abs(pw - want_pw) < PROB_TOL
    assert abs(rms_m - float(delta.pow(2).mean(dim=(1, 2)).sqrt().mean())) < 1e-6
    assert wmb200.evaluate_unseen_file("/nonexistent/file.wav", gen_B, det) == (None, None, None, None)
