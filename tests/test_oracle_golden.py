"""Pin oracle/wm_oracle.py against outputs of the reference's own definitions
(tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import wm_oracle as O
from tests import helpers as H

W = H.weights()
IO = H.io()
S = torch.from_numpy(IO["s"])
MSG = torch.from_numpy(IO["messages"])
DSD = H.det_sd(W)


def close(a, b, tol):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    assert a.shape == b.shape
    assert np.abs(a - b).max() <= tol, np.abs(a - b).max()


@pytest.mark.parametrize("tag", ["A", "B"])
def test_generator_and_post(tag):
    gsd, rows = H.gen_sd(W, tag)
    emb = H.emb_for(IO, rows, MSG)
    d_raw = O.generator_forward(gsd, S, MSG, emb_rows=emb)
    close(d_raw, IO[f"{tag}/delta_raw"], 2e-6)
    close(O.generator_forward(gsd, S[:1], None), IO[f"{tag}/delta_nomsg0"], 2e-6)
    d = O.postprocess(torch.from_numpy(IO[f"{tag}/delta_raw"]))
    close(d, IO[f"{tag}/delta"], 1e-8)
    x = O.generator_encoder(gsd, S)
    idx = torch.from_numpy(IO["sample_idx"])
    close(x[:, :, idx], IO[f"{tag}/enc_samples"], 1e-5)
    h = O.lstm(x.permute(0, 2, 1), gsd)
    close(h[:, idx, :], IO[f"{tag}/lstm_samples"], 1e-5)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_detector_heads_and_losses(tag):
    s_w = torch.from_numpy(IO[f"{tag}/s_w"])
    lg = O.detector_forward(DSD, torch.cat([s_w, S], 0))
    assert lg.shape == (10, 16000, 17)
    close(lg[0], IO[f"{tag}/logits_clip0"], 1e-5)
    close(torch.sigmoid(lg[:, :, 0]), IO[f"{tag}/probs"], 1e-6)
    close(lg[:, :, 1:].mean(1), IO[f"{tag}/msg_logits"], 1e-6)
    gsd, rows = H.gen_sd(W, tag)
    r = O.losses(gsd, DSD, S, MSG, emb_rows=H.emb_for(IO, rows, MSG))
    for k in ("l1", "mel", "loud", "loc", "bce", "hf"):
        ref = float(IO[f"{tag}/loss_{k}"])
        assert abs(float(r[k]) - ref) <= 1e-5 * max(1.0, abs(ref)), (k, float(r[k]), ref)


def test_fir_taps_and_bits():
    close(O.fir_taps(), IO["fir_taps"], 0.0)
    t = O.fir_taps()
    assert abs(float(t[50]) - 1.0) < 1e-6 and float(t.abs().sum() - t[50].abs()) < 1e-5
    b = O.bit_targets(torch.tensor([5, 40000]))
    assert b[0].tolist() == [1, 0, 1] + [0] * 13
    assert int((b[1] * (1 << torch.arange(16))).sum()) == 40000


def test_lstm_library_call_equals_step_restatement():
    gsd, _ = H.gen_sd(W, "A")
    x = torch.randn(3, 200, 64, generator=torch.Generator().manual_seed(1))
    close(O.lstm(x, gsd), O.lstm_steps(x, gsd).numpy(), 2e-6)


def test_mel_filterbank_shape_and_support():
    fb = O.mel_filterbank()
    assert fb.shape == (513, 64)
    assert int((fb > 0).sum()) == 1000
