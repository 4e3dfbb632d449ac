"""The formats and evaluation reductions either side of the path, on the GPU: the biquad + PCM16
save path of py/main15.py:850-867, torchaudio's Resample inside the file API, confusion counts / ROC / AUC of the
evaluation cells (py/main16.py:1335-1341, 2372-2386).  Checked against torchaudio and sklearn on the CPU."""
import os
import wave

import numpy as np
import pytest
import torch

import wmb200
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("rows,N", [(1, 1), (1, 2), (2, 63), (1, 64), (3, 65), (2, 8193), (1, 160001), (4, 48000)])
def test_lowpass_biquad_and_pcm16_match_torchaudio(rows, N):
    import torchaudio.functional as AF
    g = torch.Generator().manual_seed(rows * 1000 + N)
    x = (0.6 * torch.randn(rows, N, generator=g)).clamp(-1.5, 1.5)          # some samples beyond full scale: the clamp matters
    ref = AF.lowpass_biquad(x, 16000, cutoff_freq=7000)                        # py/main15.py:855
    ref_q = (ref.clamp(-1.0, 1.0) * 32767).to(torch.int16)                    # py/main15.py:859-860
    y, q = wmb200.perceptual_postprocess(x.to(DEV), 16000, 7000.0)
    assert y.shape == x.shape and q.dtype == torch.int16
    assert float((y.cpu() - ref).abs().max()) < 2e-6
    dq = (q.cpu().int() - ref_q.int()).abs()
    # truncation toward zero of two values ~1e-7 apart: they straddle an integer with probability ~2 * 1e-7 * 32767
    assert int(dq.max()) <= 1 and float((dq > 0).float().sum()) <= max(2.0, 0.02 * dq.numel())
    assert torch.equal(wmb200.lowpass_biquad(x.to(DEV), 16000, 7000.0), y)
    # a general second-order section without lfilter's clamp
    z = wmb200.biquad(x.to(DEV), 0.3, -0.1, 0.05, 1.25, -0.9, 0.4, clamp=False)
    zr = AF.lfilter(x, torch.tensor([1.25, -0.9, 0.4]), torch.tensor([0.3, -0.1, 0.05]), clamp=False)
    assert float((z.cpu() - zr).abs().max()) < 5e-6 * max(1.0, float(zr.abs().max()))


def test_save_audio_pcm16_writes_the_filtered_codes(tmp_path):
    import torchaudio.functional as AF
    x = 0.3 * torch.randn(1, 20000, generator=torch.Generator().manual_seed(5))
    path = str(tmp_path / "out.wav")
    wmb200.save_audio_pcm16(x.to(DEV), path, 16000)
    with wave.open(path, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, 20000)
        got = np.frombuffer(w.readframes(20000), dtype="<i2")
    ref = (AF.lowpass_biquad(x, 16000, cutoff_freq=7000).clamp(-1, 1) * 32767).to(torch.int16).numpy()[0]
    assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= 1


def test_file_api_resamples_on_the_device(tmp_path, monkeypatch):
    """A 44.1 kHz stereo file through generate_watermarked_audio / detect_watermark: mono mix and torchaudio's Resample
    (py/main16.py:981-985) with the resampler running in libwmb200 -- torchaudio's transform must not be called."""
    import torchaudio
    from scipy.io import wavfile
    sr, n = 44100, 44100 * 2 + 1234
    g = torch.Generator().manual_seed(9)
    x = (0.2 * torch.randn(2, n, generator=g)).clamp(-0.99, 0.99)
    path = str(tmp_path / "in44.wav")
    wavfile.write(path, sr, np.ascontiguousarray(x.numpy().T))
    want = torchaudio.transforms.Resample(sr, 16000)(x.mean(dim=0, keepdim=True))

    def boom(*a, **k):
        raise AssertionError("the file API must resample on the GPU")
    monkeypatch.setattr(torchaudio.transforms, "Resample", boom)
    W, IO = H.weights(), H.io()
    gsd, rows = H.gen_sd(W, "B")
    gen = wmb200.Generator(16)
    gen.load_state_dict(dict(gsd, **{"embedding.weight": H.full_embedding(IO, rows)}))
    det = wmb200.Detector(16)
    det.load_state_dict(torch.load(os.path.join(H.GOLDEN, "detector_best.pth")))
    gen, det = gen.to(DEV).eval(), det.to(DEV).eval()
    nseg = (want.shape[1] + 15999) // 16000
    r = wmb200.generate_watermarked_audio(path, gen, None, 16, DEV, messages=[int(IO["messages"][0])] * nseg)
    assert r["original_waveform"].shape == want.shape
    assert float((r["original_waveform"] - want).abs().max()) < 2e-6
    assert r["watermarked_waveform"].shape == want.shape
    d = wmb200.detect_watermark(path, det, 0.5, False, DEV)
    assert d["temporal_probs"].shape == (want.shape[1],) and 0.0 <= d["mean_probability"] <= 1.0


@pytest.mark.parametrize("n0,n1,ties", [(1, 1, False), (37, 53, False), (400, 300, True), (5000, 4096, True)])
def test_confusion_roc_auc_match_sklearn(n0, n1, ties):
    from sklearn import metrics as SK
    g = torch.Generator().manual_seed(n0 + 7 * n1)
    clean = torch.rand(n0, generator=g) * 0.7
    wm = 0.3 + torch.rand(n1, generator=g) * 0.7
    if ties:                                    # quantised scores: many exact ties, also across the classes
        clean, wm = (clean * 50).round() / 50, (wm * 50).round() / 50
    y_true = [0] * n0 + [1] * n1
    scores = torch.cat([clean, wm]).numpy()
    for thr in (0.5, 0.3, 0.0, 1.1):
        k = wmb200.confusion_counts(clean.to(DEV), wm.to(DEV), thr)
        y_pred = [1 if p >= thr else 0 for p in scores]                        # py/main16.py:1337
        assert np.array_equal(k["matrix"], SK.confusion_matrix(y_true, y_pred, labels=[0, 1]))
    rep = wmb200.classification_report(clean.to(DEV), wm.to(DEV), 0.5)
    ref = SK.classification_report(y_true, [1 if p >= 0.5 else 0 for p in scores], labels=[0, 1],
                                   target_names=["Clean", "Watermarked"], output_dict=True, zero_division=0)
    for name in ("Clean", "Watermarked"):
        for key in ("precision", "recall", "f1-score", "support"):
            assert abs(rep[name][key] - ref[name][key]) < 1e-12
    assert abs(rep["accuracy"] - ref["accuracy"]) < 1e-12
    fpr, tpr, thr = wmb200.roc_curve(clean.to(DEV), wm.to(DEV))
    rf, rt, rth = SK.roc_curve(y_true, scores, drop_intermediate=False)
    assert np.allclose(fpr, rf, atol=0) and np.allclose(tpr, rt, atol=0) and np.array_equal(thr[1:], rth[1:])
    a = wmb200.auc(clean.to(DEV), wm.to(DEV))
    assert abs(a - SK.auc(rf, rt)) < 1e-12 and abs(a - SK.roc_auc_score(y_true, scores)) < 1e-12
