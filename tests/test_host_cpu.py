"""CPU-side tests: host logic of the package, the packed-weight layout against the
oracle, the C-ABI library (loads, exports every declared symbol), sharding over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import wmb200
from oracle import wm_oracle as O
from tests import helpers as H
from wmb200 import _lib as L
from wmb200 import packing

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = H.weights()
IO = H.io()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "wmb200.h")).read()
    declared = set(re.findall(r"\b(wm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert L.load().wm_abi_version() == L.ABI_VERSION
    m = re.search(r"#define WM_ABI_VERSION (\d+)", header)
    assert int(m.group(1)) == L.ABI_VERSION


def test_blob_offsets_match_header():
    header = open(os.path.join(ROOT, "include", "wmb200.h")).read()
    src = "#include <stdio.h>\n" + header + "\nint main(){printf(\"%d %d %d %d %d %d\\n\", WM_RB_SIZE, WM_G_SIZE, " \
          "WM_D_SIZE, WM_G_CT_W, WM_D_HEAD_W, WM_G_LSTM_B);return 0;}\n"
    exe = os.path.join(ROOT, "tests", "golden", "_tmp_offsets")
    subprocess.run(["gcc", "-x", "c", "-", "-o", exe], input=src.encode(), check=True)
    out = subprocess.run([exe], capture_output=True, check=True).stdout.split()
    os.remove(exe)
    assert [int(v) for v in out] == [L.RB_SIZE, L.G_SIZE, L.D_SIZE, L.G_CT_W, L.D_HEAD_W, L.G_LSTM_B]


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = wmb200.Detector(16).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d(torch.zeros(1, 1, 16000))
    assert L.load().wm_device_ok() == 0
    rc = L.load().wm_conv_in_k7_fwd(None, None, None, None, 1, 1, None)
    assert rc != 0 and b"" != L.load().wm_last_error()


def test_train_mode_has_no_cpu_fallback_either():
    """Train-mode forward runs the autograd path on the GPU kernels; on CPU tensors it fails as loudly as eval mode."""
    g = wmb200.Generator(16)
    assert g.training
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.zeros(1, 1, 16000))


def test_state_dict_compat_and_prefix():
    d = wmb200.Detector(16)
    sd = torch.load(os.path.join(H.GOLDEN, "detector_best.pth"))
    assert all(k.startswith("_orig_mod.") for k in sd)
    res = d.load_state_dict(sd)                       # README path (README.md:118) works on the shipped file
    assert not res.missing_keys and not res.unexpected_keys
    d2 = wmb200.Detector(16)
    wmb200.load_state_dict_strip_prefix(d2, sd)       # py/main16.py:707-712
    for (k1, v1), (k2, v2) in zip(d.state_dict().items(), d2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    gsd, _ = H.gen_sd(W, "A")
    g = wmb200.Generator(16)
    want = set(g.state_dict()) - {"embedding.weight"}
    assert want == set(gsd)
    assert g.state_dict()["embedding.weight"].shape == (65536, 64)
    assert wmb200.Generator(0).message_bits == 0 and not hasattr(wmb200.Generator(0), "embedding")


def _conv_from_blob(x, blob, off_w, off_b, taps):
    w = blob[off_w:off_w + taps * 4096].view(taps, 64, 64).permute(2, 1, 0)   # -> (co,ci,j)
    return F.conv1d(x, w, blob[off_b:off_b + 64], padding=taps // 2)


def _rb_from_blob(x, blob, off):
    y = F.relu(_conv_from_blob(x, blob, off + L.RB_W1, off + L.RB_B1, 3))
    return F.relu(x + _conv_from_blob(y, blob, off + L.RB_W2, off + L.RB_B2, 3))


@pytest.mark.parametrize("tag", ["A", "B"])
def test_packed_generator_blob_reproduces_the_oracle(tag):
    """BN folding, tap-major layout and the ConvTranspose flip, checked on CPU."""
    gsd, rows = H.gen_sd(W, tag)
    blob = packing.pack_generator(gsd)
    assert blob.shape == (L.G_SIZE,) and blob.dtype == torch.float32
    s = torch.from_numpy(IO["s"])[:2, :, :2000]
    x = F.conv1d(s, blob[L.G_IN_W:L.G_IN_W + 448].view(7, 64).t().unsqueeze(1), blob[L.G_IN_B:L.G_IN_B + 64], padding=3)
    x = _rb_from_blob(_rb_from_blob(x, blob, L.G_RB0), blob, L.G_RB1)
    ref = O.generator_encoder(gsd, s)
    assert (x - ref).abs().max() < 2e-5
    h = torch.randn(2, 64, 500, generator=torch.Generator().manual_seed(3))
    y = _conv_from_blob(h, blob, L.G_CT_W, L.G_CT_B, 7)
    y = _rb_from_blob(y, blob, L.G_RB2)
    y = F.conv1d(y, blob[L.G_HEAD_W:L.G_HEAD_W + 64].view(1, 64, 1), blob[L.G_HEAD_B:L.G_HEAD_B + 1])
    assert (y - O.generator_decoder(gsd, h)).abs().max() < 2e-5
    b = blob[L.G_LSTM_B:L.G_LSTM_B + 256]
    assert torch.allclose(b, gsd["lstm.bias_ih_l0"] + gsd["lstm.bias_hh_l0"])


def test_packed_detector_blob_reproduces_the_oracle():
    dsd = H.det_sd(W)
    blob = packing.pack_detector(dsd)
    s = torch.from_numpy(IO["s"])[:3, :, :3000]
    x = F.conv1d(s, blob[L.D_IN_W:L.D_IN_W + 448].view(7, 64).t().unsqueeze(1), blob[L.D_IN_B:L.D_IN_B + 64], padding=3)
    x = _rb_from_blob(_rb_from_blob(x, blob, L.D_RB0), blob, L.D_RB1)
    lg = F.conv1d(x, blob[L.D_HEAD_W:L.D_HEAD_W + 17 * 64].view(17, 64, 1), blob[L.D_HEAD_B:L.D_HEAD_B + 17])
    ref = O.detector_forward(dsd, s).permute(0, 2, 1)
    assert (lg - ref).abs().max() < 5e-5


def test_fir_taps_equal_reference():
    assert np.array_equal(packing.fir_taps().numpy(), IO["fir_taps"])


def test_segment_and_wav_roundtrip(tmp_path):
    x = torch.linspace(-0.5, 0.5, 36800).unsqueeze(0)
    batch, valid = wmb200.segment(x)
    assert batch.shape == (3, 1, 16000) and valid.tolist() == [16000, 16000, 4800]
    assert torch.equal(batch.view(-1)[:36800], x[0]) and batch.view(-1)[36800:].abs().sum() == 0
    b2, v2 = wmb200.segment(x[:, :32000])
    assert b2.shape[0] == 2 and v2.tolist() == [16000, 16000]
    p = str(tmp_path / "a.wav")
    wmb200.save_audio(p, x)
    y, sr = wmb200.load_audio(p)
    assert sr == 16000 and torch.equal(y, x)


def test_shard_range_partitions():
    for n in (0, 1, 7, 36000, 4096):
        for w in (1, 2, 4, 8):
            spans = [wmb200.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        wmb200.shard_range(4, 2, 2)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["WM_ROOT"])
import wmb200
from wmb200.sharding import reduce_file_stats, shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["WM_PORT"],
                        rank=int(os.environ["WM_RANK"]), world_size=2)
rank = dist.get_rank()
g = torch.Generator().manual_seed(5)
probs = torch.rand(7, 16000, generator=g); probs[-1, 4800:] = 0
ml = torch.randn(7, 16, generator=g)
lo, hi = shard_range(7, rank, 2)
nsamp = sum(16000 if i < 6 else 4800 for i in range(lo, hi))
mp, mlog = reduce_file_stats(float(probs[lo:hi].double().sum()), nsamp, ml[lo:hi].sum(0), hi - lo)
want_p = float(probs.double().sum() / (6 * 16000 + 4800)); want_l = ml.mean(0)
assert abs(mp - want_p) < 1e-9, (mp, want_p)
assert (mlog - want_l).abs().max() < 1e-6
dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_gloo_shard_and_reduce(tmp_path):
    """world_size-2 run of the N>1 path's only exchange (file-level aggregates)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, WM_ROOT=ROOT, WM_PORT=port, WM_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0 and "ok" in out, out


def test_mel_filterbank_matches_oracle_and_band_covers_nonzeros():
    from oracle import wm_oracle as O
    from wmb200 import packing
    fb, band = packing.mel_filterbank()
    assert fb.shape == (513, 64) and torch.equal(fb, O.mel_filterbank())
    mask = torch.zeros_like(fb, dtype=torch.bool)
    for m in range(64):
        mask[band[m, 0]:band[m, 1], m] = True
    assert int(((fb != 0) & ~mask).sum()) == 0
    assert int((band[:, 1] - band[:, 0]).max()) <= 41                  # 3-41 bins per filter


def test_losses_refuse_cpu_tensors():
    import wmb200
    with pytest.raises(RuntimeError):
        wmb200.high_freq_penalty(torch.zeros(1, 1, 16000))
    with pytest.raises(RuntimeError):
        wmb200.TFLoudnessLoss()(torch.zeros(1, 1, 16000), torch.zeros(1, 1, 16000))


def test_folder_driver_planning(tmp_path):
    from wmb200 import stream
    assert stream.plan_batches([3, 2, 7, 1, 1], 5) == [(0, 2), (2, 3), (3, 5)]
    assert stream.plan_batches([9], 4) == [(0, 1)] and stream.plan_batches([], 4) == []
    assert stream.plan_batches([0, 0, 2], 4) == [(0, 3)]
    root = tmp_path / "speech"
    (root / "a").mkdir(parents=True)
    for name in ("x.wav", "a/y.FLAC", "a/notes.txt"):
        (root / name).write_bytes(b"")
    pairs = stream.list_audio_files(str(root), str(tmp_path / "watermarked_speech"))
    outs = sorted(os.path.relpath(o, str(tmp_path)) for _, o in pairs)
    assert outs == [os.path.join("watermarked_speech", ".", "watermarked_x.wav"),
                    os.path.join("watermarked_speech", "a", "watermarked_y.FLAC")] or \
        sorted(os.path.normpath(o) for o in outs) == [os.path.join("watermarked_speech", "a", "watermarked_y.FLAC"),
                                                      os.path.join("watermarked_speech", "watermarked_x.wav")]
