"""File-level API of the reference's README (py/main16.py:977-1066, 1114-1207) on libwmb200.

Same arguments, same result dictionaries.  The difference is inside: the reference runs
one Generator/Detector call per second of audio at batch 1 with two host syncs each
(py/main16.py:996-1009, 1133-1150); here all segments of a file form ONE batch.
"""
from __future__ import annotations

import os
import wave
from typing import Optional, Sequence

import numpy as np
import torch

from .functional import SAMPLE_RATE

SEG = SAMPLE_RATE


# ---- audio I/O (host side; the reference uses torchaudio.load/save) -------------
def load_audio(path: str):
    """-> (waveform (C,N) fp32 in [-1,1), sample_rate).  PCM WAV via the stdlib, other
    formats via torchaudio when its backend is available."""
    try:
        with wave.open(path, "rb") as w:
            ch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
        if width == 2:
            x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif width == 4:
            x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
        elif width == 1:
            x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif width == 3:
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v >= 1 << 23, v - (1 << 24), v)
            x = v.astype(np.float32) / 8388608.0
        else:
            raise wave.Error(f"unsupported sample width {width}")
        return torch.from_numpy(x.reshape(-1, ch).T.copy()), sr
    except (wave.Error, EOFError):
        pass
    try:
        from scipy.io import wavfile
        sr, x = wavfile.read(path)
        x = np.asarray(x)
        if x.dtype.kind == "f":
            y = x.astype(np.float32)
        else:
            y = x.astype(np.float32) / float(np.iinfo(x.dtype).max + 1)
        y = y.reshape(len(y), -1).T
        return torch.from_numpy(np.ascontiguousarray(y)), int(sr)
    except Exception:
        import torchaudio
        return torchaudio.load(path)


def save_audio(path: str, waveform: torch.Tensor, sample_rate: int = SAMPLE_RATE) -> None:
    """Write fp32 (C,N) audio.  torchaudio.save writes float WAV by default; so does this."""
    from scipy.io import wavfile
    x = waveform.detach().to("cpu", torch.float32).numpy()
    wavfile.write(path, sample_rate, np.ascontiguousarray(x.T))


def _prepare(input_file, device="cuda"):
    """load -> mono -> 16 kHz  (py/main16.py:981-985).  Returns the (1, N) waveform on the HOST (what the result
    dictionaries carry); a file at another sample rate is resampled on `device` by the library's polyphase kernel
    (audio.resample, torchaudio.transforms.Resample's arithmetic) -- there is no CPU resampler on this path."""
    if isinstance(input_file, torch.Tensor):
        waveform, sr = input_file.to("cpu", torch.float32), SAMPLE_RATE
        if waveform.dim() == 1:
            waveform = waveform.unsqueeze(0)
    else:
        waveform, sr = load_audio(input_file)
    if waveform.shape[0] > 1:
        waveform = waveform.mean(dim=0, keepdim=True)
    if sr != SAMPLE_RATE:
        from .audio import resample
        waveform = resample(waveform.to(device, torch.float32), sr, SAMPLE_RATE).cpu()
    return waveform


def segment(waveform: torch.Tensor):
    """(1,N) -> (segments (n,1,16000) zero-padded, valid_len (n,) int32)  (py/main16.py:987-990, 1011-1013)."""
    total = waveform.shape[1]
    n = (total + SEG - 1) // SEG
    batch = torch.zeros(n, 1, SEG, dtype=torch.float32)
    batch.view(-1)[:total] = waveform[0]
    valid = torch.full((n,), SEG, dtype=torch.int32)
    if total % SEG:
        valid[-1] = total % SEG
    return batch, valid


@torch.no_grad()
def generate_watermarked_audio(input_file, generator, output_file=None, message_bits=16, device="cuda",
                               messages: Optional[Sequence[int]] = None):
    """py/main16.py:977-1066.  `messages` (optional, one id per 1 s segment) replaces the
    reference's per-segment torch.randint draws (:1001) for reproducible embedding."""
    generator.eval()
    waveform = _prepare(input_file, device)
    total = waveform.shape[1]
    batch, _ = segment(waveform)
    n = batch.shape[0]
    if n == 0:
        raise ValueError("empty audio")
    if messages is None:
        # one draw per segment on `device`, exactly the reference's RNG consumption (:1001, :1017)
        msg = torch.cat([torch.randint(0, 2 ** message_bits, (1,), device=device) for _ in range(n)])
    else:
        msg = torch.as_tensor(list(messages), dtype=torch.int64, device=device)
        if msg.shape != (n,):
            raise ValueError(f"messages: expected {n} ids (one per segment), got {tuple(msg.shape)}")
    seg_dev = batch.to(device, non_blocking=True)
    delta = generator(seg_dev, msg)                       # (n,1,16000): raw delta, no fir/clamp/rms (:1005)
    wm = seg_dev + delta
    both = torch.stack([wm.reshape(-1)[:total], delta.reshape(-1)[:total]]).cpu()
    watermarked_waveform, delta_waveform = both[0:1], both[1:2]
    original_waveform = waveform

    watermark_rms = torch.sqrt((delta_waveform ** 2).mean()).item()
    s0 = original_waveform - original_waveform.mean(dim=1, keepdim=True)
    s1 = watermarked_waveform - watermarked_waveform.mean(dim=1, keepdim=True)
    alpha = (s0 * s1).sum(dim=1, keepdim=True) / ((s0 ** 2).sum(dim=1, keepdim=True) + 1e-8)
    target = alpha * s0
    noise = s1 - target
    si_snr = (10 * torch.log10((target ** 2).sum(dim=1) / ((noise ** 2).sum(dim=1) + 1e-8))).mean().item()
    power_ratio_db = 10 * np.log10(torch.mean(original_waveform ** 2).item() / torch.mean(delta_waveform ** 2).item())

    if output_file:
        out_dir = os.path.dirname(output_file)
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
        save_audio(output_file, watermarked_waveform, SAMPLE_RATE)

    return {"watermarked_waveform": watermarked_waveform, "delta_waveform": delta_waveform,
            "original_waveform": original_waveform,
            "metrics": {"watermark_rms": watermark_rms, "si_snr_db": si_snr, "power_ratio_db": power_ratio_db},
            "messages": msg.cpu()}


@torch.no_grad()
def detect_watermark(input_file, detector, detection_threshold=0.5, visualize=True, device="cuda"):
    """py/main16.py:1114-1207."""
    detector.eval()
    waveform = _prepare(input_file, device)
    total = waveform.shape[1]
    batch, valid = segment(waveform)
    if batch.shape[0] == 0:
        raise ValueError("empty audio")
    r = detector.detect(batch.to(device, non_blocking=True), valid.to(device), want_probs=True, want_votes=False)
    temporal_probs = r["probs"].reshape(-1)[:total].cpu().numpy()
    mean_prob = float(torch.from_numpy(temporal_probs).mean().item())      # over every sample (:1170-1171)
    is_watermarked = mean_prob > detection_threshold                        # strict > (:1173)
    result = {"mean_probability": mean_prob, "is_watermarked": is_watermarked, "temporal_probs": temporal_probs,
              "decision": "WATERMARKED" if is_watermarked else "NOT WATERMARKED"}
    if getattr(detector, "message_bits", 0) > 0:
        mean_logits = r["msg_logits"].cpu().mean(dim=0)                     # mean of per-segment means (:1184)
        result["predicted_message"] = (mean_logits > 0).int().tolist()
        result["message_confidence"] = torch.sigmoid(mean_logits).tolist()
    if visualize:
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            plt = None
        if plt is not None:
            name = os.path.basename(input_file) if isinstance(input_file, str) else "<tensor>"
            t = np.linspace(0, len(temporal_probs) / SAMPLE_RATE, len(temporal_probs))
            plt.figure(figsize=(12, 6))
            plt.plot(t, temporal_probs, label="Detection Probability", alpha=0.7)
            plt.axhline(y=detection_threshold, linestyle="--", label=f"Threshold ({detection_threshold})")
            plt.axhline(y=mean_prob, linestyle="-.", label=f"Mean Probability ({mean_prob:.4f})")
            plt.xlabel("Time (seconds)"); plt.ylabel("Watermark Detection Probability")
            plt.title(f"Watermark Detection Results for {name}\nDecision: {result['decision']}")
            plt.ylim(-0.05, 1.05); plt.legend(); plt.grid(True, alpha=0.3); plt.tight_layout(); plt.show()
    return result


@torch.no_grad()
def detect_prob(path, detector, device="cuda") -> float:
    """Clip-level probability of a file (py/main16.py:1575-1596)."""
    return detect_watermark(path, detector, visualize=False, device=device)["mean_probability"]


# ---- the remaining file-level callers of py/main16.py, on the batched path ------------------------------------------
def set_seed(seed: int = 42) -> None:
    """py/main16.py:21-25."""
    import random
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


@torch.no_grad()
def process_audio_file_with_delta(file_path, generator, sample_rate: int = SAMPLE_RATE, message_bits: int = 16,
                                  device="cuda", messages: Optional[Sequence[int]] = None):
    """py/main16.py:723-762: (watermarked (1,N), delta (1,N), original (1,N)) CPU tensors; every 1 s segment (the last
    one zero-padded and cropped) gets its own message, drawn as the reference draws it unless `messages` is given.
    All segments run as one generator batch."""
    if sample_rate != SAMPLE_RATE:
        raise ValueError(f"the model works at {SAMPLE_RATE} Hz (got sample_rate={sample_rate})")
    r = generate_watermarked_audio(file_path, generator, None, message_bits, device, messages)
    return r["watermarked_waveform"], r["delta_waveform"], r["original_waveform"]


@torch.no_grad()
def run_inference_on_file(file_path, generator, detector, device="cuda", messages: Optional[Sequence[int]] = None):
    """py/main16.py:775-800: embed per segment, then ONE detector pass over the whole watermarked recording (not per
    segment — the reference feeds (1,1,N)), mean detection probability, watermark RMS and SI-SNR.
    Returns (watermarked (1,N), detection_prob, watermark_rms, si_snr)."""
    from .audio import compute_si_snr
    detector.eval()
    wm, delta, orig = process_audio_file_with_delta(file_path, generator, SAMPLE_RATE, generator.message_bits or 16,
                                                    device, messages)
    x = wm.to(device).reshape(1, 1, -1)
    r = detector.detect(x, want_probs=False, want_votes=False)
    detection_prob = float(r["clip_prob"][0])
    watermark_rms = torch.sqrt((delta ** 2).mean()).item()
    si_snr_val = compute_si_snr(orig.to(device), wm.to(device))
    return wm, detection_prob, watermark_rms, si_snr_val


@torch.no_grad()
def evaluate_unseen_file(filepath, generator, detector, device="cuda", messages: Optional[Sequence[int]] = None):
    """py/main16.py:1263-1299: per zero-padded 1 s segment, detection probability of the clean and of the watermarked
    segment, SI-SNR and watermark RMS; returns the four means over the segments (clean, watermarked, si_snr, rms), or
    four Nones when the file cannot be read.  Segments run as one batch through generator and detector."""
    from .audio import file_metrics
    try:
        waveform = _prepare(filepath, device)
    except Exception:
        return None, None, None, None
    generator.eval()
    detector.eval()
    batch, _ = segment(waveform)
    n = batch.shape[0]
    if n == 0:
        return None, None, None, None
    bits = generator.message_bits or 16
    if messages is None:
        msg = torch.cat([torch.randint(0, 2 ** bits, (1,), device=device) for _ in range(n)])     # :1285, one per segment
    else:
        msg = torch.as_tensor(list(messages), dtype=torch.int64, device=device)
    seg = batch.to(device)
    delta = generator(seg, msg)
    seg_w = seg + delta
    p_clean = detector.detect(seg, want_probs=False, want_votes=False)["clip_prob"]
    p_wm = detector.detect(seg_w, want_probs=False, want_votes=False)["clip_prob"]
    m = file_metrics(seg[:, 0], seg_w[:, 0])                    # per segment: rms, si_snr, power ratio
    return (float(p_clean.mean()), float(p_wm.mean()), float(m[:, 1].mean()), float(m[:, 0].mean()))
