"""Callers of the embed+detect path at scale (SURVEY.md §8f-1 and BASELINE config 5):

* `embed_detect_stream`  — a long recording (hours) cut into 1 s segments, embedded and detected through the
  host-fed pipeline (`wm_embed_detect_host`: sub-batched kernels, H2D/D2H on copy streams), sharded by
  contiguous segment range over the ranks of a `torch.distributed` job.  No data-path collective: the only
  exchange is the handful of file-level sums of `sharding.reduce_file_stats`.
* `process_folder_with_tqdm` — the reference's folder driver (py/main16.py:1409-1446) with its signature and
  output tree, except that the segments of MANY files share one generator batch instead of one launch per
  second of audio; `detect_watermark_folder` is the detection twin (py/main14d.py:1066-1080).

Tail handling follows py/main16.py:1011-1026 / 1152-1168: the last partial segment is right-zero-padded,
processed, and cropped; message-logit means use only its valid samples.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import api, ops
from .functional import SAMPLE_RATE, fir_taps_on
from .sharding import reduce_file_stats, shard_range

SEG = SAMPLE_RATE
AUDIO_EXTS = (".wav", ".mp3", ".flac", ".ogg", ".m4a", ".aac")      # py/main16.py:1413


def _pin(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_pinned() else t.pin_memory()


_STAGING: dict = {}


def _staging(name: str, shape, dtype) -> torch.Tensor:
    """Pinned input staging buffer, cached per (name, dtype) and grown on demand (free_stream_buffers() drops it)."""
    n = 1
    for d in shape:
        n *= d
    buf = _STAGING.get((name, dtype))
    if buf is None or buf.numel() < n:
        buf = torch.empty(n, dtype=dtype, pin_memory=True)
        _STAGING[(name, dtype)] = buf
    return buf[:n].view(shape)


def _result(buffers: dict, name: str, shape) -> torch.Tensor:
    t = buffers.get(name)
    if t is None or tuple(t.shape) != tuple(shape) or not t.is_pinned():
        t = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        buffers[name] = t
    return t


def free_stream_buffers() -> None:
    _STAGING.clear()


@torch.no_grad()
def embed_detect_stream(generator, detector, waveform: torch.Tensor, messages: Optional[torch.Tensor] = None,
                        postprocess: bool = True, chunk: Optional[int] = None, rank: int = 0, world: int = 1,
                        group=None, reduce: bool = True, buffers: Optional[dict] = None) -> dict:
    """waveform: host fp32 (N,) or (1,N) at 16 kHz.  Rank `rank` of `world` embeds and detects segments
    [lo, hi) = shard_range(n_segments, rank, world) and returns, for ITS range, host tensors
    `watermarked` (samples,), `probs` (samples,), `clip_prob` (segments,), `msg_logits` (segments, bits),
    `messages` (segments,), `segment_range`; plus file-level `mean_probability` / `mean_msg_logits`
    (all-reduced over the job when torch.distributed is initialised and `reduce`).  The big results are pinned host
    tensors; pass the same `buffers` dict to successive calls to have them reused (and overwritten) instead of
    pinned afresh."""
    if generator.training or detector.training:
        raise NotImplementedError("embed_detect_stream is the eval-mode path; call .eval() on both modules")
    x = waveform.reshape(-1).to("cpu", torch.float32)
    total = x.numel()
    n_seg = (total + SEG - 1) // SEG
    lo, hi = shard_range(n_seg, rank, world)
    nb = hi - lo
    dev = next(generator.parameters()).device
    bits = detector.message_bits
    if messages is None:
        g = torch.Generator().manual_seed(1234)
        messages = torch.randint(0, 2 ** max(generator.message_bits, 1), (n_seg,), generator=g)
    messages = messages.to("cpu", torch.int64)
    if messages.shape != (n_seg,):
        raise ValueError(f"messages: expected ({n_seg},), got {tuple(messages.shape)}")
    out = {"segment_range": (lo, hi), "messages": messages[lo:hi].clone()}
    s0, s1 = lo * SEG, min(hi * SEG, total)
    if nb == 0:
        out.update(watermarked=torch.empty(0), probs=torch.empty(0), clip_prob=torch.empty(0),
                   msg_logits=torch.empty(0, bits))
        sum_prob, n_samp, sum_ml = 0.0, 0, torch.zeros(bits, device=dev)
    else:
        # host staging: the input buffer is cached between calls (pinning gigabytes costs more than the GPU pass);
        # result buffers are the caller's — fresh pinned tensors unless `buffers` hands reusable ones in
        if x.is_pinned() and x.is_contiguous():
            hs = x[s0:s1]                     # straight out of the caller's pinned recording: the ragged tail of the last
        else:                                 # segment is zero-filled on the device (wm_embed_detect_host_ragged)
            hs = _staging("hs", (nb, SEG), torch.float32)
            hs.view(-1)[:s1 - s0] = x[s0:s1]
            if s1 - s0 < nb * SEG:
                hs.view(-1)[s1 - s0:] = 0.0
        hm = _staging("hm", (nb,), torch.int64)
        hm.copy_(messages[lo:hi])
        buffers = buffers if buffers is not None else {}
        h_sw = _result(buffers, "watermarked", (nb, SEG))
        h_pr = _result(buffers, "probs", (nb, SEG))
        h_cp = _result(buffers, "clip_prob", (nb,))
        h_ml = _result(buffers, "msg_logits", (nb, max(bits, 1)))
        use_msg = generator.message_bits > 0
        pipe = ops.HostPipeline(generator.packed(), generator.embedding_table() if use_msg else None,
                                detector.packed(), fir_taps_on(dev), detector.nout, SEG,
                                chunk=min(nb, chunk or ops.max_chunk()),
                                post_mode=L.POST_ALL if postprocess else 0, device=dev)
        with torch.cuda.device(dev):
            pipe(hs, hm if use_msg else None, h_sw, h_pr, h_cp, h_ml if bits else None)
            torch.cuda.current_stream().synchronize()
        probs = h_pr.view(-1)[:s1 - s0]
        ml = h_ml[:, :bits].clone() if bits else torch.empty(nb, 0)
        tail = s1 - s0 - (nb - 1) * SEG
        if tail < SEG:
            # The reference crops the watermarked tail to its true length (py/main16.py:1011-1026) and detects on
            # that, zero-padded again, averaging message logits over the valid samples only (:1152-1168): redo the
            # detection of that one segment the same way.
            seg = h_sw[nb - 1].clone()
            seg[tail:] = 0.0
            v = torch.tensor([tail], dtype=torch.int32, device=dev)
            r = detector.detect(seg.view(1, 1, SEG).to(dev), v, want_probs=True, want_votes=False)
            h_pr[nb - 1] = r["probs"][0].cpu()
            if bits:
                ml[nb - 1] = r["msg_logits"][0].cpu()
            h_cp[nb - 1] = r["clip_prob"][0].cpu()
        out.update(watermarked=h_sw.view(-1)[:s1 - s0], probs=probs, clip_prob=h_cp, msg_logits=ml)
        # file-level probability sum from the per-segment means the detector epilogue already produced (every
        # segment has SEG valid samples except the cropped tail): no second pass over gigabytes of host memory
        w = torch.full((nb,), float(SEG), dtype=torch.float64)
        w[nb - 1] = float(tail)
        sum_prob, n_samp, sum_ml = float((h_cp.double() * w).sum()), int(probs.numel()), ml.sum(0).to(dev)
    if reduce:
        mp, mlg = reduce_file_stats(sum_prob, n_samp, sum_ml, nb, group)
    else:
        mp, mlg = sum_prob / max(n_samp, 1), (sum_ml / max(nb, 1)).float()
    out["mean_probability"], out["mean_msg_logits"] = mp, mlg.cpu()
    return out


def list_audio_files(input_folder: str, output_root: str, prefix: str = "watermarked_") -> List[Tuple[str, str]]:
    """(in, out) pairs in os.walk order, output tree mirrored (py/main16.py:1415-1425)."""
    pairs = []
    for root, _, files in os.walk(input_folder):
        rel = os.path.relpath(root, input_folder)
        for fname in files:
            if fname.lower().endswith(AUDIO_EXTS):
                pairs.append((os.path.join(root, fname), os.path.join(output_root, rel, f"{prefix}{fname}")))
    return pairs


def plan_batches(seg_counts: Sequence[int], max_clips: int) -> List[Tuple[int, int]]:
    """Greedy grouping of consecutive files into generator batches of at most `max_clips` segments (a single
    longer file gets a batch of its own): [(first_file, last_file_exclusive), ...]."""
    batches, start, acc = [], 0, 0
    for i, n in enumerate(seg_counts):
        if acc and acc + n > max_clips:
            batches.append((start, i))
            start, acc = i, 0
        acc += n
    if start < len(seg_counts):
        batches.append((start, len(seg_counts)))
    return batches


def _bar(total: int, quiet: bool):
    if quiet:
        return None
    try:
        from tqdm import tqdm
        return tqdm(total=total, desc="Watermarking audio files", unit="file")
    except ImportError:
        return None


@torch.no_grad()
def process_folder_with_tqdm(input_folder, generator, message_bits=16, device="cuda", max_clips: int = 4096,
                             rank: int = 0, world: int = 1, loader: Optional[Callable] = None,
                             saver: Optional[Callable] = None, quiet: bool = False) -> dict:
    """py/main16.py:1409-1446: watermark every audio file under `input_folder` into the sibling tree
    `watermarked_<folder>/…/watermarked_<name>`, print the average watermark RMS and power ratio.
    Segments of consecutive files are batched together (<= max_clips per generator call); with `world` > 1
    rank r handles files r, r + world, … (no collective; averages are per rank unless the caller reduces)."""
    generator.eval()
    base = os.path.basename(os.path.abspath(input_folder))
    output_root = os.path.join(os.path.dirname(input_folder), f"watermarked_{base}")
    pairs = list_audio_files(input_folder, output_root)[rank::world]
    load = loader or (lambda path: api._prepare(path, device))
    save = saver or api.save_audio
    waves = [load(p) for p, _ in pairs]
    counts = [(w.shape[1] + SEG - 1) // SEG for w in waves]
    rms_list, pr_list = [], []
    bar = _bar(len(pairs), quiet)
    for f0, f1 in plan_batches(counts, max_clips):
        segs = [api.segment(waves[i])[0] for i in range(f0, f1)]
        batch = torch.cat(segs, 0)
        if batch.shape[0] == 0:
            continue
        # one draw per segment on `device`, file by file, the reference's RNG consumption (:1386, :1394)
        msg = torch.cat([torch.randint(0, 2 ** message_bits, (1,), device=device) for _ in range(batch.shape[0])])
        x = batch.to(device, non_blocking=True)
        delta = generator(x, msg)
        wm_all, d_all = (x + delta).reshape(-1).cpu(), delta.reshape(-1).cpu()
        off = 0
        for i in range(f0, f1):
            total = waves[i].shape[1]
            wm = wm_all[off * SEG: off * SEG + total].unsqueeze(0)
            dw = d_all[off * SEG: off * SEG + total]
            off += counts[i]
            os.makedirs(os.path.dirname(pairs[i][1]) or ".", exist_ok=True)
            save(pairs[i][1], wm, SAMPLE_RATE)
            rms_list.append(float(torch.sqrt((dw ** 2).mean())))
            pr_list.append(float(10 * np.log10(float((waves[i] ** 2).mean()) / max(float((dw ** 2).mean()), 1e-30))))
            if bar is not None:
                bar.update(1)
    if bar is not None:
        bar.close()
    count = len(rms_list)
    avg_rms = float(np.mean(rms_list)) if count else 0.0
    avg_pr = float(np.mean(pr_list)) if count else 0.0
    if not quiet:
        print(f"\nProcessed {count} files")
        print(f"Average Watermark RMS:        {avg_rms:.6f}")
        print(f"Average Power Ratio (dB):     {avg_pr:.2f}")
    return {"files": count, "avg_watermark_rms": avg_rms, "avg_power_ratio_db": avg_pr, "output_root": output_root,
            "outputs": [o for _, o in pairs]}


@torch.no_grad()
def detect_watermark_folder(input_folder, detector, detection_threshold=0.5, device="cuda", max_clips: int = 4096,
                            rank: int = 0, world: int = 1, loader: Optional[Callable] = None) -> List[dict]:
    """Detection over a folder (py/main14d.py:1066-1080 shape of result: one dict per file with the keys of
    detect_watermark minus the plot), segments of consecutive files batched together."""
    detector.eval()
    pairs = list_audio_files(input_folder, input_folder, prefix="")[rank::world]
    load = loader or (lambda path: api._prepare(path, device))
    waves = [load(p) for p, _ in pairs]
    counts = [(w.shape[1] + SEG - 1) // SEG for w in waves]
    results = []
    for f0, f1 in plan_batches(counts, max_clips):
        sv = [api.segment(waves[i]) for i in range(f0, f1)]
        batch, valid = torch.cat([s for s, _ in sv], 0), torch.cat([v for _, v in sv], 0)
        if batch.shape[0] == 0:
            continue
        r = detector.detect(batch.to(device, non_blocking=True), valid.to(device), want_probs=True, want_votes=False)
        probs, ml = r["probs"].reshape(-1).cpu(), r["msg_logits"].cpu()
        off = 0
        for i in range(f0, f1):
            total, n = waves[i].shape[1], counts[i]
            tp = probs[off * SEG: off * SEG + total].numpy()
            mean_prob = float(torch.from_numpy(tp).mean()) if total else 0.0
            res = {"file": pairs[i][0], "mean_probability": mean_prob, "is_watermarked": mean_prob > detection_threshold,
                   "temporal_probs": tp, "decision": "WATERMARKED" if mean_prob > detection_threshold else "NOT WATERMARKED"}
            if getattr(detector, "message_bits", 0) > 0 and n:
                mlg = ml[off: off + n].mean(dim=0)
                res["predicted_message"] = (mlg > 0).int().tolist()
                res["message_confidence"] = torch.sigmoid(mlg).tolist()
            results.append(res)
            off += n
    return results
