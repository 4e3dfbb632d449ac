#!/usr/bin/env python
"""bench.py — clip-seconds/sec of embed+detect (1 s @ 16 kHz) on N B200s.

One "step" = one pass of the hot path over one batch of synthetic clips:
    G -> fir/clamp/rms -> s + delta -> D -> sigmoid(ch 0), clip mean, message-logit means
(the forward of the reference's evaluate_model, py/main16.py:378-398).
Workload = BASELINE.json configs[1]: main16 embed+detect, batch 4096 clips, 16-bit message,
per GPU (weak scaling: every rank owns its own 4096 clips, no data-path collective).

  value   device-resident inputs, CUDA events, max over ranks
  e2e     the same through wm_embed_detect_host with pinned HOST buffers (H2D + D2H inside)
  roofline   the dominant kernel (the fused ResBlock, 5 launches = ~55 % of a step) timed alone with CUDA events
  cpu_baseline  the oracle (torch CPU fp32 restatement of the reference) on a bounded sample

  parity     (rank 0, outside every timed region) 64 clips of the workload against the oracle: delta / probability
             errors, message-bit mismatches (all, and where the sign is decidable), vote-bit mismatches
  aux        informational numbers of the other BASELINE configs, measured after the headline: the majority-vote
             variant of the step, config 3 (main14b_2), config 4 (training step, NCCL gradient all-reduce over the
             ranks of this job), config 5 (10 h stream from host memory, sharded over the ranks)

`--impl reference` times the reference's own CPU implementation of the path on the host cores: the reference's
classes and functions themselves (`oracle/_ref/ref_main16.py`, lifted from the mount by `oracle/make_ref.py` at
build time; kind "reference"), else the oracle port (kind "port") — the notebook export cannot be imported
(SURVEY.md §8c).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T = 16000
FLOP_PER_CLIP = 5.964e9          # G 4.342 + D 1.622 GFLOP
CONV64_K3_FLOP_PER_CLIP = 2 * 64 * 64 * 3 * T      # one 64->64 k3 convolution
RESBLOCK_FLOP_PER_CLIP = 2 * CONV64_K3_FLOP_PER_CLIP   # 786.4 MFLOP: the two convolutions of a ResBlock
RESBLOCK_HBM_BYTES_PER_CLIP = 2 * 64 * T * 4       # planar bf16 hi+lo in, same out: 8.19 MB


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs: ONE `nvidia-smi -lms 50` process whose
    CSV lines are read as they come (spawning nvidia-smi per sample gives only a few samples per second)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], threading.Event(), None
        self.t_first = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.strip().split(",")]
                if len(parts) == 6:
                    if self.t_first is None:
                        self.t_first = time.perf_counter()
                    self.rows.append(parts)
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass
        finally:
            if self.proc is not None:
                try:
                    self.proc.kill()        # the exact process this object started
                except Exception:
                    pass

    def wait_ready(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.t_first is None and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        """Index of the next sample: samples from here on belong to the timed region."""
        return len(self.rows)

    def summary(self, first=0, last=None):
        rows = self.rows[first:last] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        reasons = []
        for i, name in ((2, "hw_slowdown"), (3, "hw_thermal_slowdown"), (4, "sw_thermal_slowdown"), (5, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_mhz_min": sm[0], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}


def build_models(device):
    import torch

    import wmb200
    torch.manual_seed(1234)
    gen = wmb200.Generator(16)                      # random init: models/generator_best.pth is not in the mount
    det = wmb200.Detector(16)
    det.load_state_dict(torch.load(os.path.join(ROOT, "tests", "golden", "detector_best.pth")))
    return gen.to(device).eval(), det.to(device).eval()


def synth(B, seed, device, pin=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    s = (0.1 * torch.randn(B, T, generator=g)).clamp_(-0.99, 0.99)
    m = torch.randint(0, 65536, (B,), generator=g)
    if pin:
        return s.pin_memory(), m.pin_memory()
    return s.to(device), m.to(device)


def _load_reference_module():
    """The reference's own definitions (oracle/_ref/ref_main16.py, see oracle/make_ref.py), or None."""
    import importlib.util
    path = os.path.join(ROOT, "oracle", "_ref", "ref_main16.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_main16", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_rate(sample_B, steps, warmup, threads=None, device="cpu", tf32=False, use_reference=True):
    """The reference's CPU path on `sample_B` clips per step: the reference's own Generator / Detector /
    fir_lowpass / clamp_peak / limit_rms when oracle/_ref holds them (kind "reference"), else the oracle port
    (kind "port").  device="cuda" (informational only, `--ref-device cuda`) runs the same stock-PyTorch eager
    definitions on the GPU.  Returns (clip-s/s, threads, per-step seconds, kind)."""
    import torch

    from oracle import wm_oracle as O
    if device != "cpu":
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
    if threads is None:            # all the host threads we may use (torchrun exports OMP_NUM_THREADS=1)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, threads))
    cores = torch.get_num_threads()
    gen, det = build_models("cpu")
    s, m = synth(sample_B, 1234, device)
    s = s.unsqueeze(1)
    ref = _load_reference_module() if use_reference else None
    if ref is not None:
        rg, rd = ref.Generator(16), ref.Detector(16)
        rg.load_state_dict(gen.state_dict())
        rd.load_state_dict(det.state_dict())
        rg, rd = rg.to(device).eval(), rd.to(device).eval()

        def once():          # forward of the reference's evaluate_model body, py/main16.py:385-398
            delta = ref.limit_rms(ref.clamp_peak(ref.fir_lowpass(rg(s, m))))
            logits = rd(s + delta)
            probs = torch.sigmoid(logits[:, :, 0])
            return {"clip_prob": probs.mean(dim=1), "msg_logits": logits[:, :, 1:].mean(dim=1)}
        kind = "reference"
    else:
        gsd = {k: v.detach().to(device) for k, v in gen.state_dict().items()}
        dsd = {k: v.detach().to(device) for k, v in det.state_dict().items()}

        def once():
            return O.embed_detect(gsd, dsd, s, m)
        kind = "port"
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = once()
            if device != "cpu":
                float(r["clip_prob"][0])         # device -> host read of a result
                torch.cuda.synchronize()
            del r
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sample_B * len(times) / sum(times), cores, times, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = 16 if args.ref_device == "cpu" else args.ref_batch
    rate, cores, times, kind = cpu_reference_rate(sample_B, args.steps, args.warmup, device=args.ref_device,
                                                  tf32=args.ref_tf32)
    sample = (f"{sample_B} clips per step (BASELINE configs[0]) of the {args.batch}-clip workload, torch CPU fp32, " +
              ("the reference's own classes and functions" if kind == "reference" else "oracle port of the reference")
              if args.ref_device == "cpu" else
              f"{sample_B} clips per step, stock PyTorch eager on {args.ref_device}, tf32={args.ref_tf32} (informational)")
    line = {"impl": "reference", "metric": "clip-seconds/sec embed+detect (1 s@16 kHz)", "value": rate,
            "unit": "clip-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"main16 embed+detect, batch {args.batch} clips x 1 s @ 16 kHz, 16-bit message",
                       "sample": sample},
            "cpu_baseline": {"value": rate, "unit": "clip-s/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": "clip-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)



def parity_block(gen, det, s, m, n=64):
    """Parity gates reported with every number: the first `n` clips of the workload through the CUDA
    path and through the oracle (torch CPU fp32 restatement of the reference).  Outside every timed region."""
    import numpy as np
    import torch

    import wmb200
    from oracle import wm_oracle as O          # checker only
    n = min(n, s.shape[0])
    sd, md = s[:n].contiguous(), m[:n].contiguous()
    r = wmb200.embed_detect(gen, det, sd.unsqueeze(1), md, want_delta=True, want_probs=True, want_votes=True)
    torch.cuda.synchronize()
    gsd = {k: v.detach().cpu() for k, v in gen.state_dict().items()}
    dsd = {k: v.detach().cpu() for k, v in det.state_dict().items()}
    with torch.no_grad():
        ref = O.embed_detect(gsd, dsd, sd.cpu().unsqueeze(1), md.cpu())
    err = lambda a, b: float((a.detach().float().cpu().reshape(-1) - b.detach().float().reshape(-1)).abs().max())
    ml, ml_ref = r["msg_logits"].cpu(), ref["msg_logits"]
    e_ml = err(ml, ml_ref)
    bits, bits_ref = (ml > 0).numpy(), (ml_ref > 0).numpy()
    safe = (ml_ref.abs() > 4 * max(e_ml, 1e-6)).numpy()                 # sign decidable at the measured error
    lg_ref = ref["logits"][:, :, 1:]
    e_logit = 2.0 * max(e_ml, err(torch.logit(r["probs"].cpu().clamp(1e-6, 1 - 1e-6)), ref["logits"][:, :, 0]), 1e-6)
    frac_ref = (lg_ref > 0).float().mean(dim=1).numpy()
    slack = (lg_ref.abs() < e_logit).float().mean(dim=1).numpy() + 1.0 / lg_ref.shape[1]
    vote, vote_ref = (r["vote_frac"].cpu().numpy() > 0.5), ref["bits_vote"].numpy()
    decidable = np.abs(frac_ref - 0.5) > slack
    return {"clips": n, "oracle": "oracle/wm_oracle.py (torch CPU fp32)",
            "delta_err": err(r["delta"], ref["delta"]), "prob_err": err(r["probs"], ref["probs"]),
            "clip_prob_err": err(r["clip_prob"], ref["clip_prob"]), "msg_logit_err": e_ml,
            "bits_total": int(bits.size), "bit_mismatches_all": int((bits != bits_ref).sum()),
            "bit_mismatches_safe": int((bits != bits_ref)[safe].sum()), "bits_safe": int(safe.sum()),
            "min_margin": float(ml_ref.abs().min()),
            "vote_mismatches": int((vote != vote_ref).sum()),
            "vote_mismatches_decidable": int((vote != vote_ref)[decidable].sum()), "votes_decidable": int(decidable.sum()),
            "tolerances": {"delta_err": 1e-3, "prob_err": 1e-3, "bits": "exact where decidable"}}


def _event_ms(fn, steps, warmup, world, dev):
    """ms per step of fn(), CUDA events, barrier + synchronize on both sides, max over ranks."""
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def measure_main14b2(dev, world, rank, steps, warmup, B=1024):
    """BASELINE config 3: the main14b_2 residual stack + 2-layer LSTM, 8192 clips sharded over 8 GPUs = 1024 clips
    per GPU per step (weak scaling, no collective), Generator then Detector on s + delta."""
    import torch

    from wmb200 import main14b_2 as M
    from wmb200 import ops
    torch.manual_seed(0)
    G, D = M.Generator().to(dev).eval(), M.Detector().to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    s = (0.1 * torch.randn(B, 1, 16000, device=dev, generator=g)).clamp(-0.99, 0.99)
    msg = torch.randint(0, 65536, (B,), device=dev, generator=g)

    def step():
        with torch.no_grad():
            return D(s + G(s, msg))
    # parity gate (outside the timed region): the tensor-core walk against the layer-by-layer fp32 CUDA operators
    with torch.no_grad():
        d_tc, l_tc = G(s[:4], msg[:4]), D(s[:4])
        old = ops.set_math_mode(0)
        try:
            d_32, l_32 = G(s[:4], msg[:4]), D(s[:4])
        finally:
            ops.set_math_mode(old)
    parity = {"clips": 4, "against": "the fp32 CUDA operators of the same layers (WM_MATH_FP32), pinned on the reference's goldens by tests/test_main14b2.py",
              "delta_rel_err": float((d_tc - d_32).abs().max() / d_32.abs().max().clamp_min(1e-12)),
              "logit_abs_err": float((l_tc - l_32).abs().max()), "tolerances": {"delta_rel_err": 1e-4, "logit_abs_err": 2e-4}}
    tc = ops.get_math_mode() != 0
    n0 = ops.launch_count()
    ms = _event_ms(step, steps, max(warmup, 3), world, dev)
    launches = int(ops.launch_count() - n0) // (steps + max(warmup, 3))
    graphed = None
    if tc:
        # the same pass captured once as a CUDA graph and replayed (static shapes): host cost of a pass = one launch
        try:
            ge = M.GraphedEmbedDetect(G, D, B, 16000, dev)
            out = {}

            def gstep():
                out["r"] = ge(s, msg)
            gms = _event_ms(gstep, steps, max(warmup, 3), world, dev)
            with torch.no_grad():
                same = bool(torch.equal(out["r"][1][:2], D(s[:2] + G(s[:2], msg[:2]))))
            graphed = {"value": world * B * 1000.0 / gms, "unit": "clip-s/s", "ms_per_step": gms,
                       "bit_identical_to_eager": same, "what": "main14b_2.GraphedEmbedDetect: one CUDA-graph replay per pass"}
            del ge, out
        except Exception as e:                                   # informational leg: never fail the line
            graphed = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    return {"metric": "clip-seconds/sec embed+detect (main14b_2 stack)", "value": world * B * 1000.0 / ms,
            "unit": "clip-s/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x2 (fp32 accumulate)" if tc else "f32", "data": "synthetic",
            "config": {"workload": "main14b_2 Generator+Detector, %d clips x 1 s @ 16 kHz per GPU (BASELINE configs[2])" % B,
                       "parallelism": "dp%d" % world,
                       "kernels": "tcgen05 implicit GEMMs over planar bf16-pair activations (pconv_tc_kernel)" if tc
                       else "fp32 CUDA-core operators"},
            "gpu_launches": launches, "graphed": graphed,
            "algorithmic_tflops": 4.540e9 * world * B / ms * 1e3 / 1e12, "parity": parity}


def measure_train(dev, world, rank, steps, warmup, B=16):
    """BASELINE config 4: one train_one_epoch iteration (py/main16.py:238-278) per step through wmb200.Trainer,
    per-GPU batch B, gradients averaged over the ranks with NCCL; time = max over ranks of the CUDA-event time."""
    import torch

    import wmb200
    from wmb200 import ops
    from wmb200 import train as TR
    torch.manual_seed(0)
    tr = TR.Trainer(wmb200.Generator(message_bits=16).to(dev), wmb200.Detector(message_bits=16).to(dev))
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    s = (0.1 * torch.randn(B, 16000, device=dev, generator=g)).clamp(-0.99, 0.99)
    msg = torch.randint(0, 65536, (B,), device=dev, generator=g)
    out = {}

    def step():
        out["r"] = tr.step(s, msg)
    n0 = ops.launch_count()
    ms = _event_ms(step, steps, max(warmup, 3), world, dev)
    return {"metric": "training iterations/sec (main16 train_one_epoch step: forward, backward, Adam)",
            "value": 1000.0 / ms, "unit": "it/s", "clips_per_s": world * B * 1000.0 / ms, "n_gpus": world,
            "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": (1000.0 / ms) / 5.1 if B == 16 and world == 1 else None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "main16 training step, T=16000, batch %d per GPU" % B, "parallelism": "dp%d" % world,
                       "exchange": "NCCL all_reduce of %.1f MB of gradients per step" % ((tr.g_grads.numel() + tr.d_grads.numel()) * 4 / 1e6)},
            "gpu_launches": int(ops.launch_count() - n0), "loss_total": float(out["r"]["total"]),
            "baseline_note": "BASELINE.md: reference 5.1 it/s at B=16 on its own GPU"}


def measure_stream(gen, det, dev, world, rank, hours=10.0):
    """BASELINE config 5: a 10 h 16 kHz stream (36 000 one-second segments, ragged tail) from HOST memory, segments
    sharded contiguously over the ranks (every rank synthesises only its own shard), H2D / D2H inside; wall clock of
    the whole call, max over ranks; first call (pins the buffers) and steady state."""
    import torch
    import torch.distributed as dist

    from wmb200 import stream as ST
    from wmb200.sharding import shard_range
    n = int(hours * 3600 * 16000) - 4321
    segs = (n + 15999) // 16000
    lo, hi = shard_range(segs, rank, world)
    n_local = min(n, hi * 16000) - lo * 16000
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    x = torch.empty(n_local, pin_memory=True)
    for i in range(0, n_local, 64_000_000):
        k = min(64_000_000, n_local - i)
        x[i:i + k].copy_(0.1 * torch.randn(k, device=dev, generator=g))
    torch.cuda.synchronize()
    times, bufs = [], {}
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ST.embed_detect_stream(gen, det, x, rank=0, world=1, reduce=False, buffers=bufs)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt))
    del out, bufs, x
    ST.free_stream_buffers()
    return {"metric": "wall seconds for a %.0f h stream" % hours, "segments": segs, "n_gpus": world,
            "wall_s_first_call": times[0], "wall_s_steady": times[-1], "clip_s_per_s_wall": segs / times[-1],
            "realtime_factor": hours * 3600 / times[-1], "higher_is_better": False, "scaling": "strong",
            "config": {"workload": "10 h = %d segments from pinned host memory, contiguous shards over %d GPU(s) "
                                   "(BASELINE configs[4])" % (segs, world)}}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import wmb200
    from wmb200 import _lib as L
    from wmb200 import ops
    from wmb200.functional import fir_taps_on

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B = args.batch
    gen, det = build_models(dev)
    fir = fir_taps_on(dev)
    s, m = synth(B, 1234 + rank, dev)
    g_blob, d_blob, emb = gen.packed(), det.packed(), gen.embedding_table()

    def step():
        return ops.embed_detect_fwd(g_blob, emb, d_blob, fir, m, s, det.nout, L.POST_ALL, want_delta=False,
                                    want_probs=True, want_votes=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ops.launch_count() - n0

    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_ready()
    marks = {}

    def timed_marked(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        marks["first"] = sampler.mark()
        r = timed(fn, steps, 0)
        marks["last"] = sampler.mark()
        return r

    ms_dev, launches = timed_marked(step, args.steps, args.warmup)
    time.sleep(0.06)                      # let the last in-region sample arrive
    marks["last"] = max(marks["last"], min(sampler.mark(), marks["last"] + 1))
    sampler.stop_flag.set()
    value = world * B * args.steps / (ms_dev / 1e3)

    # ---- end to end: pinned host buffers through the C-ABI host entry point -------------
    hs, hm = synth(B, 4321 + rank, dev, pin=True)
    h_sw = torch.empty(B, T).pin_memory()
    h_pr = torch.empty(B, T).pin_memory()
    h_cp = torch.empty(B).pin_memory()
    h_ml = torch.empty(B, 16).pin_memory()
    pipe = ops.HostPipeline(g_blob, emb, d_blob, fir, det.nout, T, chunk=min(B, ops.max_chunk()))

    def step_e2e():
        pipe(hs, hm, h_sw, h_pr, h_cp, h_ml)
        torch.cuda.current_stream().synchronize()     # the user's result is on the host

    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    h2d = B * T * 4 + B * 8
    d2h = 2 * B * T * 4 + B * 4 + B * 16 * 4
    del pipe, hs, hm, h_sw, h_pr, h_cp, h_ml

    # ---- the evaluate_model variant of the step (py/main16.py:385-398): delta RMS and the per-sample majority vote
    # of the message bits, fused into the detector epilogue (informational, after the headline) ----------------------
    def step_votes():
        return ops.embed_detect_fwd(g_blob, emb, d_blob, fir, m, s, det.nout, L.POST_ALL, want_delta=False,
                                    want_probs=True, want_votes=True, want_rms=True)
    ms_votes, _ = timed(step_votes, max(2, args.steps // 2), 2)
    votes_value = world * B * max(2, args.steps // 2) / (ms_votes / 1e3)

    # ---- roofline of the dominant kernel: the fused ResBlock (5 of the ~18 launches, ~55 % of a step), timed
    # alone with CUDA events on the stream it is launched on ---------------------------------------------------
    Br = min(B, 1024)
    mode = ops.get_math_mode()
    x = torch.randn(Br, T, 64, device=dev)
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    reps = 10
    if mode == L.MATH_FP32:
        y = torch.empty_like(x)
        wk = g_blob[L.G_RB0 + L.RB_W1:L.G_RB0 + L.RB_W1 + 3 * 4096]
        bk = g_blob[L.G_RB0 + L.RB_B1:L.G_RB0 + L.RB_B1 + 64]

        def kernel_once():
            L.check(lib.wm_conv64_fwd(x.data_ptr(), wk.data_ptr(), bk.data_ptr(), None, None, y.data_ptr(), Br, T, 3,
                                      1, st), "wm_conv64_fwd")
        kname, flop = "conv64_fp32_kernel (64->64 k3, CUDA-core FMA)", CONV64_K3_FLOP_PER_CLIP
    else:
        xp = ops.to_planar(x)
        y = torch.empty_like(xp)
        img = g_blob[L.G_TC:L.G_TC + 2 * L.TC_IMG3]
        b1 = g_blob[L.G_RB0 + L.RB_B1:L.G_RB0 + L.RB_B1 + 64]
        b2 = g_blob[L.G_RB0 + L.RB_B2:L.G_RB0 + L.RB_B2 + 64]

        hb1, hb2 = b1.cpu().contiguous(), b2.cpu().contiguous()       # biases by value: the variant the drivers launch

        def kernel_once():
            L.check(lib.wm_resblock_tc_hostbias_fwd(xp.data_ptr(), img.data_ptr(), hb1.data_ptr(), hb2.data_ptr(),
                                                    y.data_ptr(), None, Br, T, st), "wm_resblock_tc_hostbias_fwd")
        kname, flop = "resblock_tc_kernel (fused ResBlock, tcgen05 bf16 pairs)", RESBLOCK_FLOP_PER_CLIP
    ms_k, _ = timed(kernel_once, reps, 3)
    k_tflops = flop * Br * reps / (ms_k / 1e3) / 1e12
    del x, y
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r2_resblock_traffic.json")       # dram bytes of one ncu --set full capture
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "r1_resblock_traffic.json")
    if mode != L.MATH_FP32 and os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj["dram_bytes_per_clip"] * Br, tj["source"]
    roofline = {"bound": "tensor", "kernel": kname, "achieved": k_tflops, "peak": peaks["bf16_burst"],
                "unit": "TFLOP/s", "frac": k_tflops / peaks["bf16_burst"], "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peaks["src"] + " burst (kernel timed alone)",
                "launch_ms": ms_k / reps, "clips_per_launch": Br,
                "algorithmic_flop_per_launch": flop * Br,
                "algorithmic_hbm_bytes_per_launch": RESBLOCK_HBM_BYTES_PER_CLIP * Br,
                "hbm_gbs_algorithmic": RESBLOCK_HBM_BYTES_PER_CLIP * Br * reps / (ms_k / 1e3) / 1e9,
                "issued_flop_factor": 1.0 if mode == L.MATH_FP32 else 3.0,
                "note": "bf16 hi+lo operands, 3 partial products (hi.hi, hi.lo, lo.hi): the tensor pipe issues 3x the "
                        "algorithmic MACs, so frac <= 1/3 by construction; the kernel sits on the shared-memory operand "
                        "bandwidth of its MMAs (128x128x16 + 128x64x16 per tap and k-chunk: 112 cycles for 96 of math, "
                        "DESIGN.md §4)",
                "path_frac_of_sustained": value / world * FLOP_PER_CLIP / (peaks["bf16_sustained"] * 1e12)}

    # ---- parity gates (rank 0) and the other BASELINE configs (all ranks), outside the timed regions ----------------
    parity = None
    if rank == 0 and not args.no_parity:
        try:
            parity = parity_block(gen, det, s, m, 64)
        except Exception as e:                                   # never lose the headline line to a side measurement
            parity = {"error": repr(e)[:300]}
    del s, m
    torch.cuda.empty_cache()
    aux = {"votes_step": {"value": votes_value, "unit": "clip-s/s", "ms_per_step": ms_votes / max(2, args.steps // 2),
                          "what": "the same step with delta RMS and the fused majority-vote bit fractions "
                                  "(evaluate_model, py/main16.py:385-398)"}}
    if not args.no_aux:
        for name, fn in (("config3_main14b2", lambda: measure_main14b2(dev, world, rank, 5, 6)),
                         ("config4_train", lambda: measure_train(dev, world, rank, 5, 3, args.train_batch)),
                         ("config5_stream_10h", lambda: measure_stream(gen, det, dev, world, rank, args.stream_hours))):
            try:
                aux[name] = fn()
            except Exception as e:
                aux[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
    if rank == 0:
        cpu_rate, cores, _, _ = (cpu_reference_rate(16, 3, 1, use_reference=False) if not args.no_cpu_baseline
                                 else (None, 0, [], "port"))
        clocks = sampler.summary(marks.get("first", 0), marks.get("last"))
        line = {"metric": "clip-seconds/sec embed+detect (1 s@16 kHz)", "value": value, "unit": "clip-s/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if mode == L.MATH_FP32 else "bf16x2 (fp32 accumulate)", "data": "synthetic",
                "config": {"workload": f"main16 embed+detect, batch {B} clips x 1 s @ 16 kHz per GPU, 16-bit message "
                                       "(BASELINE configs[1])",
                           "weights": "Generator random init seed 1234 (checkpoint absent), shipped detector_best.pth",
                           "l2": "inputs (262 MB/step) and activations larger than L2",
                           "outputs": "s_w, per-sample probabilities, clip means, 16 mean message logits per clip "
                                      "(SURVEY.md §8d's unit; the majority-vote variant is aux.votes_step)",
                           "math_mode": mode, "chunk": ops.max_chunk()},
                "e2e": {"value": e2e, "unit": "clip-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "parity": parity, "aux": aux,
                "cpu_baseline": {"value": cpu_rate, "unit": "clip-s/s", "cores": cores, "kind": "port",
                                 "sample": "16 clips per pass (BASELINE configs[0]), best-effort torch CPU fp32, 3 passes"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_side(args):
    """`--workload train|main14b2|stream`: one of the aux measurements alone, as its own JSON line (informational; the
    judged line is the default workload)."""
    import torch
    import torch.distributed as dist

    import wmb200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload == "train":
        line = measure_train(dev, world, rank, args.steps, args.warmup, args.train_batch)
    elif args.workload == "main14b2":
        line = measure_main14b2(dev, world, rank, args.steps, args.warmup)
    else:
        gen, det = build_models(dev)
        line = measure_stream(gen, det, dev, world, rank, args.stream_hours)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="clips per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    ap.add_argument("--ref-device", default="cpu", help="--impl reference: cpu (the judged arm) or cuda (informational)")
    ap.add_argument("--ref-batch", type=int, default=256, help="clips per step of --impl reference --ref-device cuda")
    ap.add_argument("--ref-tf32", action="store_true", help="--ref-device cuda: allow TF32 (the reference's setting)")
    ap.add_argument("--workload", default="embed_detect", choices=["embed_detect", "train", "main14b2", "stream"],
                    help="embed_detect = BASELINE's headline metric (default, carries the others as `aux`); train = "
                         "BASELINE config 4, main14b2 = config 3, stream = config 5 alone (informational)")
    ap.add_argument("--no-aux", action="store_true", help="skip the aux block (configs 3, 4, 5)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block")
    ap.add_argument("--stream-hours", type=float, default=10.0, help="length of the config-5 stream")
    ap.add_argument("--train-batch", type=int, default=16, help="--workload train: clips per GPU per iteration")
    args = ap.parse_args()
    if args.workload in ("train", "main14b2", "stream"):
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), __file__,
                   "--workload", args.workload, "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup",
                   str(args.warmup), "--train-batch", str(args.train_batch)]
            sys.exit(subprocess.call(cmd))
        run_side(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), __file__,
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--batch", str(args.batch)]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
