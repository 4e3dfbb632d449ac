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

`--impl reference` times the reference's own CPU implementation of the path (the oracle port;
the reference's notebook export cannot be imported — SURVEY.md §8c) on the host cores.
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


def cpu_reference_rate(sample_B, steps, warmup, threads=None, device="cpu", tf32=False):
    """Oracle (port of the reference's CPU path) on `sample_B` clips per step.  device="cuda" (informational only,
    `--ref-device cuda`) runs the same stock-PyTorch eager definitions on the GPU."""
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
    gsd = {k: v.detach().to(device) for k, v in gen.state_dict().items()}
    dsd = {k: v.detach().to(device) for k, v in det.state_dict().items()}
    s, m = synth(sample_B, 1234, device)
    s = s.unsqueeze(1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = O.embed_detect(gsd, dsd, s, m)
            if device != "cpu":
                float(r["clip_prob"][0])         # device -> host read of a result
                torch.cuda.synchronize()
            del r
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sample_B * len(times) / sum(times), cores, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = 16 if args.ref_device == "cpu" else args.ref_batch
    rate, cores, times = cpu_reference_rate(sample_B, args.steps, args.warmup, device=args.ref_device, tf32=args.ref_tf32)
    sample = (f"{sample_B} clips per step (BASELINE configs[0]) of the {args.batch}-clip workload, torch CPU fp32"
              if args.ref_device == "cpu" else
              f"{sample_B} clips per step, stock PyTorch eager on {args.ref_device}, tf32={args.ref_tf32} (informational)")
    line = {"impl": "reference", "metric": "clip-seconds/sec embed+detect (1 s@16 kHz)", "value": rate,
            "unit": "clip-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"main16 embed+detect, batch {args.batch} clips x 1 s @ 16 kHz, 16-bit message",
                       "sample": sample},
            "cpu_baseline": {"value": rate, "unit": "clip-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "clip-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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

        def kernel_once():
            L.check(lib.wm_resblock_tc_fwd(xp.data_ptr(), img.data_ptr(), b1.data_ptr(), b2.data_ptr(), y.data_ptr(),
                                           None, Br, T, st), "wm_resblock_tc_fwd")
        kname, flop = "resblock_tc_kernel (fused ResBlock, tcgen05 bf16 pairs)", RESBLOCK_FLOP_PER_CLIP
    ms_k, _ = timed(kernel_once, reps, 3)
    k_tflops = flop * Br * reps / (ms_k / 1e3) / 1e12
    del x, y
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r1_resblock_traffic.json")       # dram bytes of one ncu --set full capture
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
                "issued_flop_factor": 1.0 if mode == L.MATH_FP32 else 1.5,
                "note": "bf16 hi+lo operands, 3 partial products: the tensor pipe issues 1.5x the algorithmic FLOPs; "
                        "the kernel sits on the shared-memory operand bandwidth of the 128x128x16 MMA (DESIGN.md §4)",
                "path_frac_of_sustained": value / world * FLOP_PER_CLIP / (peaks["bf16_sustained"] * 1e12)}

    if rank == 0:
        cpu_rate, cores, _ = cpu_reference_rate(16, 3, 1) if not args.no_cpu_baseline else (None, 0, [])
        clocks = sampler.summary(marks.get("first", 0), marks.get("last"))
        line = {"metric": "clip-seconds/sec embed+detect (1 s@16 kHz)", "value": value, "unit": "clip-s/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if mode == L.MATH_FP32 else "bf16x2 (fp32 accumulate)", "data": "synthetic",
                "config": {"workload": f"main16 embed+detect, batch {B} clips x 1 s @ 16 kHz per GPU, 16-bit message "
                                       "(BASELINE configs[1])",
                           "weights": "Generator random init seed 1234 (checkpoint absent), shipped detector_best.pth",
                           "l2": "inputs (262 MB/step) and activations larger than L2",
                           "math_mode": mode, "chunk": ops.max_chunk()},
                "e2e": {"value": e2e, "unit": "clip-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
                "cpu_baseline": {"value": cpu_rate, "unit": "clip-s/s", "cores": cores, "kind": "port",
                                 "sample": "16 clips per pass (BASELINE configs[0]), best-effort torch CPU fp32, 3 passes"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """BASELINE config 4 (informational; the judged line is the default workload): one train_one_epoch iteration
    (py/main16.py:238-278) per step through wmb200.Trainer, per-GPU batch --train-batch, gradients averaged over
    the ranks with NCCL; time = max over ranks of the CUDA-event time."""
    import torch
    import torch.distributed as dist

    import wmb200
    from wmb200 import ops
    from wmb200 import train as TR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    tr = TR.Trainer(wmb200.Generator(message_bits=16).to(dev), wmb200.Detector(message_bits=16).to(dev))
    B = args.train_batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    s = (0.1 * torch.randn(B, 16000, device=dev, generator=g)).clamp(-0.99, 0.99)
    msg = torch.randint(0, 65536, (B,), device=dev, generator=g)
    for _ in range(max(args.warmup, 3)):
        tr.step(s, msg)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = tr.step(s, msg)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        print(json.dumps({
            "metric": "training iterations/sec (main16 train_one_epoch step: forward, backward, Adam)",
            "value": 1000.0 / ms, "unit": "it/s", "clips_per_s": world * B * 1000.0 / ms, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": (1000.0 / ms) / 5.1 if B == 16 and world == 1 else None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "main16 training step, T=16000, batch %d per GPU" % B, "parallelism": "dp%d" % world,
                       "exchange": "NCCL all_reduce of %.1f MB of gradients per step" % ((tr.g_grads.numel() + tr.d_grads.numel()) * 4 / 1e6)},
            "gpu_launches": int(ops.launch_count() - n0), "loss_total": float(out["total"]),
            "baseline_note": "BASELINE.md: reference 5.1 it/s at B=16 on its own GPU"}))
    if world > 1:
        dist.destroy_process_group()


def run_main14b2(args):
    """BASELINE config 3 (informational): the main14b_2 residual stack + 2-layer LSTM, 8192 clips sharded over 8 GPUs
    = 1024 clips per GPU per step (weak scaling, no collective), Generator then Detector on s + delta."""
    import torch
    import torch.distributed as dist

    from wmb200 import main14b_2 as M
    from wmb200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    G, D = M.Generator().to(dev).eval(), M.Detector().to(dev).eval()
    B = 1024
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    s = (0.1 * torch.randn(B, 1, 16000, device=dev, generator=g)).clamp(-0.99, 0.99)
    msg = torch.randint(0, 65536, (B,), device=dev, generator=g)

    def step():
        with torch.no_grad():
            return D(s + G(s, msg))

    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        print(json.dumps({
            "metric": "clip-seconds/sec embed+detect (main14b_2 stack)", "value": world * B * 1000.0 / ms,
            "unit": "clip-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "main14b_2 Generator+Detector, 1024 clips x 1 s @ 16 kHz per GPU (BASELINE configs[2])",
                       "parallelism": "dp%d" % world},
            "gpu_launches": int(ops.launch_count() - n0), "algorithmic_tflops": 4.540e9 * world * B / ms * 1e3 / 1e12}))
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
    ap.add_argument("--workload", default="embed_detect", choices=["embed_detect", "train", "main14b2"],
                    help="embed_detect = BASELINE's headline metric (default); train = BASELINE config 4, "
                         "main14b2 = BASELINE config 3 (both informational)")
    ap.add_argument("--train-batch", type=int, default=16, help="--workload train: clips per GPU per iteration")
    args = ap.parse_args()
    if args.workload in ("train", "main14b2"):
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus > 1 and world == 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), __file__,
                   "--workload", args.workload, "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup",
                   str(args.warmup), "--train-batch", str(args.train_batch)]
            sys.exit(subprocess.call(cmd))
        (run_train if args.workload == "train" else run_main14b2)(args)
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
