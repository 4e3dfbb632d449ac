"""BASELINE config 5: a 10-hour 16 kHz stream (36 000 one-second segments) through embed_detect_stream — host source,
H2D / D2H inside the timed region, segments sharded contiguously over the ranks (no data-path collective; the file
aggregates are one all_reduce of <= 20 floats).

    python tools/stream_bench.py [--hours 10]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29621 \
        tools/stream_bench.py

Prints one JSON line on rank 0: wall seconds (max over ranks) of the whole call — pinning, staging and the tail
handling included — and of the pipeline pass alone (CUDA events inside the call are not exposed, so the second
figure is a repeat call on already-pinned buffers)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import wmb200
from wmb200 import stream as ST


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hours", type=float, default=10.0)
    ap.add_argument("--unpinned", action="store_true", help="pageable source: the driver stages it through its cached pinned buffer")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    g = wmb200.Generator(message_bits=16).cuda().eval()
    d = wmb200.Detector(message_bits=16).cuda().eval()
    n = int(a.hours * 3600 * 16000) - 4321                      # a ragged tail segment
    gen = torch.Generator().manual_seed(7)
    x = torch.empty(n, pin_memory=not a.unpinned)      # host-pinned source (read in place)
    for i in range(0, n, 16_000_000):                            # 0.1 * randn in slabs (host RAM friendly)
        x[i:i + 16_000_000] = 0.1 * torch.randn(min(16_000_000, n - i), generator=gen)
    times, bufs = [], {}
    for rep in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ST.embed_detect_stream(g, d, x, rank=rank, world=world, buffers=bufs)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt))
    lo, hi = out["segment_range"]
    if rank == 0:
        segs = (n + 15999) // 16000
        print(json.dumps({"workload": "config 5: %.1f h stream = %d segments, host source, sharded over %d GPU(s)" % (a.hours, segs, world),
                          "n_gpus": world, "segments": segs, "segments_this_rank": hi - lo,
                          "wall_s_first_call": round(times[0], 3), "wall_s_steady": round(times[-1], 3),
                          "clip_s_per_s_wall": round(segs / times[-1], 1), "realtime_factor": round(a.hours * 3600 / times[-1], 1),
                          "mean_probability": float(out["mean_probability"])}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
