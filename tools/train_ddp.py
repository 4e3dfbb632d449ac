"""Data-parallel training step over NCCL (BASELINE config 4): one process per GPU, per-rank batches,
gradients averaged with all_reduce over NVLink between backward and Adam.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 \
        tools/train_ddp.py [--batch 16] [--steps 5]

Checks (rank 0 prints one JSON line): every rank holds bit-identical parameters after the steps; the averaged
gradient of step 1 equals the mean of the per-rank gradients computed without the exchange; timing = max over ranks
of the CUDA-event time of `steps` iterations."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import wmb200
from wmb200 import train as TR


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)                                    # identical initial weights on every rank
    g, d = wmb200.Generator(message_bits=16).cuda(), wmb200.Detector(message_bits=16).cuda()
    tr = TR.Trainer(g, d)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)          # rank-offset data and messages
    T = 16000
    s = 0.1 * torch.randn(a.batch, T, device="cuda", generator=gen)
    msg = torch.randint(0, 65536, (a.batch,), device="cuda", generator=gen)

    # step 1 by hand: local gradients, then the exchange, against an explicit all_gather mean
    tr.forward_backward(s, msg)
    ok_mean = True
    if world > 1:
        local_g = tr.d_grads.clone()
        gathered = [torch.empty_like(local_g) for _ in range(world)]
        dist.all_gather(gathered, local_g)
        tr.all_reduce_gradients()
        want = torch.stack(gathered).double().mean(0)
        ok_mean = bool(((tr.d_grads.double() - want).abs().max() <= 1e-6 * want.abs().max()).item())
    tr.apply()

    def one():
        tr.step(s, msg)
    for _ in range(2):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        for buf in (tr.g_params, tr.d_params):
            ref = buf.clone()
            dist.broadcast(ref, 0)
            flag = torch.tensor([float(torch.equal(ref, buf))], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            same = same and bool(flag.item())
    if rank == 0:
        print(json.dumps({"n_gpus": world, "batch_per_gpu": a.batch, "ms_per_step": round(float(ms), 2),
                          "it_per_s": round(1000 / float(ms), 2), "clips_per_s": round(world * a.batch * 1000 / float(ms), 1),
                          "params_identical_across_ranks": same, "allreduce_equals_mean": ok_mean,
                          "grad_bytes_per_step": int((tr.g_grads.numel() + tr.d_grads.numel()) * 4)}))
    if world > 1:
        dist.destroy_process_group()
    if not (same and ok_mean):
        sys.exit(1)


if __name__ == "__main__":
    main()
