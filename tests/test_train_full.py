"""One whole training iteration (py/main16.py:238-278, SURVEY.md 8a-11 / BASELINE config 4).

CPU: the oracle restatement against vectors produced by the reference's own classes and loop body
(tests/golden/train_full.npz), and the host-side gradient exchange under gloo with two ranks.
GPU: wmb200.Trainer against the same vectors and against the oracle (fp64 and fp32) on seeded inputs — the tolerance
logic is the one described in tests/test_train.py (fp32 ReLU-mask flips; dead biases in front of BatchNorm)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import wm_oracle_train as OT
from tests import helpers as H

G = H.load_npz("train_full.npz")
DEAD = ("block.0.bias", "block.3.bias")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOSS_KEYS = ("l1", "mel", "loud", "loc", "bce", "hf", "total", "raw_total")


def sd_of(prefix):
    sd = {k[len(prefix):]: torch.from_numpy(np.asarray(v)) for k, v in G.items() if k.startswith(prefix)}
    if "embedding.rows" in sd:
        emb = torch.zeros(65536, 64)
        emb[torch.from_numpy(G["message"])] = sd.pop("embedding.rows")
        sd["embedding.weight"] = emb
    return sd


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_oracle_reproduces_the_reference_iteration():
    o = OT.TrainOracle(sd_of("init.g."), sd_of("init.d."))
    r = o.step(torch.from_numpy(G["s"]), torch.from_numpy(G["message"]))
    for i, k in enumerate(LOSS_KEYS):
        assert abs(float(r[k]) - G["losses"][i]) < 1e-5 * max(1.0, abs(G["losses"][i])), k
    assert float((r["s_w"] - torch.from_numpy(G["s_w"])).abs().max()) < 1e-6
    for tag, grads, want in (("g", r["g_grads"], sd_of("grad.g.")), ("d", r["d_grads"], sd_of("grad.d."))):
        gmax = max(float(v.abs().max()) for v in want.values())
        for k, v in want.items():
            if k.endswith(DEAD):
                assert float(grads[k].abs().max()) < 1e-3 * gmax, k
            else:
                assert rel(grads[k], v) < 2e-3, (tag, k)
    gsd, dsd = o.state_dicts()
    for sd, want in ((gsd, sd_of("final.g.")), (dsd, sd_of("final.d."))):
        for k, v in want.items():
            if k.endswith(DEAD) or k.endswith("num_batches_tracked"):
                continue
            assert float((sd[k] - v).abs().max()) < (2.1e-3 if "running" not in k else 1e-5), k
            assert float((sd[k] - v).abs().median()) < 1e-5, k


def test_generator_flat_layout_round_trips():
    from wmb200 import train as TR
    sd = sd_of("init.g.")
    back = TR.unflatten_generator(TR.flatten_generator(sd, "cpu"))
    for k in OT.G_PARAM_KEYS:
        assert torch.equal(back[k], sd[k]), k


_DDP_SCRIPT = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from wmb200.train import average_gradients
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
a, b = torch.full((1000,), float(r + 1)), torch.arange(10.0) * (r + 1)
average_gradients((a, b))
assert torch.allclose(a, torch.full((1000,), 1.5)) and torch.allclose(b, torch.arange(10.0) * 1.5)
dist.destroy_process_group()
print("ok", r)
"""


def test_gradient_exchange_two_ranks_gloo(tmp_path):
    script = tmp_path / "ddp.py"
    script.write_text(_DDP_SCRIPT)
    port = str(29700 + os.getpid() % 200)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


# ---------------------------------------------------------------- GPU ------------------------------------------
def _modules(gsd, dsd):
    import wmb200
    g, d = wmb200.Generator(message_bits=16), wmb200.Detector(message_bits=16)
    g.load_state_dict(gsd)
    d.load_state_dict(dsd)
    return g.cuda(), d.cuda()


def _compare_grads(got, want, base=None, tol=None):
    gmax = max(float(v.abs().max()) for v in want.values())
    for k, v in want.items():
        if k.endswith(DEAD):
            assert float(got[k].abs().max()) < 1e-3 * gmax, k
        elif k == "embedding.weight":
            rows = v.abs().sum(dim=1) > 0
            assert float(got[k][~rows.to(got[k].device)].abs().max()) == 0.0
            lim = tol if base is None else 3 * rel(base[k], v) + 2e-4
            assert rel(got[k], v) < lim, k
        else:
            lim = tol if base is None else 3 * rel(base[k], v) + 2e-4
            assert rel(got[k], v) < lim, (k, rel(got[k], v), lim)


@pytest.mark.gpu
def test_trainer_matches_the_reference_golden():
    from wmb200 import train as TR
    g, d = _modules(sd_of("init.g."), sd_of("init.d."))
    tr = TR.Trainer(g, d)
    out = tr.forward_backward(torch.from_numpy(G["s"]).cuda(), torch.from_numpy(G["message"]).cuda(), want_s_w=True)
    for i, k in enumerate(LOSS_KEYS):
        assert abs(float(out[k]) - G["losses"][i]) < 2e-4 * max(1.0, abs(G["losses"][i])), (k, float(out[k]), G["losses"][i])
    assert float((out["s_w"].cpu() - torch.from_numpy(G["s_w"])).abs().max()) < 2e-5
    gg, dg = tr.grad_dicts()
    _compare_grads(gg, sd_of("grad.g."), tol=5e-2)
    _compare_grads(dg, sd_of("grad.d."), tol=5e-2)
    # the bulk of every gradient is far closer than its worst element
    for got, want in ((gg, sd_of("grad.g.")), (dg, sd_of("grad.d."))):
        for k, v in want.items():
            if not k.endswith(DEAD) and v.numel() >= 64 and k != "embedding.weight":
                dlt = (got[k].cpu() - v).abs()
                assert float(dlt.median()) < 2e-3 * float(v.abs().max()), k
    tr.apply()
    gsd, dsd = tr.state_dicts()
    for sd, want in ((gsd, sd_of("final.g.")), (dsd, sd_of("final.d."))):
        for k, v in want.items():
            if k.endswith(DEAD) or k.endswith("num_batches_tracked"):
                continue
            dlt = (sd[k].cpu() - v).abs()
            if "running" in k:
                assert float(dlt.max()) < 2e-5 * max(1.0, float(v.abs().max())), k
            else:
                assert float(dlt.max()) <= 2.1e-3, k
                assert float(dlt.median()) < 1e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,seed", [(2, 2400, 0), (3, 16000, 1), (1, 2401, 2)])
def test_trainer_matches_oracle(B, T, seed):
    import wmb200
    from wmb200 import train as TR
    torch.manual_seed(seed)
    g, d = wmb200.Generator(message_bits=16), wmb200.Detector(message_bits=16)
    with torch.no_grad():
        for m in list(g.modules()) + list(d.modules()):
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(0.8 + 0.4 * torch.rand(64))
                m.bias.copy_(0.1 * torch.randn(64))
        g.decoder[2].weight.mul_(0.05)
        g.decoder[2].bias.mul_(0.05)
    gsd, dsd = g.state_dict(), d.state_dict()
    tr = TR.Trainer(g.cuda(), d.cuda())
    o64 = OT.TrainOracle(gsd, dsd, dtype=torch.float64, device="cuda")
    o32 = OT.TrainOracle(gsd, dsd, dtype=torch.float32, device="cuda")
    t = torch.arange(T, device="cuda") / 16000.0
    for step in range(2):
        s = 0.1 * torch.randn(B, T, device="cuda") + 0.2 * torch.sin(2 * np.pi * 300.0 * t)
        msg = torch.randint(0, 65536, (B,), device="cuda")
        want, base = o64.step(s, msg), o32.step(s, msg)
        got = tr.step(s, msg)
        for k in LOSS_KEYS:
            w = float(want[k])
            assert abs(float(got[k]) - w) < 3 * abs(float(base[k]) - w) + 2e-5 * max(1.0, abs(w)), (step, k)
        if step == 0:
            gg, dg = tr.grad_dicts()
            _compare_grads(gg, want["g_grads"], base["g_grads"])
            _compare_grads(dg, want["d_grads"], base["d_grads"])
    gsd1, dsd1 = tr.state_dicts()
    wg, wd = o64.state_dicts()
    for sd, want in ((gsd1, wg), (dsd1, wd)):
        for k, v in want.items():
            if k.endswith(DEAD) or k.endswith("num_batches_tracked"):
                continue
            dlt = (sd[k].double() - v).abs()
            if "running" in k:
                assert float(dlt.max()) < 1e-3 * max(1.0, float(v.abs().max())), k
            else:
                assert float(dlt.max()) <= 2 * 2.1e-3, k
                assert float(dlt.median()) < 1e-4, k


@pytest.mark.gpu
def test_training_reduces_the_loss_and_writes_back():
    import wmb200
    from wmb200 import train as TR
    torch.manual_seed(4)
    g, d = wmb200.Generator(message_bits=16).cuda(), wmb200.Detector(message_bits=16).cuda()
    tr = TR.Trainer(g, d)
    s = 0.1 * torch.randn(4, 4000, device="cuda")
    msg = torch.randint(0, 65536, (4,), device="cuda")
    first = {k: float(v) for k, v in tr.step(s, msg).items()}
    for _ in range(7):
        last = {k: float(v) for k, v in tr.step(s, msg).items()}
    assert last["loc"] < first["loc"] and last["total"] < first["total"]
    tr.write_back(g, d)
    g.eval(); d.eval()
    r = wmb200.embed_detect(g, d, s.unsqueeze(1), msg)
    assert torch.isfinite(r["s_w"]).all()


@pytest.mark.gpu
def test_forward_backward_is_deterministic():
    import wmb200
    from wmb200 import train as TR
    torch.manual_seed(6)
    g, d = wmb200.Generator(message_bits=16).cuda(), wmb200.Detector(message_bits=16).cuda()
    tr = TR.Trainer(g, d)
    s = 0.1 * torch.randn(3, 3000, device="cuda")
    msg = torch.tensor([7, 7, 9], device="cuda")            # a repeated message: two clips hit one embedding row
    tr.forward_backward(s, msg)
    g1, d1 = tr.g_grads.clone(), tr.d_grads.clone()
    stats = tr.g_stats.clone()
    tr.forward_backward(s, msg)
    assert torch.equal(tr.g_grads, g1) and torch.equal(tr.d_grads, d1)
    assert not torch.equal(tr.g_stats, stats)               # running stats moved a second time


@pytest.mark.gpu
def test_data_parallel_step_two_gpus():
    """Two ranks over NCCL: averaged gradients equal the mean of the per-rank ones and the replicas stay
    bit-identical (tools/train_ddp.py does the checking and exits non-zero otherwise)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = str(29800 + os.getpid() % 100)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tools", "train_ddp.py"), "--batch", "2", "--steps", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert '"params_identical_across_ranks": true' in r.stdout and '"allreduce_equals_mean": true' in r.stdout


@pytest.mark.gpu
def test_out_of_range_message_is_reported():
    import wmb200
    from wmb200 import train as TR
    torch.manual_seed(8)
    tr = TR.Trainer(wmb200.Generator(message_bits=16).cuda(), wmb200.Detector(message_bits=16).cuda())
    s = 0.1 * torch.randn(2, 2400, device="cuda")
    tr.forward_backward(s, torch.tensor([3, 65535], device="cuda"))
    tr.check_messages()
    tr.forward_backward(s, torch.tensor([3, 65536], device="cuda"))
    with pytest.raises(IndexError, match="out of range"):
        tr.state_dicts()
