"""Epoch driver around the training step: One-Cycle schedule against torch's own scheduler, early
stopping semantics (py/main16.py:511-528), and on the GPU: checkpoints in the reference's format that a genuine
torch.optim.Adam can load, and a resumed run that lands bit-for-bit where the uninterrupted one does."""
import os

import numpy as np
import pytest
import torch

from wmb200 import training as TG


@pytest.mark.parametrize("total,pct", [(50, 0.1), (37, 0.3), (200, 0.10)])
def test_one_cycle_matches_torch(total, pct):
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=3e-4 / 25)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=3e-4, total_steps=total, pct_start=pct, div_factor=25,
                                                final_div_factor=1e4, anneal_strategy="cos")
    oc = TG.OneCycle(3e-4, total, pct, 25, 1e4)
    for k in range(total):
        lr, b1 = oc.at(k)
        assert abs(lr - opt.param_groups[0]["lr"]) <= 1e-12 + 1e-9 * lr, k
        assert abs(b1 - opt.param_groups[0]["betas"][0]) <= 1e-9, k
        opt.step()
        if k + 1 < total:
            sched.step()


def test_early_stopping_semantics():
    es = TG.EarlyStopping(patience=2, min_delta=0.1)
    for v, stop in [(1.0, False), (0.95, False), (0.94, True)]:
        es.step(v)
        assert es.early_stop is stop
    es = TG.EarlyStopping(patience=3, min_delta=0.001)
    for v in (1.0, 0.9, 0.8, 0.8, 0.8):
        es.step(v)
    assert not es.early_stop and es.counter == 2 and es.best_loss == 0.8


def _loader(n, B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return [0.1 * torch.randn(B, 1, T, generator=g) for _ in range(n)]


class _Messages:
    """Deterministic per-step messages so an interrupted run can be replayed."""

    def __init__(self):
        self.k = 0

    def __call__(self, B, device):
        g = torch.Generator().manual_seed(1000 + self.k)
        self.k += 1
        return torch.randint(0, 65536, (B,), generator=g).to(device)


@pytest.mark.gpu
def test_checkpoint_resume_is_bit_exact_and_torch_loadable(tmp_path):
    import wmb200
    train, val = _loader(3, 2, 2400, 1), _loader(1, 2, 2400, 2)
    sched = TG.OneCycle(3e-4, 6, 0.34)

    def fresh():
        torch.manual_seed(0)
        return wmb200.Generator(message_bits=16), wmb200.Detector(message_bits=16)

    # uninterrupted: two epochs
    g0, d0 = fresh()
    msgs = _Messages()
    tr0, logs0, _ = TG.fit(g0, d0, train, val, 2, schedule=sched, ckpt_dir=None, log=lambda *_: None, message_fn=msgs,
                           patience=10)
    # interrupted after epoch 1, resumed from ckpt_latest.pth in a new process-like state
    g1, d1 = fresh()
    msgs = _Messages()
    TG.fit(g1, d1, train, val, 1, schedule=sched, ckpt_dir=str(tmp_path), log=lambda *_: None, message_fn=msgs,
           patience=10)
    ck = torch.load(tmp_path / "ckpt_latest.pth", map_location="cpu", weights_only=False)
    assert set(ck) >= {"epoch", "step", "best_val", "gen", "det", "opt", "sched"} and ck["epoch"] == 1 and ck["step"] == 3
    # the reference's own resume path: plain modules + torch.optim.Adam + OneCycleLR accept the file
    g2, d2 = fresh()
    g2.load_state_dict(ck["gen"]); d2.load_state_dict(ck["det"])
    opt = torch.optim.Adam(list(g2.parameters()) + list(d2.parameters()), lr=1e-3)
    opt.load_state_dict(ck["opt"])
    s = torch.optim.lr_scheduler.OneCycleLR(opt, **sched.kwargs())
    s.load_state_dict(ck["sched"])
    assert s.last_epoch == 3 and len(opt.state) == len(list(g2.parameters())) + len(list(d2.parameters()))
    assert float(next(iter(opt.state.values()))["step"]) == 3.0
    # resume with this library
    g3, d3 = fresh()
    tr3, logs3, _ = TG.fit(g3, d3, train, val, 2, schedule=sched, ckpt_dir=str(tmp_path), log=lambda *_: None,
                           message_fn=msgs, patience=10)
    assert len(logs3) == 1 and tr3.steps == 6 == tr0.steps
    assert torch.equal(tr3.g_params, tr0.g_params) and torch.equal(tr3.d_params, tr0.d_params)
    assert torch.equal(tr3.g_stats, tr0.g_stats) and torch.equal(tr3.g_m, tr0.g_m)
    assert logs3[0] == logs0[1]
    assert int(g3.state_dict()["encoder.1.block.1.num_batches_tracked"]) == 6
    assert os.path.exists(tmp_path / "generator_best.pth") and os.path.exists(tmp_path / "ckpt_best.pth")


@pytest.mark.gpu
def test_reference_signature_of_train_one_epoch():
    """The reference's loop body `train_one_epoch(generator, detector, train_loader, optimizer, losses, device)`
    (py/main16.py:538) runs unchanged and leaves trained weights in the modules."""
    import wmb200
    torch.manual_seed(3)
    g, d = wmb200.Generator(message_bits=16).cuda(), wmb200.Detector(message_bits=16).cuda()
    opt = torch.optim.Adam(list(g.parameters()) + list(d.parameters()), lr=1e-3)
    losses = {"mel": wmb200.MultiScaleMelLoss(), "loud": wmb200.TFLoudnessLoss()}
    loader = _loader(2, 2, 2400, 5)
    w0 = d.state_dict()["model.3.weight"].clone()
    m1 = TG.train_one_epoch(g, d, loader, opt, losses, "cuda")
    m2 = TG.train_one_epoch(g, d, loader, opt, losses, "cuda")
    assert set(m1) == set(TG.LOG_KEYS) and all(np.isfinite(v) for v in m2.values())
    assert g._wmb200_trainer.steps == 4                       # one Trainer across both epochs
    assert float((d.state_dict()["model.3.weight"] - w0).abs().max()) > 1e-4
    assert int(d.state_dict()["model.1.block.1.num_batches_tracked"]) == 4
    v = wmb200.validate_one_epoch(g, d, loader, losses, "cuda")
    assert np.isfinite(v["total"])
