"""Epoch-level training driver around `train.Trainer`: the reference's `train_one_epoch` loop and
loss bookkeeping (py/main16.py:223-294), `EarlyStopping` (py/main16.py:511-528), the best-model / early-stop loop
(py/main16.py:534-560), and the resumable driver of py/main14d.py:499-625 — One-Cycle learning rate (and Adam beta1)
per optimiser step, `ckpt_latest.pth` / `ckpt_best.pth` in the reference's checkpoint format
({"epoch","step","best_val","gen","det","opt","sched"}, `opt` being a genuine `torch.optim.Adam.state_dict()` so
either side can resume the other's run).
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch

from . import evaluate as _ev
from . import train as TR

LOG_KEYS = ("total", "raw_total", "l1", "mel", "loud", "loc", "bce")      # py/main16.py:228-236


class EarlyStopping:
    """py/main16.py:511-528."""

    def __init__(self, patience: int = 3, min_delta: float = 0.0):
        self.patience = patience
        self.min_delta = min_delta
        self.best_loss = float("inf")
        self.counter = 0
        self.early_stop = False

    def step(self, val_loss: float) -> None:
        if self.best_loss - val_loss > self.min_delta:
            self.best_loss = val_loss
            self.counter = 0
        else:
            self.counter += 1
            if self.counter >= self.patience:
                self.early_stop = True


class OneCycle:
    """torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr, total_steps, pct_start, anneal_strategy="cos",
    div_factor, final_div_factor) with its defaults (two phases, momentum cycled between 0.95 and 0.85 — for Adam
    that is beta1), as configured at py/main14d.py:499-507.  `at(k)` = (lr, beta1) of optimiser step k (0-based)."""

    def __init__(self, max_lr: float = 3e-4, total_steps: int = 1, pct_start: float = 0.10, div_factor: float = 25.0,
                 final_div_factor: float = 1e4, base_momentum: float = 0.85, max_momentum: float = 0.95):
        if total_steps <= 0:
            raise ValueError("Expected positive integer total_steps")
        if not 0 <= pct_start <= 1:
            raise ValueError("Expected float between 0 and 1 pct_start")
        self.max_lr, self.total_steps, self.pct_start = float(max_lr), int(total_steps), float(pct_start)
        self.div_factor, self.final_div_factor = float(div_factor), float(final_div_factor)
        self.base_momentum, self.max_momentum = float(base_momentum), float(max_momentum)
        self.initial_lr = self.max_lr / self.div_factor
        self.min_lr = self.initial_lr / self.final_div_factor

    @staticmethod
    def _cos(start: float, end: float, pct: float) -> float:
        return end + (start - end) / 2.0 * (math.cos(math.pi * pct) + 1.0)

    def at(self, step: int) -> Tuple[float, float]:
        if step > self.total_steps:
            raise ValueError(f"Tried to step {step} times. The specified number of total steps is {self.total_steps}")
        e0 = float(self.pct_start * self.total_steps) - 1.0
        e1 = float(self.total_steps - 1)
        if step <= e0:
            pct = step / e0 if e0 != 0 else 0.0
            return self._cos(self.initial_lr, self.max_lr, pct), self._cos(self.max_momentum, self.base_momentum, pct)
        pct = (step - e0) / (e1 - e0)
        return self._cos(self.max_lr, self.min_lr, pct), self._cos(self.base_momentum, self.max_momentum, pct)

    def kwargs(self) -> dict:
        return dict(max_lr=self.max_lr, total_steps=self.total_steps, pct_start=self.pct_start,
                    div_factor=self.div_factor, final_div_factor=self.final_div_factor, anneal_strategy="cos",
                    base_momentum=self.base_momentum, max_momentum=self.max_momentum)


def _draw_messages(B: int, device, bits: int = 16) -> torch.Tensor:
    return torch.randint(0, 2 ** bits, (B,), device=device)                 # py/main16.py:241


def train_one_epoch(trainer, train_loader: Iterable, *args, schedule: Optional[OneCycle] = None,
                    global_step: int = 0, message_fn: Callable = _draw_messages, progress: Callable = iter) -> Dict[str, float]:
    """One pass over `train_loader` (batches s of shape (B,1,T) or (B,T), any device): the mean of every loss term,
    as py/main16.py:238-294 returns it.  With `schedule`, optimiser step k uses schedule.at(global_step + k).

    Two call forms:
      train_one_epoch(trainer, train_loader, ...)                                        # a wmb200 Trainer
      train_one_epoch(generator, detector, train_loader, optimizer, losses, device)       # the reference's signature
    The second keeps the reference's training loop (py/main16.py:534-560) unchanged: a Trainer is created on the first
    call and kept on the generator module; lr, betas and eps are read from `optimizer.param_groups[0]` on every call
    (the torch optimizer itself is not stepped — Adam's moments live in the Trainer) and an optimizer the fused step
    does not implement (another class, several groups, weight decay, amsgrad) raises; `losses` is accepted and ignored (the
    mel / loudness kernels are built in), and the modules receive the updated parameters and BatchNorm statistics
    before the function returns, so `validate_one_epoch(generator, detector, ...)` sees the trained weights."""
    if not isinstance(trainer, TR.Trainer):
        generator, detector, loader = trainer, train_loader, args[0]
        optimizer = args[1] if len(args) > 1 else None
        device = args[3] if len(args) > 3 else "cuda"
        _check_optimizer(optimizer)
        tr = getattr(generator, "_wmb200_trainer", None)
        if tr is None or tr._detector_ref() is not detector:
            import weakref
            generator.to(device)
            detector.to(device)
            g0 = optimizer.param_groups[0] if optimizer is not None else {}
            tr = TR.Trainer(generator, detector, lr=float(g0.get("lr", TR.LR)),
                            betas=tuple(float(b) for b in g0.get("betas", (0.9, 0.999))), eps=float(g0.get("eps", 1e-8)))
            tr._detector_ref = weakref.ref(detector)
            object.__setattr__(generator, "_wmb200_trainer", tr)
        elif optimizer is not None:
            g0 = optimizer.param_groups[0]
            tr.lr = float(g0["lr"])
            tr.betas, tr.eps = tuple(float(b) for b in g0.get("betas", tr.betas)), float(g0.get("eps", tr.eps))
        out = train_one_epoch(tr, loader, schedule=schedule, global_step=global_step, message_fn=message_fn,
                              progress=progress)
        tr.write_back(generator, detector)
        return out
    sums = torch.zeros(len(LOG_KEYS), device=trainer.device, dtype=torch.float64)
    n = 0
    for s in progress(train_loader):
        s = s.to(trainer.device, non_blocking=True)
        message = message_fn(s.shape[0], trainer.device)
        if schedule is not None:
            lr, beta1 = schedule.at(global_step + n)
            trainer.lr, trainer.betas = lr, (beta1, trainer.betas[1])
        out = trainer.step(s, message)
        sums += torch.stack([out[k].double() for k in LOG_KEYS])           # stays on the device: no sync per batch
        n += 1
    if n == 0:
        raise ValueError("train_one_epoch: the loader produced no batches")
    return {k: float(v) / n for k, v in zip(LOG_KEYS, sums.tolist())}


def _check_optimizer(optimizer) -> None:
    """The Trainer implements torch.optim.Adam as the reference configures it (py/main16.py:504): one parameter group,
    no weight decay, no amsgrad / maximize.  lr, betas and eps are taken from the optimizer; anything else that would
    silently change the update is refused."""
    if optimizer is None:
        return
    if not isinstance(optimizer, torch.optim.Adam) or isinstance(optimizer, torch.optim.AdamW):
        raise TypeError(f"train_one_epoch: only torch.optim.Adam is implemented (got {type(optimizer).__name__}); "
                        "run the loop under autograd (generator.train(), loss.backward()) to use another optimizer")
    if len(optimizer.param_groups) != 1:
        raise ValueError("train_one_epoch: the fused optimizer updates one parameter group (py/main16.py:504); got "
                         f"{len(optimizer.param_groups)} — use the autograd path for per-group settings")
    g = optimizer.param_groups[0]
    bad = {k: g.get(k) for k, ok in (("weight_decay", 0), ("amsgrad", False), ("maximize", False)) if g.get(k, ok) != ok}
    if bad:
        raise ValueError(f"train_one_epoch: optimizer settings {bad} are not implemented by the fused Adam step; "
                         "use the autograd path (loss.backward(); optimizer.step()) for them")


# ---- checkpoints in the reference's format (py/main14d.py:540-560) --------------------------------------------------
def _named_params(generator, detector):
    return [("g", k, p) for k, p in generator.named_parameters()] + [("d", k, p) for k, p in detector.named_parameters()]


def _torch_adam(trainer: TR.Trainer, generator, detector) -> torch.optim.Adam:
    """A genuine torch.optim.Adam over list(generator.parameters()) + list(detector.parameters()) (py/main16.py:504)
    carrying this trainer's moments and step count."""
    opt = torch.optim.Adam([p for _, _, p in _named_params(generator, detector)], lr=trainer.lr, betas=trainer.betas,
                           eps=trainer.eps)
    if trainer.steps > 0:
        gm, gv = TR.unflatten_generator(trainer.g_m), TR.unflatten_generator(trainer.g_v)
        dm, dv = TR.unflatten_detector(trainer.d_m, trainer.nout), TR.unflatten_detector(trainer.d_v, trainer.nout)
        for tag, k, p in _named_params(generator, detector):
            m, v = (gm[k], gv[k]) if tag == "g" else (dm[k], dv[k])
            opt.state[p] = {"step": torch.tensor(float(trainer.steps)), "exp_avg": m.to(p.device),
                            "exp_avg_sq": v.to(p.device)}
    return opt


def save_ckpt(path: str, trainer: TR.Trainer, generator, detector, epoch: int, global_step: int, best_val: float,
              schedule: Optional[OneCycle] = None) -> None:
    """`epoch` is the epoch to start NEXT time, as in the reference."""
    trainer.write_back(generator, detector)
    trainer_steps = trainer.steps
    opt = _torch_adam(trainer, generator, detector)
    sched_state = None
    if schedule is not None:
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, **schedule.kwargs())
        sched.last_epoch = global_step
        sched_state = sched.state_dict()
    tmp = path + ".tmp"
    torch.save({"epoch": epoch, "step": global_step, "best_val": best_val, "gen": generator.state_dict(),
                "det": detector.state_dict(), "opt": opt.state_dict(), "sched": sched_state,
                "wmb200": {"adam_steps": trainer_steps}}, tmp)
    os.replace(tmp, path)


def load_ckpt(path: str, generator, detector, device="cuda", **trainer_kwargs):
    """-> (trainer, next_epoch, global_step, best_val).  Accepts checkpoints written by save_ckpt or by the reference's
    own driver (same keys; `_orig_mod.` prefixes from torch.compile are stripped)."""
    from .models import load_state_dict_strip_prefix
    try:                       # plain tensors / containers: the safe loader is enough for files written by save_ckpt
        ckpt = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:          # checkpoints of older torch versions pickle scheduler internals; the file is the user's own
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
    load_state_dict_strip_prefix(generator, ckpt["gen"])
    load_state_dict_strip_prefix(detector, ckpt["det"])
    generator.to(device)
    detector.to(device)
    trainer = TR.Trainer(generator, detector, **trainer_kwargs)
    named = _named_params(generator, detector)
    state = ckpt["opt"]["state"]
    if state:
        gm, gv, dm, dv = {}, {}, {}, {}
        steps = 0
        for i, (tag, k, p) in enumerate(named):
            st = state[i]
            steps = int(st["step"])
            (gm if tag == "g" else dm)[k] = st["exp_avg"]
            (gv if tag == "g" else dv)[k] = st["exp_avg_sq"]
        trainer.g_m, trainer.g_v = TR.flatten_generator(gm, trainer.device), TR.flatten_generator(gv, trainer.device)
        trainer.d_m = TR.flatten_detector(dm, trainer.nout, trainer.device)
        trainer.d_v = TR.flatten_detector(dv, trainer.nout, trainer.device)
        trainer.steps = steps
        trainer._steps0 = steps
    group = ckpt["opt"]["param_groups"][0]
    trainer.lr, trainer.betas, trainer.eps = float(group["lr"]), tuple(float(b) for b in group["betas"]), float(group["eps"])
    return trainer, int(ckpt["epoch"]), int(ckpt["step"]), float(ckpt["best_val"])


def fit(generator, detector, train_loader, val_loader, epochs: int, lr: float = TR.LR,
        schedule: Optional[OneCycle] = None, patience: int = 3, min_delta: float = 0.001, ckpt_dir: Optional[str] = None,
        resume: bool = True, device="cuda", log: Callable = print, message_fn: Callable = _draw_messages):
    """The training loop of py/main16.py:534-560 (best-model files, early stopping) with the resumable checkpoints
    and optional One-Cycle schedule of py/main14d.py:562-623.  Returns (trainer, train_logs, val_logs)."""
    latest = os.path.join(ckpt_dir, "ckpt_latest.pth") if ckpt_dir else None
    if latest and resume and os.path.exists(latest):
        trainer, start_epoch, global_step, best_val = load_ckpt(latest, generator, detector, device)
        log(f"Resumed from {latest} (next epoch = {start_epoch}, global_step = {global_step})")
    else:
        generator.to(device)
        detector.to(device)
        trainer, start_epoch, global_step, best_val = TR.Trainer(generator, detector, lr=lr), 0, 0, float("inf")
    stopper = EarlyStopping(patience=patience, min_delta=min_delta)
    train_logs: List[dict] = []
    val_logs: List[dict] = []
    for epoch in range(start_epoch, epochs):
        train_metrics = train_one_epoch(trainer, train_loader, schedule=schedule, global_step=global_step,
                                        message_fn=message_fn)
        global_step = trainer.steps
        trainer.write_back(generator, detector)
        generator.eval()
        detector.eval()
        val_metrics = _ev.validate_one_epoch(generator, detector, val_loader, None, device)
        train_logs.append(train_metrics)
        val_logs.append(val_metrics)
        log(f"Epoch {epoch + 1}:")
        for k in train_metrics:
            log(f"  [Train] {k}: {train_metrics[k]:.4f} | [Val] {val_metrics.get(k, float('nan')):.4f}")
        if val_metrics["total"] < best_val:
            best_val = val_metrics["total"]
            if ckpt_dir:
                torch.save(generator.state_dict(), os.path.join(ckpt_dir, "generator_best.pth"))
                torch.save(detector.state_dict(), os.path.join(ckpt_dir, "detector_best.pth"))
                save_ckpt(os.path.join(ckpt_dir, "ckpt_best.pth"), trainer, generator, detector, epoch + 1, global_step,
                          best_val, schedule)
            log("Saved best model")
        if latest:
            save_ckpt(latest, trainer, generator, detector, epoch + 1, global_step, best_val, schedule)
        stopper.step(val_metrics["total"])
        if stopper.early_stop:
            log("Early stopping triggered")
            break
    return trainer, train_logs, val_logs
