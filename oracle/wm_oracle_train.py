"""CPU/PyTorch restatement of the Detector's half of the reference's training step — TEST INFRASTRUCTURE ONLY
(tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import it; the product never does).

Follows py/main16.py: ResBlock :112-125 with nn.BatchNorm1d in train mode (batch statistics, running stats updated
with momentum 0.1 and the unbiased variance), Detector :170-186, the detector losses of train_one_epoch :249-264
(loss = LAMBDA_LOC * loc + LAMBDA_DEC * bce, :275-276 restricted to the terms the detector sees), loss.backward()
and torch.optim.Adam(lr=LR) :277-278,504.  Pinned against the reference's own classes by
tests/golden/train_step.npz (tests/golden/make_golden_train.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

PARAM_KEYS = ["model.0.weight", "model.0.bias"] + [
    f"model.{k}.block.{i}.{n}" for k in (1, 2) for i in (0, 1, 3, 4) for n in ("weight", "bias")] + [
    "model.3.weight", "model.3.bias"]


def _resblock_train(x: Tensor, sd: SD, p: str, momentum: float = 0.1) -> Tensor:
    """py/main16.py:112-125 with BatchNorm in training mode; running stats in sd are updated in place."""
    z1 = F.conv1d(x, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], padding=1)
    u = F.relu(F.batch_norm(z1, sd[p + ".block.1.running_mean"], sd[p + ".block.1.running_var"],
                            sd[p + ".block.1.weight"], sd[p + ".block.1.bias"], True, momentum, 1e-5))
    z2 = F.conv1d(u, sd[p + ".block.3.weight"], sd[p + ".block.3.bias"], padding=1)
    y = F.batch_norm(z2, sd[p + ".block.4.running_mean"], sd[p + ".block.4.running_var"],
                     sd[p + ".block.4.weight"], sd[p + ".block.4.bias"], True, momentum, 1e-5)
    return F.relu(x + y)


def detector_forward_train(sd: SD, x: Tensor) -> Tensor:
    """x (B,T) -> logits (B,T,nout); py/main16.py:177-186 in train mode."""
    h = F.conv1d(x.unsqueeze(1), sd["model.0.weight"], sd["model.0.bias"], padding=3)
    h = _resblock_train(h, sd, "model.1")
    h = _resblock_train(h, sd, "model.2")
    return F.conv1d(h, sd["model.3.weight"], sd["model.3.bias"]).permute(0, 2, 1)


def detector_losses(logits: Tensor, message: Optional[Tensor], n_wm: int) -> Tuple[Tensor, Tensor]:
    """(loc, bce) of py/main16.py:252-264; the first n_wm clips are the watermarked half."""
    B2, T, nout = logits.shape
    target = torch.cat([torch.ones(n_wm, T), torch.zeros(B2 - n_wm, T)]).to(logits)
    loc = F.binary_cross_entropy_with_logits(logits[:, :, 0], target)
    if nout == 1 or n_wm == 0:
        return loc, torch.zeros((), dtype=logits.dtype, device=logits.device)
    bits = ((message[:n_wm].unsqueeze(1) >> torch.arange(nout - 1, device=message.device)) & 1).to(logits.dtype)
    bce = F.binary_cross_entropy_with_logits(logits[:n_wm, :, 1:], bits.unsqueeze(1).expand(-1, T, -1))
    return loc, bce


class DetectorTrainOracle:
    """State dict in, `steps` of forward / backward / Adam out."""

    def __init__(self, sd: SD, lr: float = 1e-3, lam_loc: float = 10.0, lam_dec: float = 1.0, dtype=torch.float32,
                 device="cpu"):
        self.sd = {k: v.detach().clone().to(device=device, dtype=dtype if v.is_floating_point() else v.dtype)
                   for k, v in sd.items()}
        self.params: List[Tensor] = []
        for k in PARAM_KEYS:
            self.sd[k].requires_grad_(True)
            self.params.append(self.sd[k])
        self.opt = torch.optim.Adam(self.params, lr=lr)
        self.lam_loc, self.lam_dec = lam_loc, lam_dec

    def step(self, x: Tensor, message: Optional[Tensor], n_wm: int, update: bool = True):
        x = x.detach().clone().to(self.params[0]).requires_grad_(True)
        self.opt.zero_grad()
        logits = detector_forward_train(self.sd, x)
        loc, bce = detector_losses(logits, message, n_wm)
        (self.lam_loc * loc + self.lam_dec * bce).backward()
        grads = {k: (self.sd[k].grad.detach().clone() if self.sd[k].grad is not None else torch.zeros_like(self.sd[k]))
                 for k in PARAM_KEYS}
        if update:
            self.opt.step()
        return {"loc": loc.detach(), "bce": bce.detach(), "grads": grads, "d_input": x.grad.detach().clone(),
                "logits": logits.detach()}

    def state_dict(self) -> SD:
        return {k: v.detach().clone() for k, v in self.sd.items()}


# ---- the whole step: generator + post-processing + detector + every loss (py/main16.py:244-278) ------------------
from oracle import wm_oracle as _O  # noqa: E402  (post-processing and loss definitions, py/main16.py:53-81,192-217)

G_PARAM_KEYS = (["encoder.0.weight", "encoder.0.bias"]
                + [f"encoder.{k}.block.{i}.{n}" for k in (1, 2) for i in (0, 1, 3, 4) for n in ("weight", "bias")]
                + ["lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0", "embedding.weight",
                   "decoder.0.weight", "decoder.0.bias"]
                + [f"decoder.1.block.{i}.{n}" for i in (0, 1, 3, 4) for n in ("weight", "bias")]
                + ["decoder.2.weight", "decoder.2.bias"])
LAMBDAS = dict(l1=1.0, msspec=4.0, loud=20.0, loc=10.0, dec=1.0, hf=5.0)       # py/main16.py:38-43


def generator_forward_train(sd: SD, s: Tensor, message: Tensor) -> Tensor:
    """s (B,1,T), message (B,) -> raw delta (B,1,T); py/main16.py:148-162 with BatchNorm in train mode."""
    x = F.conv1d(s, sd["encoder.0.weight"], sd["encoder.0.bias"], padding=3)
    x = _resblock_train(x, sd, "encoder.1")
    x = _resblock_train(x, sd, "encoder.2")
    lstm = torch.nn.LSTM(64, 64, batch_first=True).to(device=s.device, dtype=s.dtype)
    x, _ = torch.func.functional_call(lstm, {k[len("lstm."):]: v for k, v in sd.items() if k.startswith("lstm.")},
                                      (x.permute(0, 2, 1),))
    x = x.permute(0, 2, 1) + F.embedding(message, sd["embedding.weight"]).unsqueeze(-1)
    x = F.conv_transpose1d(x, sd["decoder.0.weight"], sd["decoder.0.bias"], padding=3)
    x = _resblock_train(x, sd, "decoder.1")
    return F.conv1d(x, sd["decoder.2.weight"], sd["decoder.2.bias"])


class TrainOracle:
    """Generator and Detector state dicts in; `step(s, message)` = one iteration of train_one_epoch."""

    def __init__(self, gsd: SD, dsd: SD, lr: float = 1e-3, lambdas=None, dtype=torch.float32, device="cpu"):
        conv = lambda sd: {k: v.detach().clone().to(device=device, dtype=dtype if v.is_floating_point() else v.dtype)
                           for k, v in sd.items()}
        self.gsd, self.dsd = conv(gsd), conv(dsd)
        self.params: List[Tensor] = []
        for sd, keys in ((self.gsd, G_PARAM_KEYS), (self.dsd, PARAM_KEYS)):
            for k in keys:
                sd[k].requires_grad_(True)
                self.params.append(sd[k])
        self.opt = torch.optim.Adam(self.params, lr=lr)                       # py/main16.py:504
        self.lam = dict(LAMBDAS, **(lambdas or {}))

    def step(self, s: Tensor, message: Tensor, update: bool = True):
        s = s.detach().to(self.params[0])
        if s.dim() == 2:
            s = s.unsqueeze(1)
        B = s.shape[0]
        self.opt.zero_grad()
        delta = _O.limit_rms(_O.clamp_peak(_O.fir_lowpass(generator_forward_train(self.gsd, s, message))))
        s_w = s + delta
        logits = detector_forward_train(self.dsd, torch.cat([s_w, s], dim=0)[:, 0])
        loc, bce = detector_losses(logits, message, B)
        l1 = delta.abs().mean()
        mel, loud, hf = _O.mel_loss(s, s_w), _O.loudness_loss(s, s_w), _O.high_freq_penalty(delta)
        lam = self.lam
        total = (lam["l1"] * l1 + lam["msspec"] * mel + lam["loud"] * loud + lam["loc"] * loc + lam["dec"] * bce
                 + lam["hf"] * hf)
        total.backward()
        zero = lambda t: t.grad.detach().clone() if t.grad is not None else torch.zeros_like(t)
        out = {"l1": l1, "mel": mel, "loud": loud, "loc": loc, "bce": bce, "hf": hf, "total": total,
               "raw_total": l1 + mel + loud + loc + bce}
        out = {k: v.detach() for k, v in out.items()}
        out["g_grads"] = {k: zero(self.gsd[k]) for k in G_PARAM_KEYS}
        out["d_grads"] = {k: zero(self.dsd[k]) for k in PARAM_KEYS}
        out["s_w"] = s_w.detach()[:, 0]
        if update:
            self.opt.step()
        return out

    def state_dicts(self):
        return ({k: v.detach().clone() for k, v in self.gsd.items()}, {k: v.detach().clone() for k, v in self.dsd.items()})
