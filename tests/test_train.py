"""Training step of the Detector (py/main16.py:249-264,275-278,504 — BASELINE config 4, detector half).

CPU: the oracle restatement against vectors produced by the reference's own classes (tests/golden/train_step.npz).
GPU: the CUDA operators against PyTorch autograd, and the fused `wm_detector_train_step` against the golden vectors
and against the oracle on seeded inputs.

Convolution biases that feed a BatchNorm have an analytically ZERO gradient (the batch mean removes them); what
autograd — and this library — computes there is round-off noise, which Adam's g / (|g| + eps) normalisation turns
into updates of arbitrary sign.  The parameter comparisons therefore skip `block.0.bias` / `block.3.bias`
(the gradients themselves are checked to be ~0 on both sides).

fp32 against fp64: a handful of pre-activations sit within round-off of zero, their ReLU masks flip between two
arithmetics and the early layers' gradients move by 1e-3..1e-2 relative (the input gradient by a few 1e-2 at single
samples) — PyTorch's own fp32 autograd shows the same deviations from its fp64 run (tools/train_precision.py,
profiles/r1_train_precision.txt).  The end-to-end gradient checks are therefore RELATIVE TO THAT: this library's
distance from the fp64 oracle must stay within 3x the distance of the fp32 oracle from the fp64 oracle (+1e-4); the
single operators, which have no such discontinuity at random inputs, are held to 2e-5."""
import numpy as np
import pytest

L_MATH_FP32 = 0      # WM_MATH_FP32 (include/wmb200.h)
import torch
import torch.nn.functional as F

from oracle import wm_oracle_train as OT
from tests import helpers as H

G = H.load_npz("train_step.npz")
LR, LAM_LOC, LAM_DEC = (float(v) for v in G["hyper"])
DEAD = ("block.0.bias", "block.3.bias")


def sd_of(prefix):
    return {k[len(prefix):]: torch.from_numpy(np.asarray(v)) for k, v in G.items() if k.startswith(prefix)}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))



@pytest.fixture(autouse=True)
def _fp32_math_mode(request):
    """The operator-level checks of this module compare with autograd at fp32 round-off level, so they run the library
    in WM_MATH_FP32 (exact-order fp32 FMA kernels).  The default mode -- training convolutions on tcgen05 with bf16
    hi+lo operand pairs -- has its own tests at the end of the module (`tensor_core` in the name) with gates stated
    relative to PyTorch's own fp32 (TF32) step, and is what tests/test_train_full.py and test_autograd_loop.py run."""
    if "tensor_core" in request.node.name or not torch.cuda.is_available():
        yield
        return
    from wmb200 import ops
    prev = ops.set_math_mode(L_MATH_FP32)
    try:
        yield
    finally:
        ops.set_math_mode(prev)


def test_oracle_reproduces_the_reference_training_steps():
    o = OT.DetectorTrainOracle(sd_of("init."), lr=LR, lam_loc=LAM_LOC, lam_dec=LAM_DEC)
    for step in range(2):
        x, msg = torch.from_numpy(G[f"x{step}"]), torch.from_numpy(G[f"message{step}"])
        r = o.step(x, msg, n_wm=x.shape[0] // 2)
        assert abs(float(r["loc"]) - G[f"losses{step}"][0]) < 1e-6
        assert abs(float(r["bce"]) - G[f"losses{step}"][1]) < 1e-6
        if step == 0:
            gmax = max(float(np.abs(G["grad1." + k]).max()) for k in OT.PARAM_KEYS)
            for k in OT.PARAM_KEYS:
                want = torch.from_numpy(G["grad1." + k])
                if k.endswith(DEAD):
                    assert float(r["grads"][k].abs().max()) < 2e-4 * gmax and float(want.abs().max()) < 2e-4 * gmax
                else:
                    assert rel(r["grads"][k], want) < 1e-4, k
            assert rel(r["d_input"], torch.from_numpy(G["grad1.input"])) < 1e-4
    final, want = o.state_dict(), sd_of("final.")
    for k, v in want.items():
        if k.endswith(DEAD) or k.endswith("num_batches_tracked"):
            continue
        assert float((final[k] - v).abs().max()) < 2e-5, k


def test_flat_layout_round_trips():
    from wmb200 import train as TR
    sd = sd_of("init.")
    flat = TR.flatten_detector(sd, 17, "cpu")
    back = TR.unflatten_detector(flat, 17)
    for k in OT.PARAM_KEYS:
        assert torch.equal(back[k], sd[k]), k
    assert flat.numel() == TR.L.DT_SIZE


def test_trainer_refuses_cpu():
    import wmb200
    from wmb200 import train as TR
    det = wmb200.Detector(message_bits=16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TR.DetectorTrainer(det)


# ---------------------------------------------------------------- GPU ------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False)])
def test_bn_train_forward_backward_match_autograd(relu, res):
    from wmb200 import train as TR
    torch.manual_seed(0)
    B, T = 3, 1111
    z = (torch.randn(B, T, 64, device="cuda") * 1.5 + 0.3).requires_grad_(True)
    r = torch.randn(B, T, 64, device="cuda").requires_grad_(True) if res else None
    gamma = (0.5 + torch.rand(64, device="cuda")).requires_grad_(True)
    beta = (0.1 * torch.randn(64, device="cuda")).requires_grad_(True)
    rm, rv = 0.1 * torch.randn(64, device="cuda"), 0.5 + torch.rand(64, device="cuda")
    rm_t, rv_t = rm.clone(), rv.clone()
    y = F.batch_norm(z.permute(0, 2, 1).double(), rm_t.double(), rv_t.double(), gamma.double(), beta.double(), True, 0.1,
                     1e-5).permute(0, 2, 1)
    rm_t, rv_t = rm.clone().double(), rv.clone().double()
    F.batch_norm(z.detach().permute(0, 2, 1).double(), rm_t, rv_t, None, None, True, 0.1, 1e-5)
    if res:
        y = y + r.double()
    if relu:
        y = F.relu(y)
    dout = torch.randn(B, T, 64, device="cuda")
    y.backward(dout.double())
    out, mean, rstd = TR.bn_train_fwd(z.detach(), gamma.detach(), beta.detach(), r.detach() if res else None, relu, rm, rv)
    assert float((out.double() - y.detach()).abs().max()) < 5e-6
    assert float((rm.double() - rm_t).abs().max()) < 1e-6 and float((rv.double() - rv_t).abs().max()) < 1e-6
    dz, dres, dg, db = TR.bn_train_bwd(dout, out if relu else None, z.detach(), mean, rstd, gamma.detach(), res)
    assert rel(dz, z.grad) < 2e-5
    assert rel(dg, gamma.grad) < 2e-5 and rel(db, beta.grad) < 2e-5
    if res:
        assert rel(dres, r.grad) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("K,B,T", [(3, 2, 700), (7, 1, 1300), (1, 3, 64), (3, 5, 513)])
def test_conv64_backward_matches_autograd(K, B, T):
    from wmb200 import train as TR
    torch.manual_seed(K)
    x = torch.randn(B, T, 64, device="cuda")
    w = (torch.randn(64, 64, K, device="cuda") / (64 * K) ** 0.5)
    dy = torch.randn(B, T, 64, device="cuda")
    xd = x.double().permute(0, 2, 1).requires_grad_(True)
    wd = w.double().requires_grad_(True)
    bd = torch.zeros(64, device="cuda", dtype=torch.float64, requires_grad=True)
    F.conv1d(xd, wd, bd, padding=K // 2).backward(dy.double().permute(0, 2, 1))
    dw, db, dx = TR.conv64_bwd(x, dy, w)
    assert rel(dw, wd.grad) < 2e-5
    assert rel(db, bd.grad) < 2e-5
    assert rel(dx, xd.grad.permute(0, 2, 1)) < 2e-5


@pytest.mark.gpu
def test_adam_matches_torch_optim():
    from wmb200 import train as TR
    torch.manual_seed(3)
    n = 100_003
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(n, device="cuda") * 10.0 ** float(torch.randint(-4, 2, ()))
        ref.grad = g.clone()
        opt.step()
        TR.adam_step(p, g, m, v, step, lr=1e-3)
        assert float((p - ref.detach()).abs().max()) < 2e-6


def _check_against(tr, sd, steps):
    """Three trajectories on the same batches: this library, the oracle in fp32 and the oracle in fp64."""
    o64 = OT.DetectorTrainOracle(sd, dtype=torch.float64, device="cuda")
    o32 = OT.DetectorTrainOracle(sd, dtype=torch.float32, device="cuda")
    for step, (x, msg, n_wm) in enumerate(steps):
        want, base = o64.step(x, msg, n_wm), o32.step(x, msg, n_wm)
        got = tr.step(x, msg, n_wm, want_input_grad=True)
        assert abs(float(got["loc"]) - float(want["loc"])) < 2e-5
        assert abs(float(got["bce"]) - float(want["bce"])) < 2e-5
        gd = tr.grad_dict()
        gmax = max(float(v.abs().max()) for v in want["grads"].values())
        for k in OT.PARAM_KEYS:
            if k.endswith(DEAD):
                assert float(gd[k].abs().max()) < 2e-4 * gmax, k
            else:
                assert rel(gd[k], want["grads"][k]) < 3 * rel(base["grads"][k], want["grads"][k]) + 1e-4, (step, k)
        assert rel(got["d_input"], want["d_input"]) < 3 * rel(base["d_input"], want["d_input"]) + 1e-4
        # the bulk of the input gradient is resolved far better than its worst sample
        if step == 0:   # (later steps run on parameters that Adam has already moved apart, see below)
            d = (got["d_input"].double() - want["d_input"]).abs()
            assert float(d.median()) < 1e-3 * float(want["d_input"].abs().max())
    final, wantsd = tr.state_dict(), o64.state_dict()
    for k, v in wantsd.items():
        if k.endswith(DEAD) or k.endswith("num_batches_tracked"):
            continue
        d = (final[k].double().cpu() - v.double().cpu()).abs()
        if "running" in k:   # the running mean carries the (arbitrarily moved) dead bias: 0.1 * lr per step
            tol = 0.1 * 2.1 * LR * len(steps) if k.endswith("mean") else 1e-3 * max(1.0, float(v.abs().max()))
            assert float(d.max()) < tol, k
            continue
        # Adam moves every weight by ~lr per step whatever the gradient's size: a gradient whose sign is not
        # resolved costs O(lr); everything else lands on the fp64 trajectory
        assert float(d.max()) <= 2.1 * LR * len(steps), k
        assert float(d.median()) < 0.1 * LR, k


@pytest.mark.gpu
def test_detector_train_step_matches_the_reference_golden():
    import wmb200
    from wmb200 import train as TR
    det = wmb200.Detector(message_bits=16)
    det.load_state_dict(sd_of("init."))
    tr = TR.DetectorTrainer(det.cuda(), lr=LR, lambda_loc=LAM_LOC, lambda_dec=LAM_DEC)
    for step in range(2):
        x, msg = torch.from_numpy(G[f"x{step}"]).cuda(), torch.from_numpy(G[f"message{step}"]).cuda()
        r = tr.step(x, msg, want_input_grad=(step == 0))
        assert abs(float(r["loc"]) - G[f"losses{step}"][0]) < 2e-5
        assert abs(float(r["bce"]) - G[f"losses{step}"][1]) < 2e-5
        if step == 0:
            gd = tr.grad_dict()
            gmax = max(float(np.abs(G["grad1." + k]).max()) for k in OT.PARAM_KEYS)
            for k in OT.PARAM_KEYS:
                want = torch.from_numpy(G["grad1." + k])
                if k.endswith(DEAD):
                    assert float(gd[k].abs().max()) < 2e-4 * gmax
                else:   # fp32 here against fp32 on the reference's CPU: ReLU-mask flips, see the module docstring
                    assert rel(gd[k], want) < (2e-5 if k.startswith("model.3") else 3e-2), k
            d = (r["d_input"].cpu().double() - torch.from_numpy(G["grad1.input"]).double()).abs()
            assert float(d.median()) < 1e-4 * float(np.abs(G["grad1.input"]).max())
    final, want = tr.state_dict(), sd_of("final.")
    for k in ("model.1.block.1.running_mean", "model.1.block.1.running_var", "model.2.block.4.running_mean",
              "model.2.block.4.running_var", "model.3.weight", "model.0.weight", "model.1.block.1.weight"):
        assert float((final[k].cpu() - want[k]).abs().max()) < (5e-4 if "running" in k else 2.1 * LR * 2), k
    # parameters moved the same way as the reference's wherever the gradient is resolved
    for k in ("model.3.weight", "model.1.block.1.weight", "model.2.block.3.weight"):
        d = (final[k].cpu() - want[k]).abs()
        assert float(d.median()) < 1e-5, k


@pytest.mark.gpu
@pytest.mark.parametrize("bits,B2,n_wm,T", [(16, 6, 3, 1000), (0, 4, 2, 777), (16, 5, 2, 2049), (4, 2, 2, 16000)])
def test_detector_trainer_matches_oracle(bits, B2, n_wm, T):
    import wmb200
    from wmb200 import train as TR
    torch.manual_seed(bits + B2)
    det = wmb200.Detector(message_bits=bits)
    with torch.no_grad():
        for m in det.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(0.8 + 0.4 * torch.rand(64))
                m.bias.copy_(0.1 * torch.randn(64))
    sd = det.state_dict()
    tr = TR.DetectorTrainer(det.cuda())
    steps = []
    for _ in range(3):
        x = 0.1 * torch.randn(B2, T, device="cuda")
        msg = torch.randint(0, 2 ** max(bits, 1), (n_wm,), device="cuda") if bits else None
        steps.append((x, msg, n_wm))
    _check_against(tr, sd, steps)


@pytest.mark.gpu
def test_write_back_and_eval_inference():
    """After training steps the module runs the inference path with the updated parameters and running stats."""
    import wmb200
    from oracle import wm_oracle as O
    from wmb200 import train as TR
    torch.manual_seed(5)
    det = wmb200.Detector(message_bits=16).cuda()
    tr = TR.DetectorTrainer(det)
    x = 0.1 * torch.randn(4, 4000, device="cuda")
    msg = torch.randint(0, 65536, (2,), device="cuda")
    before = {k: v.clone() for k, v in det.state_dict().items()}
    l0 = float(tr.step(x, msg)["loc"])
    for _ in range(4):
        last = tr.step(x, msg)
    assert float(last["loc"]) < l0                     # the same batch five times: the loss goes down
    tr.write_back(det)
    after = det.state_dict()
    assert int(after["model.1.block.1.num_batches_tracked"]) == int(before["model.1.block.1.num_batches_tracked"]) + 5
    assert float((after["model.3.weight"] - before["model.3.weight"]).abs().max()) > 1e-4
    det.eval()
    got = det(x.unsqueeze(1))
    want = O.detector_forward({k: v.cpu() for k, v in after.items()}, x.cpu().unsqueeze(1))
    assert float((got.cpu() - want).abs().max()) < 2e-3


@pytest.mark.gpu
def test_update_false_leaves_parameters_alone_and_is_deterministic():
    import wmb200
    from wmb200 import train as TR
    torch.manual_seed(9)
    det = wmb200.Detector(message_bits=16).cuda()
    tr = TR.DetectorTrainer(det)
    p0 = tr.params.clone()
    x = 0.1 * torch.randn(4, 3000, device="cuda")
    msg = torch.randint(0, 65536, (2,), device="cuda")
    tr.step(x, msg, update=False)
    g1 = tr.grads.clone()
    assert torch.equal(tr.params, p0) and tr.steps == 0
    tr.step(x, msg, update=False)
    assert torch.equal(tr.grads, g1)                  # fixed-order reductions: bit-identical gradients


@pytest.mark.gpu
@pytest.mark.parametrize("B,T", [(1, 1), (3, 300), (2, 4000)])
def test_lstm_train_forward_backward_match_autograd(B, T):
    from wmb200 import train as TR
    torch.manual_seed(T)
    lstm = torch.nn.LSTM(64, 64, batch_first=True).cuda().double()
    x = torch.randn(B, T, 64, device="cuda", dtype=torch.float64, requires_grad=True)
    dy = torch.randn(B, T, 64, device="cuda")
    y, _ = lstm(x)
    y.backward(dy.double())
    p = {k: v.detach().float() for k, v in lstm.named_parameters()}
    h, saved = TR.lstm_train_fwd(x.detach().float(), p["weight_ih_l0"], p["weight_hh_l0"], p["bias_ih_l0"], p["bias_hh_l0"])
    assert float((h.double() - y.detach()).abs().max()) < 2e-5
    dx, dwi, dwh, db = TR.lstm_train_bwd(dy, saved)
    assert rel(dx, x.grad) < 1e-4
    assert rel(dwi, lstm.weight_ih_l0.grad) < 1e-4
    assert rel(dwh, lstm.weight_hh_l0.grad) < 1e-4
    assert rel(db, lstm.bias_ih_l0.grad) < 1e-4 and rel(db, lstm.bias_hh_l0.grad) < 1e-4


# ---- backward of the losses and of the post-processing against autograd of the oracle's definitions -------------
def _signals(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(T) / 16000.0
    s = 0.1 * torch.randn(B, T, generator=g) + 0.2 * torch.sin(2 * np.pi * 300.0 * t)[None]
    delta = 0.004 * torch.randn(B, T, generator=g)
    delta[0] *= 4.0             # one clip above the RMS cap with samples beyond the peak clamp
    return s.cuda(), delta.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("B,T", [(2, 16000), (3, 4097)])
def test_loss_gradients_match_autograd(B, T):
    from oracle import wm_oracle as O
    from wmb200 import packing
    from wmb200 import train as TR
    s, delta = _signals(B, T, T)
    sd, dd = s.double(), delta.double().requires_grad_(True)
    swd = sd + dd
    # hf penalty w.r.t. delta
    (g_ref,) = torch.autograd.grad(O.high_freq_penalty(dd.unsqueeze(1)), dd)
    got = TR.hf_penalty_bwd(delta, 512, 113, weight=5.0)
    assert rel(got, 5.0 * g_ref) < 2e-4
    # loudness and mel w.r.t. the watermarked signal
    (g_ref,) = torch.autograd.grad(O.loudness_loss(sd, swd), dd)
    got = TR.loudness_bwd(s, s + delta, weight=20.0)
    assert rel(got, 20.0 * g_ref) < 2e-4
    (g_ref,) = torch.autograd.grad(O.mel_loss(sd.unsqueeze(1), swd.unsqueeze(1)), dd)
    fb, band = packing.mel_filterbank(513, 64, 16000)
    got = TR.mel_log_l1_bwd(s, s + delta, fb.cuda(), band.cuda(), weight=4.0)
    assert rel(got, 4.0 * g_ref) < 1e-3
    # |delta| mean, accumulated on top of the previous gradient
    assert rel(TR.abs_mean_bwd(delta, weight=2.0), 2.0 * torch.sign(delta) / delta.numel()) < 1e-6
    acc = got.clone()
    TR.abs_mean_bwd(delta, weight=1.0, into=acc)
    assert rel(acc, got + torch.sign(delta) / delta.numel()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [7, 6, 4, 1, 0])
def test_postprocess_backward_matches_autograd(mode):
    from oracle import wm_oracle as O
    from wmb200 import train as TR
    _, delta = _signals(3, 16000, 3)
    raw = (delta * 1.5).double().requires_grad_(True)
    d = raw.unsqueeze(1)
    if mode & 1:
        d = O.fir_lowpass(d)
    if mode & 2:
        d = O.clamp_peak(d)
    if mode & 4:
        d = O.limit_rms(d)
    g = torch.randn(3, 16000, device="cuda")
    (want,) = torch.autograd.grad((d[:, 0] * g.double()).sum(), raw)
    fir = O.fir_taps().float().cuda()
    got = TR.postprocess_bwd(g, raw.detach().float(), fir, mode)
    assert rel(got, want) < 2e-4


@pytest.mark.gpu
def test_head_bce_and_input_conv_backward_match_autograd():
    import torch.nn.functional as Fn
    from wmb200 import train as TR
    torch.manual_seed(12)
    B2, n_wm, T, nout = 4, 2, 1501, 17
    y = torch.randn(B2, T, 64, device="cuda", dtype=torch.float64, requires_grad=True)
    w = (torch.randn(nout, 64, 1, device="cuda", dtype=torch.float64) / 8).requires_grad_(True)
    b = torch.zeros(nout, device="cuda", dtype=torch.float64, requires_grad=True)
    msg = torch.randint(0, 65536, (n_wm,), device="cuda")
    logits = Fn.conv1d(y.permute(0, 2, 1), w, b).permute(0, 2, 1)
    loc, bce = OT.detector_losses(logits, msg, n_wm)
    logits.retain_grad()
    (10.0 * loc + 1.0 * bce).backward()
    dl = TR.bce_heads_bwd(logits.detach().float(), msg, n_wm, 10.0, 1.0)
    assert rel(dl, logits.grad) < 2e-5
    dy, dw, db = TR.head_bwd(dl, y.detach().float(), w.detach().float())
    assert rel(dy, y.grad) < 2e-5 and rel(dw, w.grad) < 2e-5
    assert rel(db, b.grad) < 1e-4        # 6004 signed terms per output, fp32 within a 1024-row block
    s = torch.randn(3, 2000, device="cuda", dtype=torch.float64, requires_grad=True)
    wi = torch.randn(64, 1, 7, device="cuda", dtype=torch.float64, requires_grad=True)
    bi = torch.zeros(64, device="cuda", dtype=torch.float64, requires_grad=True)
    dx = torch.randn(3, 2000, 64, device="cuda")
    Fn.conv1d(s.unsqueeze(1), wi, bi, padding=3).backward(dx.double().permute(0, 2, 1))
    dwi, dbi, ds = TR.conv_in_k7_bwd(s.detach().float(), dx, wi.detach().float())
    assert rel(dwi, wi.grad) < 2e-5 and rel(dbi, bi.grad) < 2e-5 and rel(ds, s.grad) < 2e-5


# ---- the default math mode: training convolutions on the tensor cores ---------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("K,B,T", [(3, 2, 700), (7, 1, 1300), (3, 3, 16000), (7, 2, 16000), (3, 1, 129)])
def test_tensor_core_conv64_forward_backward(K, B, T):
    """wm_conv64_train_fwd / wm_conv64_bwd in WM_MATH_BF16X2: forward and data gradient through the tcgen05
    convolution (3 partial products), weight gradient through the MN-major tcgen05 kernel (4 partial products), against
    fp64 autograd.  bf16 hi+lo operands carry 16 mantissa bits: 2^-17 relative per product."""
    from wmb200 import ops
    from wmb200 import train as TR
    assert ops.get_math_mode() == 1
    g = torch.Generator().manual_seed(K * 100 + T)
    x = torch.randn(B, T, 64, generator=g).cuda()
    w = (torch.randn(64, 64, K, generator=g) / (8 * K ** 0.5)).cuda()
    b = torch.randn(64, generator=g).cuda()
    dy = torch.randn(B, T, 64, generator=g).cuda()
    xd = x.double().permute(0, 2, 1).requires_grad_()
    wd, bd = w.double().requires_grad_(), b.double().requires_grad_()
    yd = torch.nn.functional.conv1d(xd, wd, bd, padding=K // 2)
    yd.backward(dy.double().permute(0, 2, 1))
    y = TR.conv64_train_fwd(x, w, b)
    rel = lambda a, r: float((a.double() - r).abs().max() / r.abs().max())
    assert rel(y.permute(0, 2, 1), yd.detach()) < 3e-5
    dw, db, dx = TR.conv64_bwd(x, dy, w)
    assert rel(dx.permute(0, 2, 1), xd.grad) < 3e-5
    assert rel(dw, wd.grad) < 3e-5
    assert rel(db, bd.grad) < 1e-5
    dw2, db2, _ = TR.conv64_bwd(x, dy, w)                     # deterministic: fixed-order partial sums
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


@pytest.mark.gpu
@pytest.mark.parametrize("bits,B2,n_wm,T", [(16, 6, 3, 1000), (16, 5, 2, 2049), (4, 2, 2, 16000)])
def test_tensor_core_detector_trainer_stays_within_pytorch_fp32(bits, B2, n_wm, T):
    """Three detector training steps in the default math mode next to the fp64 oracle and to PyTorch's own fp32 step on
    the same GPU (cuDNN, TF32 allowed -- the reference's setting, py/main16.py:44): losses and every resolved gradient
    stay within 3x PyTorch's distance to fp64."""
    import wmb200
    from wmb200 import train as TR
    torch.manual_seed(bits + B2)
    det = wmb200.Detector(message_bits=bits)
    sd = det.state_dict()
    tr = TR.DetectorTrainer(det.cuda())
    o64 = OT.DetectorTrainOracle(sd, dtype=torch.float64, device="cuda")
    o32 = OT.DetectorTrainOracle(sd, dtype=torch.float32, device="cuda")
    for step in range(3):
        x = 0.1 * torch.randn(B2, T, device="cuda")
        msg = torch.randint(0, 2 ** bits, (n_wm,), device="cuda")
        want, base = o64.step(x, msg, n_wm), o32.step(x, msg, n_wm)
        got = tr.step(x, msg, n_wm, want_input_grad=True)
        for k in ("loc", "bce"):
            assert abs(float(got[k]) - float(want[k])) < 3 * abs(float(base[k]) - float(want[k])) + 1e-4, (step, k)
        gd = tr.grad_dict()
        for k in OT.PARAM_KEYS:
            if not k.endswith(DEAD):
                assert rel(gd[k], want["grads"][k]) < 3 * rel(base["grads"][k], want["grads"][k]) + 1e-3, (step, k)
        assert rel(got["d_input"], want["d_input"]) < 3 * rel(base["d_input"], want["d_input"]) + 1e-3
