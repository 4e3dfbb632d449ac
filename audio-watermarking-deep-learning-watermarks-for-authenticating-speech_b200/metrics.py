"""Detection statistics of the reference's evaluation cells on device scores: confusion counts and
the classification report at a threshold (py/main16.py:1335-1341), ROC curve and AUC (py/main16.py:2372-2386).  The
counting runs on the GPU (wm_eval.cu); only a handful of integers reach the host."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .ops import _req, _stream


def confusion_counts(clean_probs: torch.Tensor, wm_probs: torch.Tensor, thresh: float = 0.5) -> dict:
    """{tn, fp, fn, tp} with prediction = p >= thresh, labels clean = 0 / watermarked = 1 (py/main16.py:1335-1339);
    `matrix` is sklearn's confusion_matrix layout [[tn, fp], [fn, tp]]."""
    lib = L.load()
    c, w = _req(clean_probs.reshape(-1), "clean_probs"), _req(wm_probs.reshape(-1), "wm_probs")
    out = torch.empty(4, dtype=torch.int64, device=c.device)
    L.check(lib.wm_confusion_counts_fwd(L.ptr(c), c.numel(), L.ptr(w), w.numel(), float(thresh), L.ptr(out), _stream()),
            "wm_confusion_counts_fwd")
    tn, fp, fn, tp = (int(v) for v in out.cpu())
    return {"tn": tn, "fp": fp, "fn": fn, "tp": tp, "matrix": np.array([[tn, fp], [fn, tp]])}


def classification_report(clean_probs: torch.Tensor, wm_probs: torch.Tensor, thresh: float = 0.5) -> dict:
    """Precision / recall / F1 / support per class and accuracy, as sklearn.metrics.classification_report(...,
    output_dict=True) reports them for target_names ["Clean", "Watermarked"] (py/main16.py:1341)."""
    k = confusion_counts(clean_probs, wm_probs, thresh)
    tn, fp, fn, tp = k["tn"], k["fp"], k["fn"], k["tp"]
    div = lambda a, b: a / b if b else 0.0
    rep = {}
    for name, t, f_pred, f_miss, sup in (("Clean", tn, fn, fp, tn + fp), ("Watermarked", tp, fp, fn, fn + tp)):
        prec, rec = div(t, t + f_pred), div(t, t + f_miss)
        rep[name] = {"precision": prec, "recall": rec, "f1-score": div(2 * prec * rec, prec + rec), "support": sup}
    rep["accuracy"] = div(tn + tp, tn + fp + fn + tp)
    rep["confusion_matrix"] = k["matrix"]
    return rep


def roc_curve(clean_probs: torch.Tensor, wm_probs: torch.Tensor):
    """(fpr, tpr, thresholds) as sklearn.metrics.roc_curve(y_true, y_score, drop_intermediate=False) returns them:
    one point per distinct score in decreasing order, preceded by (0, 0) at threshold +inf.  The thresholds are the
    scores themselves (sorted on the device by torch), the counting per threshold is one kernel."""
    lib = L.load()
    c, w = _req(clean_probs.reshape(-1), "clean_probs"), _req(wm_probs.reshape(-1), "wm_probs")
    thr = torch.unique(torch.cat([c, w])).flip(0).contiguous()          # distinct scores, descending
    nt = thr.numel()
    fp = torch.empty(nt, dtype=torch.int32, device=c.device)
    tp = torch.empty(nt, dtype=torch.int32, device=c.device)
    L.check(lib.wm_roc_points_fwd(L.ptr(c), c.numel(), L.ptr(w), w.numel(), L.ptr(thr), nt, L.ptr(fp), L.ptr(tp),
                                  _stream()), "wm_roc_points_fwd")
    fpr = np.concatenate([[0.0], fp.cpu().numpy() / max(c.numel(), 1)])
    tpr = np.concatenate([[0.0], tp.cpu().numpy() / max(w.numel(), 1)])
    return fpr, tpr, np.concatenate([[np.inf], thr.cpu().numpy()])


def auc(clean_probs: torch.Tensor, wm_probs: torch.Tensor) -> float:
    """Area under the ROC curve, exactly (rank statistic over all clean x watermarked pairs, ties count half) — what
    sklearn.metrics.auc(fpr, tpr) gives on the full curve (py/main16.py:2377)."""
    lib = L.load()
    c, w = _req(clean_probs.reshape(-1), "clean_probs"), _req(wm_probs.reshape(-1), "wm_probs")
    out = torch.empty(1, dtype=torch.int64, device=c.device)
    L.check(lib.wm_auc_pairs_fwd(L.ptr(c), c.numel(), L.ptr(w), w.numel(), L.ptr(out), _stream()), "wm_auc_pairs_fwd")
    n = c.numel() * w.numel()
    return float(out.item()) / (2.0 * n) if n else float("nan")
