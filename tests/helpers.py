"""Shared fixture loading for the parity tests (oracle side only)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def weights():
    return load_npz("main16_weights.npz")


def io():
    return load_npz("main16_io.npz")


def gen_sd(w, tag):
    """Generator state dict (no embedding table) + the stored embedding rows."""
    p = f"gen{tag}/"
    sd = {k[len(p):]: torch.from_numpy(v) for k, v in w.items() if k.startswith(p)}
    rows = sd.pop("emb_rows")
    return sd, rows


def det_sd(w):
    return {k[4:]: torch.from_numpy(v) for k, v in w.items() if k.startswith("det/")}


def emb_for(io_d, rows, messages):
    ids = io_d["emb_row_ids"].tolist()
    return torch.stack([rows[ids.index(int(m))] for m in messages])


def full_embedding(io_d, rows):
    """A 65536x64 table that is zero except for the rows the fixtures use."""
    tab = torch.zeros(65536, 64)
    tab[torch.from_numpy(io_d["emb_row_ids"])] = rows
    return tab
