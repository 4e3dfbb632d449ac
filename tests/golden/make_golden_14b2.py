"""Golden fixtures for main14b_2 FROM THE REFERENCE ITSELF (build container only: needs /root/reference).

    python tests/golden/make_golden_14b2.py

AST-extracts make_conv1d / ResidualBlock / Generator / Detector and the hyper-parameter constants from
/root/reference/py/main14b_2.py (the file trains at import time), builds both models under fixed seeds on CPU
and stores inputs, messages, outputs and per-tensor checksums of the seeded parameters in main14b2_io.npz
(the parameters themselves are ~20 M floats and are NOT stored: the drop-in modules reproduce them draw by draw).
"""
import ast
import os
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference/py/main14b_2.py"
HERE = os.path.dirname(os.path.abspath(__file__))
WANT = {"make_conv1d", "ResidualBlock", "Generator", "Detector"}
CONSTS = {"HIDDEN_DIM", "NUM_BITS", "CHANNELS", "OUTPUT_CH", "STRIDES", "LSTM_LAYERS"}
SEED_G, SEED_D = 1402, 1403


def extract():
    tree = ast.parse(open(REF).read())
    seen, body = set(), []
    for node in tree.body:
        name = None
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in WANT:
            name = node.name
        elif (isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name)
              and node.targets[0].id in CONSTS):
            name = node.targets[0].id
        if name and name not in seen:
            seen.add(name)
            body.append(node)
    mod = types.ModuleType("ref_main14b_2")
    mod.__dict__.update(dict(torch=torch, nn=nn, F=F))
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), mod.__dict__)
    assert not (WANT | CONSTS) - seen, (WANT | CONSTS) - seen
    return mod


def checksums(sd):
    return {k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in sd.items()}


def main():
    ref = extract()
    torch.manual_seed(SEED_G)
    G = ref.Generator().eval()
    torch.manual_seed(SEED_D)
    D = ref.Detector().eval()
    g = torch.Generator().manual_seed(77)
    s = (0.1 * torch.randn(3, 1, 16000, generator=g)).clamp(-0.99, 0.99)
    s_short = (0.1 * torch.randn(2, 1, 5003, generator=g)).clamp(-0.99, 0.99)     # exercises the crop / pad branch
    msg = torch.tensor([0, 40000, 65535])
    out = {"s": s.numpy(), "s_short": s_short.numpy(), "messages": msg.numpy(),
           "seed_g": np.array(SEED_G), "seed_d": np.array(SEED_D)}
    with torch.no_grad():
        out["delta"] = G(s, msg).numpy()
        out["delta_nomsg"] = G(s[:1]).numpy()
        out["delta_short"] = G(s_short, msg[:2]).numpy()
        out["logits"] = D(s[:2]).numpy()
        out["logits_short"] = D(s_short[:1]).numpy()
    for k, v in checksums(G.state_dict()).items():
        out["gsum/" + k] = v
    for k, v in checksums(D.state_dict()).items():
        out["dsum/" + k] = v
    np.savez_compressed(os.path.join(HERE, "main14b2_io.npz"), **out)
    print({k: v.shape for k, v in out.items() if not k.startswith(("gsum", "dsum"))})


if __name__ == "__main__":
    main()
