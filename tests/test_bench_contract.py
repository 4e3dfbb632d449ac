"""bench.py prints ONE JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                       timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the reference's own definitions (oracle/_ref, built from the mount by oracle/make_ref.py) or,
    without them, the oracle port, on the host cores; a bounded sample, the same metric and config."""
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "clip-s/s" and d["value"] > 0
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_main16.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "clip-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_own_arm_line():
    d = run_bench("--steps", "2", "--warmup", "3", "--batch", "256", "--no-cpu-baseline", "--no-aux")
    assert BASE_KEYS | {"roofline", "clocks", "gpu_launches", "parity", "aux"} <= set(d)
    par = d["parity"]
    assert par["delta_err"] < 1e-3 and par["prob_err"] < 1e-3 and par["bit_mismatches_safe"] == 0
    assert par["vote_mismatches_decidable"] == 0 and par["clips"] == 64
    assert d["aux"]["votes_step"]["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] == 2 * 10                       # ten kernels of this library per embed+detect step
    rf = d["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["unit"] == "TFLOP/s"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["vs_baseline"] is None and "workload" in d["config"]
