"""Build libwmb200.so (sm_100a only) in-tree with nvcc.  `python build.py` or
`__graft_entry__.build()`; the .so lands next to this file so it travels with the
repo snapshot to the GPU box (it is git-ignored)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwmb200.so")
STAMP = os.path.join(HERE, "build", "stamp.txt")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"] + os.environ.get("WMB200_DEFS", "").split()


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:  # noqa
    h = hashlib.sha256()
    files = sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "wmb200.h"))
    for f in files:
        h.update(f.encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def _src_digest(src: str) -> str:
    """Digest of one translation unit: its source, every header of csrc/ and include/, and the flags."""
    h = hashlib.sha256()
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "wmb200.h"))
    for f in [src] + hdrs:
        h.update(f.encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a (one nvcc per file, in parallel, only the files whose digest changed)
    and link libwmb200.so.  build/stamp.txt holds the digest of the whole source set the library was linked from."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        tag, sd = obj + ".digest", _src_digest(src)
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read().strip() == sd:
            continue
        cmd = [nvcc, *ARCH, *[f for f in FLAGS if not f.startswith("--use_fast_math")], "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, tag, sd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = None
    for src, tag, sd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = failed or src
            continue
        with open(tag, "w") as f:
            f.write(sd)
    if failed:
        raise RuntimeError(f"nvcc failed on {failed}")
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart", "-lcuda"]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
