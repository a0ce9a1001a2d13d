"""Build libfastvocoder_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m fastvocoder_b200.build [--force] [-v]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
OUT = os.path.join(OUT_DIR, "libfastvocoder_b200.so")
SOURCES = ["fv_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "fastvocoder_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defs=None, out: str | None = None) -> str:
    """Default: the in-tree library.  `defs` (e.g. ["-DFV_PACKED_F32=1"]) + `out` build an experimental variant next to it;
    load it with FV_LIB=<path> (fastvocoder_b200/_lib.py) to A/B two binaries inside one GPU call."""
    if out is None and not defs:
        if not force and not _stale():
            return OUT
    target = out or OUT
    os.makedirs(os.path.dirname(target), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(defs or []) + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + SOURCES
    r = subprocess.run(cmd, cwd=SRC_DIR, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return target


if __name__ == "__main__":
    argv = sys.argv[1:]
    defs = [a for a in argv if a.startswith("-D")]
    out = argv[argv.index("--out") + 1] if "--out" in argv else None
    print(build(force="--force" in argv, verbose="-v" in argv, defs=defs, out=out))
