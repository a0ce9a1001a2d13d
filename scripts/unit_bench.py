#!/usr/bin/env python
"""Time (and, with FV_STALL_DEBUG=1, stall-profile) ONE ResBlock1 branch (3 fused units) through fv_resblock1.
    python scripts/unit_bench.py C K L B mode     mode 2 = fp32 I/O fused units, 3 = TMA-fed split chain
Prints ms per unit (CUDA events, best of 5 after warm-up).  Planner knobs (FV_TC3_*) are read from the environment."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fastvocoder_b200 import _lib  # noqa: E402

C, K, L, B, mode = (int(v) for v in sys.argv[1:6])
rng = np.random.default_rng(1)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
x = dev(rng.standard_normal((B, C, L)).astype(np.float32))
w1 = [dev((rng.standard_normal((C, C, K)) * (0.6 / np.sqrt(C * K))).astype(np.float32)) for _ in range(3)]
w2 = [dev((rng.standard_normal((C, C, K)) * (0.6 / np.sqrt(C * K))).astype(np.float32)) for _ in range(3)]
b1 = [dev((rng.standard_normal(C) * 0.1).astype(np.float32)) for _ in range(3)]
b2 = [dev((rng.standard_normal(C) * 0.1).astype(np.float32)) for _ in range(3)]
arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
dil = (ctypes.c_int * 3)(1, 3, 5)
y = torch.empty(B, C, L, device="cuda")
scratch = torch.empty(2 * B * C * L, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    _lib.check(_lib.lib().fv_resblock1(_lib.ptr(x), arr(w1), arr(b1), arr(w2), arr(b2), dil, 3, _lib.ptr(y), _lib.ptr(scratch),
                                       B, C, L, K, mode, st))


run()
torch.cuda.synchronize()
if os.environ.get("FV_STALL_DEBUG"):
    sys.exit(0)
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
# fv_resblock1 derives weight images and synchronises per unit: the event time includes that host work; the per-unit kernel
# time is what FV_STALL_DEBUG reports.  Use for A/B only.
print(f"C={C} K={K} L={L} B={B} mode={mode} env={ {k: v for k, v in os.environ.items() if k.startswith('FV_')} }: {best / 3:.3f} ms per unit (incl. host overhead)")
