"""Diagnostic for the opt-in cluster/multicast weight ring (FV_CLUSTER=1): which half of each ring stage is missing?

C=128, k=3 -> k-block = 8 KB, 2 k-blocks per 16 KB stage: rank 0 fetches the even k-blocks (channels [32m, 32m+16)),
rank 1 the odd ones.  Compare the wrong outputs with partial sums over even-only / odd-only channels.
"""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvocoder_b200 import _lib
rng = np.random.default_rng(0)
B, C, K, d, L = 1, 128, 3, 1, 2048
x = rng.standard_normal((B, C, L)).astype(np.float32)
w = (rng.standard_normal((C, C, K)) / np.sqrt(C * K)).astype(np.float32)
def conv(wm):
    return torch.nn.functional.conv1d(torch.from_numpy(x).double(), torch.from_numpy(wm).double(), padding=1).numpy()
even = np.zeros(C, bool); even[[c for c in range(C) if (c // 16) % 2 == 0]] = True
w_even = w * even[None, :, None]; w_odd = w * (~even)[None, :, None]
want, want_even, want_odd = conv(w), conv(w_even), conv(w_odd)
dx, dw = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
y = torch.zeros(B, C, L, device="cuda")
_lib.check(_lib.lib().fv_conv1d(_lib.ptr(dx), _lib.ptr(dw), None, None, _lib.ptr(y), B, C, C, L, K, d, 0, -1.0, 0, 1, None))
got = y.cpu().numpy()
bad = ~(np.abs(got - want) < 1e-3)
print("nan", int(np.isnan(got).sum()), "bad", int(bad.sum()), "of", got.size)
pos_bad = bad[0].any(axis=0)
edges = np.nonzero(np.diff(pos_bad.astype(int)))[0]
print("bad position runs (edges):", edges[:16])
if pos_bad.any():
    sel = pos_bad
    for name, ref in (("even-channel half only", want_even), ("odd-channel half only", want_odd), ("zero", np.zeros_like(want))):
        print("  bad region vs %-24s max|diff| = %.3e" % (name, np.abs(got[0][:, sel] - ref[0][:, sel]).max()))
