"""Diagnostic: run one ring-mode conv with FV_CLUSTER=1 and report where the output is wrong."""
import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvocoder_b200 import _lib
rng = np.random.default_rng(0)
B, C, K, d, L = 1, 128, 3, 1, 1024
x = rng.standard_normal((B, C, L)).astype(np.float32)
w = (rng.standard_normal((C, C, K)) / np.sqrt(C * K)).astype(np.float32)
want = torch.nn.functional.conv1d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), padding=1).numpy()
dx, dw = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
y = torch.zeros(B, C, L, device="cuda")
_lib.check(_lib.lib().fv_conv1d(_lib.ptr(dx), _lib.ptr(dw), None, None, _lib.ptr(y), B, C, C, L, K, d, 0, -1.0, 0, 1, None))
got = y.cpu().numpy()
err = np.abs(got - want)
bad = ~(err < 1e-3)
print("nan count", int(np.isnan(got).sum()), "bad count", int(bad.sum()), "of", got.size)
if bad.any():
    pos_bad = bad[0].any(axis=0)          # per position
    ch_bad = bad[0].any(axis=1)
    idx = np.nonzero(pos_bad)[0]
    print("bad positions: n=%d first=%d last=%d" % (idx.size, idx[0], idx[-1]), "runs:", np.nonzero(np.diff(pos_bad.astype(int)))[0][:20])
    print("bad channels n=%d" % ch_bad.sum(), np.nonzero(ch_bad)[0][:40])
    print("sample got/want", got[0, :4, idx[0]], want[0, :4, idx[0]])
