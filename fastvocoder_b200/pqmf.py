"""4-band pseudo-QMF filterbank (drop-in for model/generator/pqmf.py).

Filter design stays on the host in float64 and is cast to fp32 exactly like the
reference (pqmf.py:61-92), so the coefficient bytes are identical (sha256 pinned
in tests/golden/specs.json); analysis / synthesis run as CUDA kernels through
the C ABI (fv_pqmf_analysis / fv_pqmf_synthesis) with exact index arithmetic.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.signal.windows import kaiser

from . import _lib


def design_prototype_filter(taps: int = 62, cutoff_ratio: float = 0.142, beta: float = 9.0) -> np.ndarray:
    """Kaiser-windowed ideal low-pass, taps+1 coefficients, float64 (pqmf.py:15-48)."""
    assert taps % 2 == 0, "The number of taps mush be even number."
    assert 0.0 < cutoff_ratio < 1.0, "Cutoff ratio must be > 0.0 and < 1.0."
    n = np.arange(taps + 1) - 0.5 * taps
    omega_c = np.pi * cutoff_ratio
    with np.errstate(invalid="ignore", divide="ignore"):
        ideal = np.sin(omega_c * n) / (np.pi * n)
    ideal[taps // 2] = np.cos(0) * cutoff_ratio        # limit at n = 0
    return ideal * kaiser(taps + 1, beta)


def design_filters(subbands: int = 4, taps: int = 62, cutoff_ratio: float = 0.142, beta: float = 9.0):
    """Cosine-modulated analysis [S,1,taps+1] and synthesis [1,S,taps+1] banks as fp32 tensors."""
    proto = design_prototype_filter(taps, cutoff_ratio, beta)
    h_ana = np.zeros((subbands, taps + 1))
    h_syn = np.zeros((subbands, taps + 1))
    pos = np.arange(taps + 1) - (taps / 2)
    for k in range(subbands):
        mod = (2 * k + 1) * (np.pi / (2 * subbands)) * pos
        phase = (-1) ** k * np.pi / 4
        h_ana[k] = 2 * proto * np.cos(mod + phase)
        h_syn[k] = 2 * proto * np.cos(mod - phase)
    return (torch.from_numpy(h_ana).float().unsqueeze(1), torch.from_numpy(h_syn).float().unsqueeze(0))


class PQMF(torch.nn.Module):
    """Same constructor, buffers and methods as the reference PQMF (pqmf.py:51-135)."""

    def __init__(self, subbands=4, taps=62, cutoff_ratio=0.142, beta=9.0):
        super().__init__()
        ana, syn = design_filters(subbands, taps, cutoff_ratio, beta)
        self.register_buffer("analysis_filter", ana)
        self.register_buffer("synthesis_filter", syn)
        updown = torch.zeros((subbands, subbands, subbands)).float()
        for k in range(subbands):
            updown[k, k, 0] = 1.0
        self.register_buffer("updown_filter", updown)   # kept for state_dict compatibility
        self.subbands = subbands
        self.taps = taps

    def _check(self, x):
        if not x.is_cuda:
            raise _lib.FvError("PQMF runs on CUDA only (no CPU fallback); move the module and input to a GPU")
        if self.analysis_filter.device != x.device:
            raise _lib.FvError("PQMF buffers and input are on different devices")
        return x.contiguous().float()

    def analysis(self, x):
        """(B, 1, T) -> (B, subbands, T // subbands)   (pqmf.py:108-119)"""
        x = self._check(x)
        B, one, L = x.shape
        assert one == 1
        Lb = (L - self.subbands) // self.subbands + 1
        y = torch.empty(B, self.subbands, Lb, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().fv_pqmf_analysis(_lib.ptr(x), _lib.ptr(self.analysis_filter), B, self.subbands,
                                                   self.taps, L, _lib.ptr(y), _lib.current_stream_ptr()),
                       "fv_pqmf_analysis")
        return y

    def synthesis(self, x):
        """(B, subbands, T // subbands) -> (B, 1, T)   (pqmf.py:121-135)"""
        x = self._check(x)
        B, S, Lb = x.shape
        assert S == self.subbands
        y = torch.empty(B, 1, Lb * S, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().fv_pqmf_synthesis(_lib.ptr(x), _lib.ptr(self.synthesis_filter), B, S, self.taps,
                                                    Lb, _lib.ptr(y), _lib.current_stream_ptr()),
                       "fv_pqmf_synthesis")
        return y
