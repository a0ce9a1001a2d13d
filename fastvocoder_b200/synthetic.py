"""Deterministic synthetic weights and mel inputs.

The reference ships no checkpoint (README.md:18 is a release URL) and there is
no network, so benchmarks and parity tests run on seeded synthetic weights of
the exact reference architecture.  The reference's own random init makes the
waveform tiny (|y| ~ 0.004-0.05, SURVEY.md §7 hard part 1), which would make a
1e-4 max-abs parity bar meaningless, so this generator scales every layer by
its fan-in such that activations stay O(1) and the waveform peaks near 0.5-0.9.

Everything is numpy PCG64 (bit-stable across machines), so the build container
(where the reference is imported to make golden outputs) and the GPU box
regenerate identical weights from (spec, seed) without shipping them.
"""
from __future__ import annotations

import re
import zlib

import numpy as np

_CONVT_MELGAN = re.compile(r"^melgan\.(\d+)\.weight$")


def _is_conv_transpose(name: str) -> bool:
    if name.startswith("ups.") and name.endswith(".weight") and ".conv." not in name:
        return True
    m = _CONVT_MELGAN.match(name)
    return bool(m) and int(m.group(1)) != 1


def _gain(name: str, shape=()) -> float:
    """Per-layer gain (found empirically so that the four shipped configs stay O(1))."""
    if ".conv.weight" in name and len(shape) == 3 and shape[0] > 4 and ".stack." not in name:
        return 1.3                                                    # UpsampleLayer's conv (modules.py:173)
    if name.startswith("conv_post"):                                  # last conv -> pre-tanh scale
        return 1.0
    if ".conv.weight" in name:                                        # MelGAN LastLayer
        return 0.5
    if ".convs2." in name or ".convs." in name:                       # residual branch output (ResBlock1 / ResBlock2)
        return 0.5
    if ".stack.4." in name:
        return 0.8
    if ".skip_layer." in name:
        return 0.95
    if name.startswith("basis_signal"):
        return 0.2
    return 1.3


def synth_param(name: str, shape, seed: int) -> np.ndarray:
    """One tensor, keyed by (seed, name) so that specs may be generated in any order."""
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
    shape = tuple(int(s) for s in shape)
    if ".bn_" in name:                                                # LastLinear's BatchNorm1d (modules.py:120-122)
        if name.endswith("running_var"):
            return rng.uniform(0.5, 1.5, shape).astype(np.float32)
        if name.endswith("running_mean") or name.endswith(".bias"):
            return (rng.standard_normal(shape) * 0.1).astype(np.float32)
        return (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)   # affine weight
    if name.endswith(".bias"):
        return (rng.standard_normal(shape) * 0.05).astype(np.float32)
    if len(shape) == 3:
        if _is_conv_transpose(name):
            fan_in = shape[0] * 2          # k = 2*stride: two taps per output sample
        else:
            fan_in = shape[1] * shape[2]
    elif len(shape) == 2:
        fan_in = shape[1]
    else:
        fan_in = 1
    std = _gain(name, shape) / np.sqrt(fan_in)
    return (rng.standard_normal(shape) * std).astype(np.float32)


def synth_state_dict(spec, seed: int = 0) -> dict:
    """spec: iterable of (name, shape) for the *folded* (weight-norm removed) state_dict."""
    return {name: synth_param(name, shape, seed) for name, shape in spec}


def synth_mel(batch: int, frames: int, seed: int = 0, n_mels: int = 80) -> np.ndarray:
    """Uniform [0,1) mel of shape [B, 80, T] (normalised-mel range, data/audio.py:159-160)."""
    rng = np.random.Generator(np.random.PCG64([seed, 0x6D656C]))
    return rng.random((batch, n_mels, frames), dtype=np.float32)
