"""ctypes binding of libfastvocoder_b200.so (include/fastvocoder_b200.h).

There is deliberately no fallback: if the shared library is missing the import
of any compute entry point raises, and every compute call needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FV_LIB=<path>: load an experimental variant built with `python -m fastvocoder_b200.build -D... --out <path>` (A/B runs)
LIB_PATH = os.environ.get("FV_LIB") or os.path.join(_HERE, "_C", "libfastvocoder_b200.so")

FV_MAX_STAGES, FV_MAX_BRANCH, FV_MAX_DIL = 8, 8, 8
FV_HIFIGAN, FV_MB_HIFIGAN, FV_MELGAN, FV_BASIS_MELGAN = 0, 1, 2, 3
FV_FWD_DEFAULT, FV_FWD_BASIS_INFERENCE, FV_FWD_NO_TENSOR_CORES = 0, 1, 2
FV_ABI_VERSION = 1

KIND_BY_NAME = {  # bin/synthesize.py:25-68 model names
    "hifigan": FV_HIFIGAN,
    "multiband-hifigan": FV_MB_HIFIGAN,
    "melgan": FV_MELGAN,
    "basis-melgan": FV_BASIS_MELGAN,
}


class FvConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("in_channels", C.c_int32),
        ("bias", C.c_int32),
        ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * FV_MAX_STAGES),
        ("upsample_kernel_sizes", C.c_int32 * FV_MAX_STAGES),
        ("channels", C.c_int32 * (FV_MAX_STAGES + 1)),
        ("pre_kernel_size", C.c_int32),
        ("post_kernel_size", C.c_int32),
        ("out_channels", C.c_int32),
        ("num_kernels", C.c_int32),
        ("resblock_type", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * FV_MAX_BRANCH),
        ("resblock_num_dilations", C.c_int32 * FV_MAX_BRANCH),
        ("resblock_dilations", (C.c_int32 * FV_MAX_DIL) * FV_MAX_BRANCH),
        ("stacks", C.c_int32),
        ("stack_kernel_size", C.c_int32),
        ("use_final_activation", C.c_int32),
        ("basis_L", C.c_int32),
        ("pqmf_subbands", C.c_int32),
        ("pqmf_taps", C.c_int32),
        ("upsample_layer", C.c_int32),
        ("use_causal_conv", C.c_int32),
        ("lastlinear", C.c_int32),
        ("negative_slope_set", C.c_int32),
        ("negative_slope", C.c_float),
        ("reserved", C.c_int32 * 3),
    ]


class FvProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("kernel", C.c_int32), ("Cin", C.c_int32), ("N", C.c_int32),
                ("K", C.c_int32), ("dil", C.c_int32), ("positions", C.c_int64), ("flops", C.c_double),
                ("bytes", C.c_double), ("ms", C.c_float), ("reserved", C.c_float)]


# name -> (restype, argtypes); exactly the symbols include/fastvocoder_b200.h declares
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_I = C.c_int
_F = C.c_float
SIGNATURES = {
    "fv_last_error": (C.c_char_p, []),
    "fv_abi_version": (_I, []),
    "fv_tc_usable": (C.c_int, [C.c_void_p]),
    "fv_launch_count": (C.c_int64, []),
    "fv_tc_launch_count": (C.c_int64, []),
    "fv_create": (_I, [C.POINTER(FvConfig), C.POINTER(_P)]),
    "fv_destroy": (None, [_P]),
    "fv_num_params": (_I, [_P]),
    "fv_param_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(C.c_int64), C.POINTER(_I), C.POINTER(C.c_int64)]),
    "fv_param_total_floats": (C.c_int64, [_P]),
    "fv_bind_weights": (_I, [_P, _P, C.c_int64, _P, _P, _P]),
    "fv_out_length": (_I, [_P, _I, _I, C.POINTER(C.c_int64)]),
    "fv_workspace_bytes": (_I, [_P, _I, _I, C.POINTER(C.c_size_t)]),
    "fv_forward": (_I, [_P, _P, _I, _I, _P, _P, _P, C.c_size_t, _I, _P]),
    "fv_forward_ragged": (_I, [_P, _P, _I, _I, C.POINTER(C.c_int32), _P, _P, _P, C.c_size_t, _I, _P]),
    "fv_forward_flops": (_I, [_P, _I, _I, _I, C.POINTER(C.c_double)]),
    "fv_forward_profile": (_I, [_P, _P, _I, _I, _P, _P, _P, C.c_size_t, _I, _P, C.POINTER(FvProfileEntry), _I,
                                C.POINTER(_I)]),
    "fv_conv1d": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P]),
    "fv_conv_transpose1d": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P]),
    "fv_resblock1": (_I, [_P, _PP, _PP, _PP, _PP, C.POINTER(_I), _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "fv_residual_stack": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "fv_basis_signal": (_I, [_P, _P, _I, _I, _P, _I, _P]),
    "fv_overlap_add": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "fv_pqmf_synthesis": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "fv_pqmf_analysis": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "fv_encode_16bits": (_I, [_P, C.c_int64, _F, _P, _P, _P]),
}

_lib = None


class FvError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FvError(
                f"{LIB_PATH} is missing: build it with `python -m fastvocoder_b200.build` "
                "(there is no CPU / PyTorch fallback for the generator forward path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.fv_abi_version() != FV_ABI_VERSION:
            raise FvError(f"ABI mismatch: library {L.fv_abi_version()} vs binding {FV_ABI_VERSION}")
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().fv_last_error().decode("utf-8", "replace")
        raise FvError(f"{what or 'fastvocoder_b200'} failed (code {rc}): {msg}")


def ptr(t):
    """Raw device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
