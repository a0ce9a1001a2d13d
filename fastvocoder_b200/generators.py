"""Drop-in generator classes: same constructors (YAML keys), state_dict keys, ``forward`` /
``inference`` / ``remove_weight_norm`` / ``apply_weight_norm`` as the reference's
model/generator/{hifigan,multiband_hifigan,melgan,basis_melgan}.py — but the forward path is the
hand-written sm_100a CUDA of libfastvocoder_b200.so reached through its C ABI.

There is no PyTorch / CPU fallback: ``forward`` on a CPU tensor raises.

Host-side responsibilities kept here (all load-time, none per step):
  * weight-norm folding ``w = g * v / ||v||`` with ``torch._weight_norm`` — the very function
    ``torch.nn.utils.remove_weight_norm`` uses (hifigan.py:58-67), so folded weights are bit-identical;
  * packing every folded parameter into ONE flat fp32 buffer (the unit a multi-GPU launch broadcasts
    once over NCCL, see sharding.py) at the offsets the C library reports;
  * allocating outputs / workspace as torch tensors.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .pqmf import PQMF

__all__ = ["HiFiGANGenerator", "MultiBandHiFiGANGenerator", "MelGANGenerator", "BasisMelGANGenerator",
           "build_generator"]


def _set_arr(arr, values):
    for i, v in enumerate(values):
        arr[i] = int(v)


class _NativeGenerator(torch.nn.Module):
    """Shared machinery; subclasses fill an FvConfig from their reference-style kwargs."""

    model_name = ""

    def __init__(self, cfg: _lib.FvConfig, use_weight_norm: bool = True):
        super().__init__()
        self._cfg = cfg
        self._handle = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.fv_create(C.byref(cfg), C.byref(self._handle)), "fv_create")
        self._spec = []  # (name, shape, offset)
        name_buf = C.create_string_buffer(256)
        shape = (C.c_int64 * 4)()
        ndim = C.c_int()
        off = C.c_int64()
        for i in range(L.fv_num_params(self._handle)):
            _lib.check(L.fv_param_info(self._handle, i, name_buf, 256, shape, C.byref(ndim), C.byref(off)))
            self._spec.append((name_buf.value.decode(), tuple(int(shape[d]) for d in range(ndim.value)),
                               int(off.value)))
        total = int(L.fv_param_total_floats(self._handle))
        self.register_buffer("packed_weights", torch.zeros(total, dtype=torch.float32), persistent=False)
        self._wn = OrderedDict()      # layer prefix -> (weight_g, weight_v): present while weight norm is applied
        self._bound_key = None
        self._workspace = None
        self.use_tensor_cores = True
        # inference(): replay a CUDA graph captured per input length instead of issuing the launch chain (Synthesizer /
        # test.sh turn it on: batch-1 calls are launch-bound)
        self.use_cuda_graphs = False
        self._reset_parameters()
        if use_weight_norm:
            self.apply_weight_norm()

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().fv_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    # ---- parameters --------------------------------------------------------------------------
    def _view(self, name):
        for n, shape, off in self._spec:
            if n == name:
                return self.packed_weights[off: off + int(np.prod(shape))].view(shape)
        raise KeyError(name)

    def _wn_layers(self):
        """Layers the reference wraps in weight norm: every Conv1d / ConvTranspose1d (hifigan.py:69-77)."""
        return [n[:-len(".weight")] for n, shape, _ in self._spec if n.endswith(".weight") and len(shape) == 3]

    def _reset_parameters(self):
        """Random init (PyTorch-default-like uniform(+-1/sqrt(fan_in))); real use loads a checkpoint."""
        with torch.no_grad():
            for name, shape, _ in self._spec:
                v = self._view(name)
                if name.endswith(".bias"):
                    w_shape = dict((n, s) for n, s, _ in self._spec)[name[:-5] + ".weight"]
                    bound = 1.0 / math.sqrt(w_shape[1] * w_shape[2])
                else:
                    bound = 1.0 / math.sqrt(shape[1] * (shape[2] if len(shape) == 3 else 1))
                v.uniform_(-bound, bound)

    def apply_weight_norm(self):
        """Re-parametrise conv weights as (g, v) like torch.nn.utils.weight_norm(dim=0) (hifigan.py:69-77)."""
        with torch.no_grad():
            for p in self._wn_layers():
                w = self._view(p + ".weight").detach().clone()
                g = torch.norm_except_dim(w, 2, 0)
                self._wn[p] = (g, w)

    def remove_weight_norm(self):
        """Fold g*v/||v|| into .weight and drop the re-parametrisation (hifigan.py:58-67)."""
        self._fold()
        self._wn.clear()

    def _fold(self):
        with torch.no_grad():
            for p, (g, v) in self._wn.items():
                w = torch._weight_norm(v.float().cpu(), g.float().cpu(), 0)
                self._view(p + ".weight").copy_(w)
        self._bound_key = None

    def reset_parameters(self):
        self._reset_parameters()
        if self._wn:
            self.apply_weight_norm()
        self._bound_key = None

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        """Reference key set: weight-norm form while it is applied, folded form after remove_weight_norm().
        Honours nn.Module's (destination, prefix) protocol, so a generator nested in another Module is saved too."""
        if args:                                  # legacy positional form: (destination, prefix, keep_vars)
            destination = args[0]
            prefix = args[1] if len(args) > 1 else prefix
        sd = self._own_state_dict()
        if destination is None:
            destination = OrderedDict()
        for k, v in sd.items():
            destination[prefix + k] = v
        return destination

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for k, v in self._own_state_dict().items():      # a parent Module's state_dict() reaches us through this hook
            destination[prefix + k] = v

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        own = OrderedDict((k[len(prefix):], v) for k, v in state_dict.items() if k.startswith(prefix))
        try:
            res = self.load_state_dict(own, strict=False)
            missing_keys.extend(prefix + k for k in res.missing_keys)
            unexpected_keys.extend(prefix + k for k in res.unexpected_keys)
        except RuntimeError as e:
            error_msgs.append(str(e))

    def _apply(self, fn, recurse=True):
        """.to(device) moves the packed weights; dtype casts (.half() / .double()) are refused for them: the kernels read the
        buffer as float32 (the arithmetic type of this path)."""
        super()._apply(fn, recurse)
        if self.packed_weights.dtype != torch.float32:
            self.packed_weights.data = self.packed_weights.data.float()
        self._bound_key = None
        return self

    def _own_state_dict(self):
        sd = OrderedDict()
        for name, shape, _ in self._spec:
            prefix = name[:-len(".weight")] if name.endswith(".weight") else None
            if prefix is not None and prefix in self._wn:
                g, v = self._wn[prefix]
                sd[prefix + ".weight_g"] = g.detach().clone()
                sd[prefix + ".weight_v"] = v.detach().clone()
            else:
                sd[name] = self._view(name).detach().clone()
        for k, v in self._extra_state_tensors().items():
            sd[k] = v.detach().clone()
        return sd

    def _extra_state_tensors(self):
        return {}

    def load_state_dict(self, state_dict, strict: bool = True):
        """Accepts reference checkpoints in either form (``*.weight_g/_v`` or folded ``*.weight``)."""
        missing, used = [], set()
        new_wn = OrderedDict()
        with torch.no_grad():
            for name, shape, _ in self._spec:
                prefix = name[:-len(".weight")] if name.endswith(".weight") else None
                if name in state_dict:
                    t = torch.as_tensor(state_dict[name]).detach().float().cpu()
                    if tuple(t.shape) != shape:
                        raise RuntimeError(f"size mismatch for {name}: checkpoint {tuple(t.shape)} vs model {shape}")
                    self._view(name).copy_(t)
                    used.add(name)
                elif prefix is not None and prefix + ".weight_g" in state_dict and prefix + ".weight_v" in state_dict:
                    g = torch.as_tensor(state_dict[prefix + ".weight_g"]).detach().float().cpu()
                    v = torch.as_tensor(state_dict[prefix + ".weight_v"]).detach().float().cpu()
                    if tuple(v.shape) != shape:
                        raise RuntimeError(f"size mismatch for {prefix}.weight_v: {tuple(v.shape)} vs {shape}")
                    new_wn[prefix] = (g, v)
                    used.update((prefix + ".weight_g", prefix + ".weight_v"))
                else:
                    missing.append(name)
        unexpected = [k for k in state_dict if k not in used and not k.startswith("pqmf.")]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for {type(self).__name__}: "
                               f"missing keys {missing}, unexpected keys {unexpected}")
        self._wn = new_wn
        self._fold()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    # ---- device binding -------------------------------------------------------------------------
    @property
    def device(self):
        return self.packed_weights.device

    def _pqmf_ptrs(self):
        return None, None

    def _ensure_bound(self):
        pw = self.packed_weights
        if pw.dtype != torch.float32:
            raise _lib.FvError(f"{type(self).__name__}: packed weights are {pw.dtype}; the kernels read float32")
        if not pw.is_cuda:
            raise _lib.FvError(f"{type(self).__name__}: weights are on {pw.device}; this implementation has no CPU "
                               "path — call .to('cuda') first")
        if self._wn and self._bound_key is None:
            self._fold()
        key = (pw.data_ptr(), pw._version, pw.device.index)
        if key != self._bound_key:
            ana, syn = self._pqmf_ptrs()
            with torch.cuda.device(pw.device):
                _lib.check(_lib.lib().fv_bind_weights(self._handle, _lib.ptr(pw), pw.numel(), ana, syn,
                                                      _lib.current_stream_ptr()), "fv_bind_weights")
            self._bound_key = key

    @property
    def tensor_cores_usable(self) -> bool:
        """False when a bound weight does not fit the fp16 hi/lo split (|w| > 65504 or non-finite): the library then dropped the
        tensor-core images and every layer runs on the exact-fp32 kernels (fv_tc_usable, include/fastvocoder_b200.h)."""
        self._ensure_bound()
        return bool(_lib.lib().fv_tc_usable(self._handle))

    def _get_workspace(self, B, T):
        """One workspace per CUDA stream: two streams driving the same model never share activation buffers."""
        need = C.c_size_t()
        _lib.check(_lib.lib().fv_workspace_bytes(self._handle, B, T, C.byref(need)), "fv_workspace_bytes")
        if self._workspace is None:
            self._workspace = {}
        key = (self.device, torch.cuda.current_stream(self.device).cuda_stream)
        ws = self._workspace.get(key)
        if ws is None or ws.numel() < need.value:
            if len(self._workspace) >= 8:                 # bounded: drop the oldest stream's buffer
                self._workspace.pop(next(iter(self._workspace)))
            self._workspace[key] = ws = torch.empty(int(need.value), dtype=torch.uint8, device=self.device)
        return ws

    def out_length(self, T: int, flags: int = 0) -> int:
        n = C.c_int64()
        _lib.check(_lib.lib().fv_out_length(self._handle, int(T), int(flags), C.byref(n)), "fv_out_length")
        return int(n.value)

    def forward_flops(self, B: int, T: int, flags: int = 0) -> float:
        f = C.c_double()
        _lib.check(_lib.lib().fv_forward_flops(self._handle, int(B), int(T), int(flags), C.byref(f)))
        return float(f.value)

    def _run(self, x, out, out2, flags=0):
        B, _, T = x.shape
        ws = self._get_workspace(B, T)
        if not self.use_tensor_cores:
            flags |= _lib.FV_FWD_NO_TENSOR_CORES
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().fv_forward(self._handle, _lib.ptr(x), B, T, _lib.ptr(out), _lib.ptr(out2),
                                             _lib.ptr(ws), ws.numel(), flags, _lib.current_stream_ptr()),
                       "fv_forward")

    def inference_batch(self, mels):
        """Ragged batch: a list of (T_i, in_channels) mels (ndarray / tensor, lengths may differ) -> list of 1-D
        waveforms, utterance i being what ``inference(mels[i])`` returns — but ONE launch chain over the padded
        batch (fv_forward_ragged) instead of the per-file loop of bin/test.py:123-131.  Each kernel treats
        samples beyond an utterance's own length as sequence padding (zero / reflect about ITS end), so there
        are no edge artefacts from the padding.  Multi-band: PQMF-synthesised waveform; Basis-MelGAN:
        ``inference`` semantics (untruncated, no zero-input subtraction)."""
        if len(mels) == 0:
            return []
        self._ensure_bound()
        dev = self.device
        ts = [m if isinstance(m, torch.Tensor) else torch.as_tensor(np.asarray(m)) for m in mels]
        for t in ts:
            if t.dim() != 2 or t.shape[1] != self._cfg.in_channels:
                raise RuntimeError(f"expected (T, {self._cfg.in_channels}) mels, got {tuple(t.shape)}")
        lens = [int(t.shape[0]) for t in ts]
        B, T = len(ts), max(lens)
        x = torch.zeros(B, self._cfg.in_channels, T, dtype=torch.float32, device=dev)
        for i, t in enumerate(ts):
            x[i, :, : lens[i]] = t.to(dev).float().transpose(0, 1)
        kind = self._cfg.kind
        flags = _lib.FV_FWD_BASIS_INFERENCE if kind == _lib.FV_BASIS_MELGAN else 0
        if not self.use_tensor_cores:
            flags |= _lib.FV_FWD_NO_TENSOR_CORES
        n = self.out_length(T, flags)
        oc = self._cfg.out_channels if kind in (_lib.FV_MB_HIFIGAN, _lib.FV_MELGAN) else 1
        out = torch.empty(B, oc, n, device=dev, dtype=torch.float32)
        wav = torch.empty(B, 1, oc * n, device=dev, dtype=torch.float32) if kind == _lib.FV_MB_HIFIGAN else None
        ws = self._get_workspace(B, T)
        lens_c = (C.c_int32 * B)(*lens)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fv_forward_ragged(self._handle, _lib.ptr(x), B, T, lens_c, _lib.ptr(out),
                                                    _lib.ptr(wav), _lib.ptr(ws), ws.numel(), flags,
                                                    _lib.current_stream_ptr()), "fv_forward_ragged")
        res = []
        for i, l in enumerate(lens):
            ni = self.out_length(l, flags)
            if kind == _lib.FV_MB_HIFIGAN:
                res.append(wav[i, 0, : oc * ni].clone())
            else:
                res.append(out[i, 0, :ni].clone())
        return res

    def profile_forward(self, x, flags: int = 0):
        """One forward with per-layer CUDA-event timing (fv_forward_profile). Returns a list of dicts."""
        x = self._prep_input(x)
        B, _, T = x.shape
        ws = self._get_workspace(B, T)
        out = torch.empty(B * max(1, self._cfg.out_channels if self._cfg.kind != _lib.FV_BASIS_MELGAN else 1)
                          * self.out_length(T, flags), device=x.device, dtype=torch.float32)
        if not self.use_tensor_cores:
            flags |= _lib.FV_FWD_NO_TENSOR_CORES
        cap = 1024
        entries = (_lib.FvProfileEntry * cap)()
        n = C.c_int()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().fv_forward_profile(self._handle, _lib.ptr(x), B, T, _lib.ptr(out), None,
                                                     _lib.ptr(ws), ws.numel(), flags, _lib.current_stream_ptr(),
                                                     entries, cap, C.byref(n)), "fv_forward_profile")
        kn = {0: "ffma", 1: "tcgen05", 2: "tcgen05-fused-unit", 3: "tcgen05-fused-stack"}
        return [dict(name=e.name.decode(), kernel=kn.get(e.kernel, str(e.kernel)), Cin=e.Cin, N=e.N, K=e.K,
                     dil=e.dil, positions=e.positions, flops=e.flops, bytes=e.bytes, ms=e.ms)
                for e in entries[: n.value]]

    def _prep_input(self, x):
        if not isinstance(x, torch.Tensor):
            raise TypeError("forward expects a torch.Tensor [B, in_channels, T]")
        self._ensure_bound()
        if x.device != self.device:
            raise _lib.FvError(f"input on {x.device} but model on {self.device} (no CPU fallback)")
        if x.dim() != 3 or x.shape[1] != self._cfg.in_channels:
            raise RuntimeError(f"expected input [B, {self._cfg.in_channels}, T], got {tuple(x.shape)}")
        return x.detach().contiguous().float()

    # ---- CUDA-graph replay of the launch chain (batch-1 latency path) -----------------------------------------
    def graphed(self, example, **fwd_kwargs):
        """Capture ``forward(example, **fwd_kwargs)`` into a CUDA graph and return ``run(x)``, which copies ``x`` into the
        captured input, replays the whole launch chain (~50-100 dependent kernels) with ONE graph launch and returns the
        captured outputs (valid until the next ``run``; clone them to keep them).  Same kernels, same plans, same
        buffers -> bit-identical to the eager call.  The graph is tied to the input shape and to the current weights."""
        x = self._prep_input(example)
        static_in = x.clone()
        cur = torch.cuda.current_stream(x.device)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():      # warm-up: weight binding, workspace, lazy attributes
            for _ in range(2):
                self.forward(static_in, **fwd_kwargs)
        cur.wait_stream(side)
        torch.cuda.synchronize(x.device)
        graph = torch.cuda.CUDAGraph()
        key = self._bound_key
        with torch.cuda.graph(graph), torch.no_grad():
            static_out = self.forward(static_in, **fwd_kwargs)

        def run(inp):
            if self._bound_key != key or self._bound_key is None:
                raise _lib.FvError("graphed(): the weights changed after capture; capture again")
            static_in.copy_(inp if inp.shape == static_in.shape else self._prep_input(inp))
            graph.replay()
            return static_out
        run.graph, run.static_input, run.static_output = graph, static_in, static_out
        return run

    def _graphed_inference(self, x, **fwd_kwargs):
        """inference() through a per-(shape, kwargs) cache of captured graphs (``use_cuda_graphs``); outputs are cloned."""
        self._ensure_bound()
        key = (tuple(x.shape), tuple(sorted(fwd_kwargs.items())), self._bound_key)
        cache = self.__dict__.setdefault("_graphs", OrderedDict())
        g = cache.get(key)
        if g is None:
            if len(cache) >= 16:
                cache.popitem(last=False)
            g = cache[key] = self.graphed(x, **fwd_kwargs)
        out = g(x)
        return tuple(o.clone() if o is not None else None for o in out) if isinstance(out, tuple) else out.clone()

    def _prep_inference_input(self, c):
        """(T, in_channels) ndarray or tensor -> (1, in_channels, T) on the model device (hifigan.py:111-113)."""
        if not isinstance(c, torch.Tensor):
            c = torch.tensor(c, dtype=torch.float).to(self.device)
        return c.to(self.device).float().transpose(1, 0).unsqueeze(0)


def _fill_hifi(cfg, kind, resblock_kernel_sizes, upsample_rates, upsample_initial_channel, resblock_type,
               upsample_kernel_sizes, resblock_dilation_sizes, transposedconv, bias, out_channels):
    cfg.kind = kind
    cfg.in_channels = 80
    # hifigan.py:31-46: `transposedconv == False` -> UpsampleLayer (nearest stretch + Conv1d(k, padding k//2))
    cfg.upsample_layer = 1 if transposedconv == False else 0   # noqa: E712  (the reference's own comparison)
    cfg.bias = 1 if bias else 0
    cfg.num_upsamples = len(upsample_rates)
    if len(upsample_kernel_sizes) != len(upsample_rates):
        raise ValueError("upsample_rates and upsample_kernel_sizes differ in length")
    _set_arr(cfg.upsample_rates, upsample_rates)
    _set_arr(cfg.upsample_kernel_sizes, upsample_kernel_sizes)
    _set_arr(cfg.channels, [upsample_initial_channel // (2 ** i) for i in range(len(upsample_rates) + 1)])
    cfg.pre_kernel_size = 7
    cfg.post_kernel_size = 7
    cfg.out_channels = out_channels
    cfg.num_kernels = len(resblock_kernel_sizes)
    # hifigan.py:28: `ResBlock1 if resblock_type == '1' else ResBlock2` — the reference's own comparison, so an unquoted
    # YAML 1 (int) selects ResBlock2 exactly as it does there.
    cfg.resblock_type = 1 if resblock_type == '1' else 2
    _set_arr(cfg.resblock_kernel_sizes, resblock_kernel_sizes)
    if len(resblock_dilation_sizes) < len(resblock_kernel_sizes):
        raise ValueError("resblock_dilation_sizes needs one entry per resblock kernel size")
    # ResBlock1 hard-codes three (convs1, convs2) pairs from dilation[0..2], ResBlock2 two convs from dilation[0..1]
    # (modules.py:190-251) whatever the list length: extra entries are ignored, missing ones raise IndexError there.
    n_convs = 3 if cfg.resblock_type == 1 else 2
    for j, dils in enumerate(resblock_dilation_sizes[: len(resblock_kernel_sizes)]):
        if len(dils) < n_convs:
            raise IndexError("tuple index out of range")   # what dilation[n] raises in the reference constructor
        cfg.resblock_num_dilations[j] = n_convs
        _set_arr(cfg.resblock_dilations[j], dils[:n_convs])
    cfg.use_final_activation = 1
    return cfg


class HiFiGANGenerator(_NativeGenerator):
    """Drop-in for model/generator/hifigan.py:13-129."""

    model_name = "hifigan"

    def __init__(self, resblock_kernel_sizes=[3, 7, 11], upsample_rates=[8, 5, 3, 2], upsample_initial_channel=256,
                 resblock_type="1", upsample_kernel_sizes=[16, 10, 6, 4],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], transposedconv=True, bias=True):
        cfg = _fill_hifi(_lib.FvConfig(), _lib.FV_HIFIGAN, resblock_kernel_sizes, upsample_rates,
                         upsample_initial_channel, resblock_type, upsample_kernel_sizes, resblock_dilation_sizes,
                         transposedconv, bias, out_channels=1)
        super().__init__(cfg)
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)

    def forward(self, x):
        """[B, 80, T] -> [B, prod(rates) * T]   (hifigan.py:92-108)"""
        x = self._prep_input(x)
        B, _, T = x.shape
        out = torch.empty(B, self.out_length(T), device=x.device, dtype=torch.float32)
        self._run(x, out, None)
        return out

    def inference(self, x):
        """[T, 80] ndarray/tensor -> 1-D waveform   (hifigan.py:110-129)"""
        x = self._prep_inference_input(x)
        if self.use_cuda_graphs:
            return self._graphed_inference(self._prep_input(x)).squeeze()
        return self.forward(x).squeeze()


class MultiBandHiFiGANGenerator(_NativeGenerator):
    """Drop-in for model/generator/multiband_hifigan.py:14-137 (4 sub-bands + PQMF synthesis in inference)."""

    model_name = "multiband-hifigan"

    def __init__(self, resblock_kernel_sizes=[3, 7, 11], upsample_rates=[10, 6], upsample_initial_channel=256,
                 resblock_type="1", upsample_kernel_sizes=[20, 12],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], transposedconv=True, bias=True):
        cfg = _fill_hifi(_lib.FvConfig(), _lib.FV_MB_HIFIGAN, resblock_kernel_sizes, upsample_rates,
                         upsample_initial_channel, resblock_type, upsample_kernel_sizes, resblock_dilation_sizes,
                         transposedconv, bias, out_channels=4)
        cfg.pqmf_subbands = 4
        cfg.pqmf_taps = 62
        super().__init__(cfg)
        self.pqmf = PQMF()  # 4 band
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)

    def _extra_state_tensors(self):
        return {"pqmf.analysis_filter": self.pqmf.analysis_filter, "pqmf.synthesis_filter": self.pqmf.synthesis_filter,
                "pqmf.updown_filter": self.pqmf.updown_filter}

    def _pqmf_ptrs(self):
        if self.pqmf.synthesis_filter.device != self.device:
            self.pqmf.to(self.device)
        return _lib.ptr(self.pqmf.analysis_filter), _lib.ptr(self.pqmf.synthesis_filter)

    def forward(self, x, synthesize=False):
        """[B, 80, T] -> sub-bands [B, 4, L]   (multiband_hifigan.py:101-116).
        ``synthesize=True`` additionally returns the PQMF-synthesised waveform [B, 1, 4L] from the same call."""
        x = self._prep_input(x)
        B, _, T = x.shape
        Lb = self.out_length(T)
        out = torch.empty(B, 4, Lb, device=x.device, dtype=torch.float32)
        wav = torch.empty(B, 1, 4 * Lb, device=x.device, dtype=torch.float32) if synthesize else None
        self._run(x, out, wav)
        return (out, wav) if synthesize else out

    def inference(self, x):
        """[T, 80] -> 1-D waveform through PQMF synthesis   (multiband_hifigan.py:118-137)"""
        x = self._prep_inference_input(x)
        if self.use_cuda_graphs:
            _, wav = self._graphed_inference(self._prep_input(x), synthesize=True)
        else:
            _, wav = self.forward(x, synthesize=True)
        return wav.squeeze()


def _fill_melgan(cfg, kind, in_channels, out_channels, kernel_size, channels, upsample_scales, stack_kernel_size,
                 stacks, use_final_nonlinear_activation, use_causal_conv, nonlinear_activation,
                 nonlinear_activation_params, pad, bias=True):
    if nonlinear_activation != "LeakyReLU" or pad != "ReflectionPad1d":
        raise NotImplementedError("only LeakyReLU + ReflectionPad1d (the shipped configs) are implemented")
    if not use_causal_conv:
        assert (kernel_size - 1) % 2 == 0, "Not support even number kernel size."   # melgan.py:63-64
    elif (kernel_size - 1) % 2 != 0:
        raise NotImplementedError("use_causal_conv with an even kernel_size changes the sequence length in the "
                                  "first / last layer (melgan.py:68-71); not wired")
    cfg.use_causal_conv = 1 if use_causal_conv else 0
    if len(channels) != len(upsample_scales) + 1:
        raise ValueError("channels must have len(upsample_scales) + 1 entries")
    slope = float(nonlinear_activation_params.get("negative_slope", 0.01))   # nn.LeakyReLU's own default
    if not (0.0 <= slope <= 1.0):
        raise NotImplementedError("negative_slope outside [0, 1] is not implemented (LeakyReLU as max(x, slope*x))")
    cfg.negative_slope_set = 1
    cfg.negative_slope = slope
    cfg.kind = kind
    cfg.in_channels = in_channels
    cfg.bias = 1 if bias else 0
    cfg.num_upsamples = len(upsample_scales)
    _set_arr(cfg.upsample_rates, upsample_scales)
    _set_arr(cfg.upsample_kernel_sizes, [2 * s for s in upsample_scales])   # melgan.py:81
    _set_arr(cfg.channels, channels)
    cfg.pre_kernel_size = kernel_size
    cfg.post_kernel_size = kernel_size
    cfg.out_channels = out_channels
    cfg.stacks = stacks
    cfg.stack_kernel_size = stack_kernel_size
    cfg.use_final_activation = 1 if use_final_nonlinear_activation else 0
    return cfg


class MelGANGenerator(_NativeGenerator):
    """Drop-in for model/generator/melgan.py:17-185."""

    model_name = "melgan"

    def __init__(self, in_channels=80, out_channels=1, kernel_size=7, channels=[512, 256, 128, 64, 32], bias=True,
                 upsample_scales=[10, 6, 2, 2], stack_kernel_size=3, stacks=3, nonlinear_activation="LeakyReLU",
                 nonlinear_activation_params={"negative_slope": 0.2}, pad="ReflectionPad1d", pad_params={},
                 use_final_nonlinear_activation=True, use_weight_norm=True, use_causal_conv=False):
        cfg = _fill_melgan(_lib.FvConfig(), _lib.FV_MELGAN, in_channels, out_channels, kernel_size, channels,
                           upsample_scales, stack_kernel_size, stacks, use_final_nonlinear_activation,
                           use_causal_conv, nonlinear_activation, nonlinear_activation_params, pad, bias=bias)
        super().__init__(cfg, use_weight_norm=use_weight_norm)
        self.pqmf = None

    def forward(self, c, all_channels=False):
        """[B, 80, T] -> [B, prod(scales) * T]   (melgan.py:125-136; channel 0 of the 1-channel output).
        ``all_channels=True`` returns the whole [B, out_channels, L] tensor (what ``inference`` squeezes, melgan.py:172-185)."""
        c = self._prep_input(c)
        B, _, T = c.shape
        Lo = self.out_length(T)
        oc = self._cfg.out_channels
        out = torch.empty(B, oc, Lo, device=c.device, dtype=torch.float32)
        self._run(c, out, None)
        return out if all_channels else out[:, 0, :]

    def inference(self, c):
        """[T, 80] -> waveform   (melgan.py:172-185)"""
        c = self._prep_input(self._prep_inference_input(c))
        if self.use_cuda_graphs:
            return self._graphed_inference(c, all_channels=True).squeeze()
        return self.forward(c, all_channels=True).squeeze()


class BasisMelGANGenerator(_NativeGenerator):
    """Drop-in for model/generator/basis_melgan.py:19-212."""

    model_name = "basis-melgan"

    def __init__(self, basis_signal_weight, L=30, in_channels=80, out_channels=256, kernel_size=7,
                 channels=[256, 256, 256], bias=True, upsample_scales=[4, 4], stack_kernel_size=3, stacks=3,
                 nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.2},
                 pad="ReflectionPad1d", pad_params={}, use_final_nonlinear_activation=True, use_weight_norm=True,
                 use_causal_conv=False, transposedconv=True, lastlinear=False):
        if not bias and lastlinear:
            raise NotImplementedError("lastlinear=True with bias=False: the folded eval-mode BatchNorm needs the bias slot")
        # without LastLinear the predictor's width is channels[-1] (basis_melgan.py:70-121 never uses out_channels)
        width = out_channels if lastlinear else channels[-1]
        cfg = _fill_melgan(_lib.FvConfig(), _lib.FV_BASIS_MELGAN, in_channels, width, kernel_size, channels,
                           upsample_scales, stack_kernel_size, stacks, use_final_nonlinear_activation,
                           use_causal_conv, nonlinear_activation, nonlinear_activation_params, pad, bias=bias)
        cfg.basis_L = L
        cfg.lastlinear = 1 if lastlinear else 0
        # LastLinear (modules.py:116-132) sits at index 2 + sum(2 + stacks) of the nn.Sequential
        self._ll = f"melgan.{2 + len(upsample_scales) * (2 + stacks)}" if lastlinear else None
        self._ll_folded = False
        # basis_melgan.py:82-99: `transposedconv == False` -> UpsampleLayer(k = 2*scale+1, padding = scale)
        cfg.upsample_layer = 1 if transposedconv == False else 0   # noqa: E712
        bw = torch.as_tensor(basis_signal_weight).float()
        if tuple(bw.shape) != (L, width):
            raise ValueError(f"basis_signal_weight must be [L={L}, {width}], got {tuple(bw.shape)}")
        super().__init__(cfg, use_weight_norm=use_weight_norm)
        if self._ll:   # fresh nn.BatchNorm1d state (identity up to eps)
            self._bn = OrderedDict()
            for q in ("bn_1", "bn_2"):
                C_ = channels[-1]
                self._bn[f"{self._ll}.{q}.weight"] = torch.ones(C_)
                self._bn[f"{self._ll}.{q}.bias"] = torch.zeros(C_)
                self._bn[f"{self._ll}.{q}.running_mean"] = torch.zeros(C_)
                self._bn[f"{self._ll}.{q}.running_var"] = torch.ones(C_)
                self._bn[f"{self._ll}.{q}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        with torch.no_grad():
            self._view("basis_signal.layer.weight").copy_(bw)   # nn.Linear weight, no weight norm (modules.py:260-261)
        self.L = L
        self.pqmf = None

    # ---- LastLinear: eval-mode BatchNorm1d folded into the 1x1 conv that follows it (host side, load time) ----------
    _LL_NAMES = ("linear_1.weight", "linear_1.bias", "linear_2.weight", "linear_2.bias")

    def _ll_unfold(self):
        """Put the raw (checkpoint) linear_1/linear_2 parameters back into the packed buffer."""
        if getattr(self, "_ll", None) and getattr(self, "_ll_folded", False):
            with torch.no_grad():
                for n in self._LL_NAMES:
                    self._view(f"{self._ll}.{n}").copy_(self._ll_raw[n])
            self._ll_folded = False

    def _fold(self):
        self._ll_unfold()
        super()._fold()                       # weight norm first: g*v/||v|| -> raw .weight
        if not getattr(self, "_ll", None) or not hasattr(self, "_bn"):
            return
        with torch.no_grad():
            self._ll_raw = {n: self._view(f"{self._ll}.{n}").detach().cpu().clone() for n in self._LL_NAMES}
            for q, lin in (("bn_1", "linear_1"), ("bn_2", "linear_2")):
                g = self._bn[f"{self._ll}.{q}.weight"].double().cpu()
                b = self._bn[f"{self._ll}.{q}.bias"].double().cpu()
                mu = self._bn[f"{self._ll}.{q}.running_mean"].double().cpu()
                var = self._bn[f"{self._ll}.{q}.running_var"].double().cpu()
                a = g / torch.sqrt(var + 1e-5)                       # bn(x) = a*x + c   (eps = nn.BatchNorm1d default)
                c = b - mu * a
                W = self._ll_raw[f"{lin}.weight"].double()[:, :, 0]   # [Cout, Cin]
                self._view(f"{self._ll}.{lin}.weight").copy_((W * a[None, :]).float().unsqueeze(-1))
                self._view(f"{self._ll}.{lin}.bias").copy_((self._ll_raw[f"{lin}.bias"].double() + W @ c).float())
        self._ll_folded = True
        self._bound_key = None

    def _reset_parameters(self):
        if getattr(self, "_ll_folded", False):
            self._ll_folded = False           # the views are about to be overwritten with fresh raw values
        super()._reset_parameters()

    def apply_weight_norm(self):
        self._ll_unfold()
        super().apply_weight_norm()

    def load_state_dict(self, state_dict, strict: bool = True):
        if getattr(self, "_ll", None):
            self._ll_unfold()
            state_dict = OrderedDict(state_dict)
            missing_bn = []
            for k in list(self._bn):
                if k in state_dict:
                    t = torch.as_tensor(state_dict.pop(k)).detach().cpu()
                    self._bn[k] = t.clone() if k.endswith("num_batches_tracked") else t.float().clone()
                elif not k.endswith("num_batches_tracked"):
                    missing_bn.append(k)
            if strict and missing_bn:
                raise RuntimeError(f"Error(s) in loading state_dict for {type(self).__name__}: missing keys {missing_bn}")
        return super().load_state_dict(state_dict, strict)

    def _own_state_dict(self):
        sd = super()._own_state_dict()
        if getattr(self, "_ll", None):
            if self._ll_folded:               # report checkpoint-form (un-folded) LastLinear parameters
                for n in self._LL_NAMES:
                    if f"{self._ll}.{n}" in sd:
                        sd[f"{self._ll}.{n}"] = self._ll_raw[n].clone()
            for k, v in self._bn.items():
                sd[k] = v.clone()
        return sd

    def forward(self, c, return_weight=True, _inference=False):
        """[B, 80, T] -> (est_source - zero_est [B, 16T*15], weight - zero_weight [B, 16T, C])
        (basis_melgan.py:140-162).  The input-independent zero pass rides along as one extra utterance.
        ``_inference=True`` (used by ``inference``): one pass, no subtraction, untruncated [B, (16T+1)*15]."""
        c = self._prep_input(c)
        B, _, T = c.shape
        if _inference:
            n = self.out_length(T, _lib.FV_FWD_BASIS_INFERENCE)
            out = torch.empty(B, n, device=c.device, dtype=torch.float32)
            self._run(c, out, None, flags=_lib.FV_FWD_BASIS_INFERENCE)
            return out
        n = self.out_length(T)
        hop = self.L // 2
        est = torch.empty(B, n, device=c.device, dtype=torch.float32)
        weight = torch.empty(B, n // hop, self._cfg.out_channels, device=c.device,
                             dtype=torch.float32) if return_weight else None
        self._run(c, est, weight)
        return est, weight

    def inference(self, c):
        """[T, 80] -> untruncated (16T+1)*15 samples, no bias subtraction   (basis_melgan.py:196-208)"""
        c = self._prep_input(self._prep_inference_input(c))
        if self.use_cuda_graphs:
            return self._graphed_inference(c, return_weight=False, _inference=True).squeeze()
        return self.forward(c, return_weight=False, _inference=True).squeeze()

    def test(self, weight):
        """basis_signal(weight): Linear + overlap-add on a given weight tensor (basis_melgan.py:210-212)."""
        self._ensure_bound()
        w = weight.detach().to(self.device).float().contiguous()
        if w.dim() != 3 or w.shape[2] != self._cfg.out_channels:
            raise RuntimeError(f"expected weight [B, frames, {self._cfg.out_channels}], got {tuple(w.shape)}")
        B, n, Cw = w.shape
        hop = self.L // 2
        w = w.transpose(1, 2).contiguous()                 # the kernels read [B, C, frames]
        out = torch.empty(B, (n + 1) * hop, device=w.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().fv_basis_signal(self._handle, _lib.ptr(w), B, n, _lib.ptr(out),
                                                  1 if self.use_tensor_cores else 0, _lib.current_stream_ptr()),
                       "fv_basis_signal")
        return out


def build_generator(model_name: str, config: dict):
    """Construct a generator from a reference YAML dict exactly as bin/synthesize.py:25-68 does.  Superset: the MelGAN-family
    constructor kwargs the reference CLI never forwards (`bias`, `nonlinear_activation_params`,
    `use_final_nonlinear_activation`, melgan.py:20-36) are passed on when the dict carries them."""
    extra = {k: config[k] for k in ("bias", "nonlinear_activation_params", "use_final_nonlinear_activation") if k in config}
    if model_name == "melgan":
        return MelGANGenerator(in_channels=config["in_channels"], out_channels=config["out_channels"],
                               kernel_size=config["kernel_size"], channels=config["channels"],
                               upsample_scales=config["upsample_scales"],
                               stack_kernel_size=config["stack_kernel_size"], stacks=config["stacks"],
                               use_weight_norm=config["use_weight_norm"], use_causal_conv=config["use_causal_conv"],
                               **extra)
    if model_name == "hifigan":
        cls = HiFiGANGenerator
    elif model_name == "multiband-hifigan":
        cls = MultiBandHiFiGANGenerator
    elif model_name == "basis-melgan":
        basis_signal_weight = torch.zeros(config["L"], config["out_channels"]).float()
        return BasisMelGANGenerator(basis_signal_weight=basis_signal_weight, L=config["L"],
                                    in_channels=config["in_channels"], out_channels=config["out_channels"],
                                    kernel_size=config["kernel_size"], channels=config["channels"],
                                    upsample_scales=config["upsample_scales"],
                                    stack_kernel_size=config["stack_kernel_size"], stacks=config["stacks"],
                                    use_weight_norm=config["use_weight_norm"],
                                    use_causal_conv=config["use_causal_conv"],
                                    transposedconv=config["transposedconv"],
                                    lastlinear=config.get("lastlinear", False), **extra)
    else:
        raise Exception("no model find!")
    return cls(resblock_kernel_sizes=config["resblock_kernel_sizes"], upsample_rates=config["upsample_rates"],
               upsample_initial_channel=config["upsample_initial_channel"], resblock_type=config["resblock_type"],
               upsample_kernel_sizes=config["upsample_kernel_sizes"],
               resblock_dilation_sizes=config["resblock_dilation_sizes"], transposedconv=config["transposedconv"],
               bias=config["bias"])
