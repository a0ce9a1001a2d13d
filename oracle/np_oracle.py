"""CPU oracle for the FastVocoder generator forward path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``fastvocoder_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs use it, and only as the checker / the reported CPU baseline.

What it restates
----------------
The reference (xcmyz/FastVocoder, /root/reference) is pure PyTorch; the
arithmetic of the hot path lives in a third-party dependency that is not under
/root/reference: **PyTorch ATen** (``conv1d``, ``conv_transpose1d``,
``leaky_relu``, ``tanh``, ``ReflectionPad1d``, ``linear``, ``index_add_``), no
version pinned by the reference (no requirements file; this image carries torch
2.11.0).  This file restates the published semantics of those ops in numpy
(cross-correlation, zero / reflect padding, transposed-conv scatter form) and
then the reference's own call sites on top of them, each citing file:line.

Pinning
-------
The reference has no tests and ships no checkpoint, so parity is pinned by
running the reference classes themselves (``oracle/gen_golden.py`` imports
/root/reference in the build container) on seeded weights/inputs and committing
the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against every one of those fixtures.

All functions take/return numpy arrays laid out like the reference tensors:
activations ``[B, C, L]``, Conv1d weights ``[Cout, Cin, K]``, ConvTranspose1d
weights ``[Cin, Cout, K]``, Linear weights ``[out, in]``.  ``params`` is a
*folded* state_dict (after ``remove_weight_norm``): name -> ndarray.
"""
from __future__ import annotations

import math

import numpy as np

LRELU_SLOPE = 0.1  # model/generator/modules.py:9


# --------------------------------------------------------------------------
# ATen op restatements
# --------------------------------------------------------------------------
def leaky_relu(x, slope=0.01):
    """F.leaky_relu: x if x > 0 else slope * x (default slope 0.01)."""
    return np.where(x > 0, x, x * np.asarray(slope, dtype=x.dtype))


def conv1d(x, w, b=None, dilation=1, padding=0, stride=1):
    """torch.nn.functional.conv1d (cross-correlation, zero padding).

    y[b,o,t] = bias[o] + sum_{c,j} w[o,c,j] * xpad[b,c, t*stride + j*dilation]
    """
    B, Cin, L = x.shape
    Cout, Cin2, K = w.shape
    assert Cin == Cin2
    Lout = (L + 2 * padding - dilation * (K - 1) - 1) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (padding, padding)))
    y = np.zeros((B, Cout, Lout), dtype=x.dtype)
    for j in range(K):
        seg = xp[:, :, j * dilation: j * dilation + (Lout - 1) * stride + 1: stride]
        y += np.einsum("oc,bcl->bol", w[:, :, j], seg, optimize=True)
    if b is not None:
        y += b[None, :, None]
    return y


def conv_transpose1d(x, w, b=None, stride=1, padding=0, output_padding=0):
    """torch.nn.functional.conv_transpose1d, scatter form.

    full[b,o, i*stride + kk] += x[b,c,i] * w[c,o,kk];  y = full[padding : padding+Lout]
    Lout = (L-1)*stride - 2*padding + K + output_padding
    """
    B, Cin, L = x.shape
    Cin2, Cout, K = w.shape
    assert Cin == Cin2
    Lout = (L - 1) * stride - 2 * padding + K + output_padding
    full = np.zeros((B, Cout, (L - 1) * stride + K + output_padding + padding), dtype=x.dtype)
    for kk in range(K):
        full[:, :, kk: kk + (L - 1) * stride + 1: stride] += np.einsum(
            "co,bcl->bol", w[:, :, kk], x, optimize=True)
    y = full[:, :, padding: padding + Lout].copy()
    if b is not None:
        y += b[None, :, None]
    return y


def reflection_pad1d(x, p):
    """torch.nn.ReflectionPad1d: x[-i] = x[i], x[L-1+i] = x[L-1-i] (edge not repeated)."""
    assert p < x.shape[-1], "ReflectionPad1d needs pad < length"
    return np.pad(x, ((0, 0), (0, 0), (p, p)), mode="reflect")


# --------------------------------------------------------------------------
# model/generator/modules.py
# --------------------------------------------------------------------------
def get_padding(kernel_size, dilation=1):
    """modules.py:186"""
    return int((kernel_size * dilation - dilation) / 2)


def resblock1(x, params, prefix, kernel_size, dilations):
    """ResBlock1.forward, modules.py:223-230.  The constructor (modules.py:190-221) hard-codes THREE (convs1, convs2)
    pairs built from dilation[0..2], whatever the length of the list (shorter: IndexError, longer: ignored)."""
    for i, d in enumerate((dilations[0], dilations[1], dilations[2])):
        xt = leaky_relu(x, LRELU_SLOPE)
        xt = conv1d(xt, params[f"{prefix}.convs1.{i}.weight"], params[f"{prefix}.convs1.{i}.bias"],
                    dilation=d, padding=get_padding(kernel_size, d))
        xt = leaky_relu(xt, LRELU_SLOPE)
        xt = conv1d(xt, params[f"{prefix}.convs2.{i}.weight"], params[f"{prefix}.convs2.{i}.bias"],
                    dilation=1, padding=get_padding(kernel_size, 1))
        x = xt + x
    return x


def resblock2(x, params, prefix, kernel_size, dilations):
    """ResBlock2.forward, modules.py:247-252.  The constructor (modules.py:233-245) hard-codes TWO convs from
    dilation[0..1], whatever the length of the list."""
    for i, d in enumerate((dilations[0], dilations[1])):
        xt = leaky_relu(x, LRELU_SLOPE)
        xt = conv1d(xt, params[f"{prefix}.convs.{i}.weight"], params[f"{prefix}.convs.{i}.bias"],
                    dilation=d, padding=get_padding(kernel_size, d))
        x = xt + x
    return x


def upsample_layer(x, w, b, upsample_rate, padding):
    """UpsampleLayer.forward, modules.py:160-177: Stretch2d(nearest, x_scale = rate) on the time axis, then
    Conv1d(kernel k, stride 1, zero padding `padding`).  F.interpolate(mode="nearest") with an integer scale
    repeats every sample `rate` times: xu[t] = x[t // rate]."""
    return conv1d(np.repeat(x, upsample_rate, axis=-1), w, b, padding=padding)


def causal_conv1d(x, w, b, dilation, pad="reflect"):
    """CausalConv1d.forward, modules.py:273-297: pad((k-1)*d) on BOTH sides with the stack's pad class
    (ReflectionPad1d in every config), Conv1d(dilation d), then keep the first T outputs -> only the left
    padding is ever read."""
    K = w.shape[-1]
    p = (K - 1) * dilation
    xp = reflection_pad1d(x, p) if pad == "reflect" else np.pad(x, ((0, 0), (0, 0), (p, p)))
    return conv1d(xp, w, b, dilation=dilation)[:, :, : x.shape[-1]]


def residual_stack(c, params, prefix, kernel_size, dilation, slope=0.2, use_causal_conv=False):
    """ResidualStack.forward, modules.py:343-382.  Non-causal:
    stack = LReLU -> ReflectionPad1d((k-1)//2*d) -> Conv1d(k, dil d) -> LReLU -> Conv1d 1x1; + skip_layer(c).
    Causal (modules.py:355-361): stack = LReLU -> CausalConv1d -> LReLU -> Conv1d 1x1 (keys stack.1.conv / stack.3)."""
    h = leaky_relu(c, slope)
    if use_causal_conv:
        h = causal_conv1d(h, params[f"{prefix}.stack.1.conv.weight"], params.get(f"{prefix}.stack.1.conv.bias"), dilation)
        h = leaky_relu(h, slope)
        h = conv1d(h, params[f"{prefix}.stack.3.weight"], params.get(f"{prefix}.stack.3.bias"))
    else:
        h = reflection_pad1d(h, (kernel_size - 1) // 2 * dilation)
        h = conv1d(h, params[f"{prefix}.stack.2.weight"], params.get(f"{prefix}.stack.2.bias"), dilation=dilation)
        h = leaky_relu(h, slope)
        h = conv1d(h, params[f"{prefix}.stack.4.weight"], params.get(f"{prefix}.stack.4.bias"))
    s = conv1d(c, params[f"{prefix}.skip_layer.weight"], params.get(f"{prefix}.skip_layer.bias"))
    return h + s


def batch_norm1d_eval(x, params, prefix, eps=1e-5):
    """nn.BatchNorm1d in eval mode: (x - running_mean) / sqrt(running_var + eps) * weight + bias per channel."""
    inv = 1.0 / np.sqrt(params[f"{prefix}.running_var"].astype(x.dtype) + np.asarray(eps, dtype=x.dtype))
    return ((x - params[f"{prefix}.running_mean"].astype(x.dtype)[None, :, None]) * inv[None, :, None]
            * params[f"{prefix}.weight"].astype(x.dtype)[None, :, None]
            + params[f"{prefix}.bias"].astype(x.dtype)[None, :, None])


def last_linear(x, params, prefix, slope=0.2):
    """LastLinear.forward (eval mode), modules.py:116-132: LReLU -> BN -> 1x1 -> LReLU -> BN -> 1x1."""
    x = leaky_relu(x, slope)
    x = batch_norm1d_eval(x, params, f"{prefix}.bn_1")
    x = conv1d(x, params[f"{prefix}.linear_1.weight"], params.get(f"{prefix}.linear_1.bias"))
    x = leaky_relu(x, slope)
    x = batch_norm1d_eval(x, params, f"{prefix}.bn_2")
    return conv1d(x, params[f"{prefix}.linear_2.weight"], params.get(f"{prefix}.linear_2.bias"))


def last_layer(x, params, prefix, kernel_size, slope=0.2):
    """LastLayer.forward, modules.py:85-89: LReLU -> ReflectionPad1d((k-1)//2) -> Conv1d."""
    x = leaky_relu(x, slope)
    x = reflection_pad1d(x, (kernel_size - 1) // 2)
    return conv1d(x, params[f"{prefix}.conv.weight"], params.get(f"{prefix}.conv.bias"))


def overlap_and_add(signal, frame_step):
    """overlap_and_add, modules.py:34-73 (gcd sub-frames + index_add_).

    signal [..., frames, frame_length] -> [..., (frames-1)*frame_step + frame_length]
    """
    outer = signal.shape[:-2]
    frames, frame_length = signal.shape[-2:]
    subframe_length = math.gcd(frame_length, frame_step)
    subframe_step = frame_step // subframe_length
    subframes_per_frame = frame_length // subframe_length
    output_size = frame_step * (frames - 1) + frame_length
    output_subframes = output_size // subframe_length
    subframe_signal = signal.reshape(*outer, -1, subframe_length)
    # torch.arange(0, output_subframes).unfold(0, subframes_per_frame, subframe_step)
    nwin = (output_subframes - subframes_per_frame) // subframe_step + 1
    frame = (np.arange(nwin)[:, None] * subframe_step + np.arange(subframes_per_frame)[None, :]).reshape(-1)
    result = np.zeros((*outer, output_subframes, subframe_length), dtype=signal.dtype)
    flat_r = result.reshape(-1, output_subframes, subframe_length)
    flat_s = subframe_signal.reshape(-1, subframe_signal.shape[-2], subframe_length)
    for b in range(flat_r.shape[0]):  # index_add_ along dim -2, sequential order
        np.add.at(flat_r[b], frame, flat_s[b])
    return result.reshape(*outer, -1)


def basis_signal_layer(weight, basis_w, L):
    """BasisSignalLayer.forward, modules.py:264-267: Linear(no bias) then OLA(L//2)."""
    source = weight @ basis_w.T
    return overlap_and_add(source, L // 2)


# --------------------------------------------------------------------------
# model/generator/pqmf.py
# --------------------------------------------------------------------------
def design_prototype_filter(taps=62, cutoff_ratio=0.142, beta=9.0):
    """pqmf.py:15-48 (float64)."""
    from scipy.signal.windows import kaiser  # the reference's `scipy.signal.kaiser` is this function
    assert taps % 2 == 0
    assert 0.0 < cutoff_ratio < 1.0
    omega_c = np.pi * cutoff_ratio
    with np.errstate(invalid="ignore"):
        h_i = np.sin(omega_c * (np.arange(taps + 1) - 0.5 * taps)) \
            / (np.pi * (np.arange(taps + 1) - 0.5 * taps))
    h_i[taps // 2] = np.cos(0) * cutoff_ratio
    return h_i * kaiser(taps + 1, beta)


def pqmf_filters(subbands=4, taps=62, cutoff_ratio=0.142, beta=9.0):
    """PQMF.__init__, pqmf.py:61-106. Returns (analysis [S,1,taps+1], synthesis [1,S,taps+1]) float32."""
    h_proto = design_prototype_filter(taps, cutoff_ratio, beta)
    h_analysis = np.zeros((subbands, len(h_proto)))
    h_synthesis = np.zeros((subbands, len(h_proto)))
    for k in range(subbands):
        h_analysis[k] = 2 * h_proto * np.cos(
            (2 * k + 1) * (np.pi / (2 * subbands)) * (np.arange(taps + 1) - (taps / 2))
            + (-1) ** k * np.pi / 4)
        h_synthesis[k] = 2 * h_proto * np.cos(
            (2 * k + 1) * (np.pi / (2 * subbands)) * (np.arange(taps + 1) - (taps / 2))
            - (-1) ** k * np.pi / 4)
    return (h_analysis.astype(np.float32)[:, None, :], h_synthesis.astype(np.float32)[None, :, :])


def pqmf_analysis(x, subbands=4, taps=62):
    """PQMF.analysis, pqmf.py:108-119: conv1d(pad(x), analysis) then stride-S pick via updown filter."""
    ana, _ = pqmf_filters(subbands, taps)
    ana = ana.astype(x.dtype)
    y = conv1d(x, ana, padding=taps // 2)
    updown = np.zeros((subbands, subbands, subbands), dtype=x.dtype)
    for k in range(subbands):
        updown[k, k, 0] = 1.0
    return conv1d(y, updown, stride=subbands)


def pqmf_synthesis(x, subbands=4, taps=62):
    """PQMF.synthesis, pqmf.py:121-135: conv_transpose1d(x, updown*S, stride=S) then conv1d(pad(.), synthesis)."""
    _, syn = pqmf_filters(subbands, taps)
    syn = syn.astype(x.dtype)
    updown = np.zeros((subbands, subbands, subbands), dtype=x.dtype)
    for k in range(subbands):
        updown[k, k, 0] = 1.0
    up = conv_transpose1d(x, updown * subbands, stride=subbands)
    return conv1d(up, syn, padding=taps // 2)


# --------------------------------------------------------------------------
# generators
# --------------------------------------------------------------------------
def _hifigan_trunk(params, cfg, x):
    """Shared trunk of hifigan.py:92-106 and multiband_hifigan.py:101-116."""
    rates = cfg["upsample_rates"]
    ksz = cfg["upsample_kernel_sizes"]
    rks = cfg["resblock_kernel_sizes"]
    rds = cfg["resblock_dilation_sizes"]
    num_kernels = len(rks)
    rb = resblock1 if cfg.get("resblock_type", "1") == '1' else resblock2   # hifigan.py:28, the reference's comparison
    x = conv1d(x, params["conv_pre.weight"], params.get("conv_pre.bias"), padding=3)
    for i, (u, k) in enumerate(zip(rates, ksz)):
        x = leaky_relu(x, LRELU_SLOPE)
        if cfg.get("transposedconv", True) == False:   # noqa: E712  hifigan.py:31-38
            x = upsample_layer(x, params[f"ups.{i}.conv.weight"], params.get(f"ups.{i}.conv.bias"), u, k // 2)
        else:
            x = conv_transpose1d(x, params[f"ups.{i}.weight"], params.get(f"ups.{i}.bias"),
                                 stride=u, padding=u // 2 + u % 2, output_padding=u % 2)
        xs = None
        for j in range(num_kernels):
            r = rb(x, params, f"resblocks.{i * num_kernels + j}", rks[j], rds[j])
            xs = r if xs is None else xs + r
        x = xs / np.asarray(num_kernels, dtype=x.dtype)
    x = leaky_relu(x)  # default slope 0.01 (hifigan.py:104)
    x = conv1d(x, params["conv_post.weight"], params.get("conv_post.bias"), padding=3)
    return np.tanh(x)


def hifigan_forward(params, cfg, x):
    """HiFiGANGenerator.forward, hifigan.py:92-108: [B,80,T] -> [B, prod(rates)*T]."""
    return _hifigan_trunk(params, cfg, x)[:, 0, :]


def hifigan_inference(params, cfg, c):
    """HiFiGANGenerator.inference, hifigan.py:110-129: [T,80] -> squeeze."""
    return np.squeeze(_hifigan_trunk(params, cfg, c.T[None]))


def mb_hifigan_forward(params, cfg, x):
    """MultiBandHiFiGANGenerator.forward, multiband_hifigan.py:101-116: [B,80,T] -> [B,4,L] (no PQMF)."""
    return _hifigan_trunk(params, cfg, x)


def mb_hifigan_inference(params, cfg, c):
    """MultiBandHiFiGANGenerator.inference, multiband_hifigan.py:118-137 (applies PQMF synthesis)."""
    return np.squeeze(pqmf_synthesis(_hifigan_trunk(params, cfg, c.T[None])))


def mel_slope(cfg):
    """LeakyReLU slope of the MelGAN family: nonlinear_activation_params (default {"negative_slope": 0.2}, melgan.py:30;
    a dict without the key falls back to nn.LeakyReLU's own default 0.01)."""
    return float(cfg.get("nonlinear_activation_params", {"negative_slope": 0.2}).get("negative_slope", 0.01))


def _melgan_body(params, cfg, c, n_prefix="melgan"):
    """The nn.Sequential built in melgan.py:66-112 / basis_melgan.py:70-125 up to (excluding) the final layer."""
    ch = cfg["channels"]
    scales = cfg["upsample_scales"]
    ksize = cfg["kernel_size"]
    stacks = cfg["stacks"]
    sk = cfg["stack_kernel_size"]
    idx = 0
    slope = mel_slope(cfg)
    x = reflection_pad1d(c, (ksize - 1) // 2)                                   # melgan.0
    x = conv1d(x, params[f"{n_prefix}.1.weight"], params.get(f"{n_prefix}.1.bias"))  # melgan.1
    idx = 2
    for i, u in enumerate(scales):
        x = leaky_relu(x, slope)                                                # idx
        if cfg.get("transposedconv", True) == False and "L" in cfg:   # noqa: E712  basis_melgan.py:82-88 only
            x = upsample_layer(x, params[f"{n_prefix}.{idx + 1}.conv.weight"],
                               params.get(f"{n_prefix}.{idx + 1}.conv.bias"), u, u)
        else:
            x = conv_transpose1d(x, params[f"{n_prefix}.{idx + 1}.weight"], params.get(f"{n_prefix}.{idx + 1}.bias"),
                                 stride=u, padding=u // 2 + u % 2, output_padding=u % 2)
        idx += 2
        for j in range(stacks):
            x = residual_stack(x, params, f"{n_prefix}.{idx}", sk, sk ** j, slope=slope,
                               use_causal_conv=cfg.get("use_causal_conv", False))
            idx += 1
    return x, idx


def melgan_forward(params, cfg, c):
    """MelGANGenerator.forward, melgan.py:125-136: [B,80,T] -> [B, prod(scales)*T]."""
    x, idx = _melgan_body(params, cfg, c)
    x = last_layer(x, params, f"melgan.{idx}", cfg["kernel_size"], slope=mel_slope(cfg))
    if cfg.get("use_final_nonlinear_activation", True):
        x = np.tanh(x)
    return x[:, 0, :]


def melgan_inference(params, cfg, c):
    """MelGANGenerator.inference, melgan.py:172-185."""
    return np.squeeze(melgan_forward(params, cfg, c.T[None]))


def _basis_pass(params, cfg, c):
    x, idx = _melgan_body(params, cfg, c)
    if cfg.get("lastlinear", False):                                            # basis_melgan.py:117-118
        x = last_linear(x, params, f"melgan.{idx}")
    if cfg.get("use_final_nonlinear_activation", True):
        x = np.maximum(x, 0)                                                    # ReLU, basis_melgan.py:121
    weight = np.ascontiguousarray(x.transpose(0, 2, 1))
    est = basis_signal_layer(weight, params["basis_signal.layer.weight"], cfg["L"])
    return est, weight


def basis_melgan_forward(params, cfg, c):
    """BasisMelGANGenerator.forward, basis_melgan.py:140-162: returns (est - zero_est, weight - zero_weight)."""
    L = cfg["L"]
    zero_est, zero_weight = _basis_pass(params, cfg, np.zeros_like(c))
    zero_est = zero_est[:, : zero_weight.shape[1] * (L // 2)]
    est, weight = _basis_pass(params, cfg, c)
    est = est[:, : weight.shape[1] * (L // 2)]
    return est - zero_est, weight - zero_weight


def basis_melgan_inference(params, cfg, c):
    """BasisMelGANGenerator.inference, basis_melgan.py:196-208: one pass, untruncated."""
    est, _ = _basis_pass(params, cfg, c.T[None])
    return np.squeeze(est)


FORWARD = {
    "hifigan": hifigan_forward,
    "multiband-hifigan": mb_hifigan_forward,
    "melgan": melgan_forward,
    "basis-melgan": basis_melgan_forward,
}
INFERENCE = {
    "hifigan": hifigan_inference,
    "multiband-hifigan": mb_hifigan_inference,
    "melgan": melgan_inference,
    "basis-melgan": basis_melgan_inference,
}


def encode_16bits(x, rescale_out=1.0):
    """data/audio.py:12-14 (save_wav's quantiser)."""
    x = x * (32767 / max(0.01, np.max(np.abs(x))) * rescale_out)
    return x.astype(np.int16)
