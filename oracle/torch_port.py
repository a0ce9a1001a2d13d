"""CPU port of the reference generators on the SAME ATen ops the reference calls.

TEST / BASELINE INFRASTRUCTURE ONLY (see np_oracle.py header for the rules).

The reference's CPU path *is* ``torch.nn.functional.conv1d`` & co. on the host
cores (synthesize.sh:9 forces CPU).  /root/reference does not exist on the GPU
box, so the reported CPU baseline (``bench.py`` ``cpu_baseline`` and
``--impl reference``) times this functional port: identical call sites
(cited per function), identical op order, running on ATen/oneDNN exactly like
the reference does.  It is pinned against the golden outputs of the real
reference (tests/test_oracle_golden.py) and is also used as the checker at
sizes where the numpy oracle would be slow.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # modules.py:9


def _p(params, name):
    return params.get(name)


def get_padding(kernel_size, dilation=1):  # modules.py:186
    return int((kernel_size * dilation - dilation) / 2)


def resblock1(x, params, prefix, k, dils):  # modules.py:190-230 (three pairs hard-coded from dilation[0..2])
    for i, d in enumerate((dils[0], dils[1], dils[2])):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, params[f"{prefix}.convs1.{i}.weight"], _p(params, f"{prefix}.convs1.{i}.bias"),
                      dilation=d, padding=get_padding(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, params[f"{prefix}.convs2.{i}.weight"], _p(params, f"{prefix}.convs2.{i}.bias"),
                      padding=get_padding(k, 1))
        x = xt + x
    return x


def resblock2(x, params, prefix, k, dils):  # modules.py:233-252 (two convs hard-coded from dilation[0..1])
    for i, d in enumerate((dils[0], dils[1])):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, params[f"{prefix}.convs.{i}.weight"], _p(params, f"{prefix}.convs.{i}.bias"),
                      dilation=d, padding=get_padding(k, d))
        x = xt + x
    return x


def upsample_layer(x, w, b, rate, padding):  # modules.py:160-177 (Stretch2d nearest on time, then Conv1d)
    x = F.interpolate(x.unsqueeze(1), scale_factor=(1, rate), mode="nearest").squeeze(1)
    return F.conv1d(x, w, b, padding=padding)


def mel_slope(cfg):  # nonlinear_activation_params default {"negative_slope": 0.2} (melgan.py:30); nn.LeakyReLU default 0.01
    return float(cfg.get("nonlinear_activation_params", {"negative_slope": 0.2}).get("negative_slope", 0.01))


def residual_stack(c, params, prefix, k, d, causal=False, slope=0.2):  # modules.py:343-382
    h = F.leaky_relu(c, slope)
    if causal:  # CausalConv1d, modules.py:273-297: ReflectionPad1d((k-1)*d) both sides, conv, keep the first T
        T = h.size(2)
        h = F.pad(h, ((k - 1) * d,) * 2, mode="reflect")
        h = F.conv1d(h, params[f"{prefix}.stack.1.conv.weight"], _p(params, f"{prefix}.stack.1.conv.bias"),
                     dilation=d)[:, :, :T]
        h = F.leaky_relu(h, slope)
        h = F.conv1d(h, params[f"{prefix}.stack.3.weight"], _p(params, f"{prefix}.stack.3.bias"))
    else:
        h = F.pad(h, ((k - 1) // 2 * d,) * 2, mode="reflect")
        h = F.conv1d(h, params[f"{prefix}.stack.2.weight"], _p(params, f"{prefix}.stack.2.bias"), dilation=d)
        h = F.leaky_relu(h, slope)
        h = F.conv1d(h, params[f"{prefix}.stack.4.weight"], _p(params, f"{prefix}.stack.4.bias"))
    return h + F.conv1d(c, params[f"{prefix}.skip_layer.weight"], _p(params, f"{prefix}.skip_layer.bias"))


def last_linear(x, params, prefix):  # modules.py:116-132, eval-mode BatchNorm1d
    def bn(x, q):
        return F.batch_norm(x, params[f"{q}.running_mean"], params[f"{q}.running_var"], params[f"{q}.weight"],
                            params[f"{q}.bias"], False, 0.1, 1e-5)
    x = bn(F.leaky_relu(x, 0.2), f"{prefix}.bn_1")
    x = F.conv1d(x, params[f"{prefix}.linear_1.weight"], _p(params, f"{prefix}.linear_1.bias"))
    x = bn(F.leaky_relu(x, 0.2), f"{prefix}.bn_2")
    return F.conv1d(x, params[f"{prefix}.linear_2.weight"], _p(params, f"{prefix}.linear_2.bias"))


def overlap_and_add(signal, frame_step):  # modules.py:34-73
    outer = signal.size()[:-2]
    frames, frame_length = signal.size()[-2:]
    sub = math.gcd(frame_length, frame_step)
    sub_step = frame_step // sub
    per_frame = frame_length // sub
    output_size = frame_step * (frames - 1) + frame_length
    out_sub = output_size // sub
    subframe_signal = signal.reshape(*outer, -1, sub)
    frame = torch.arange(0, out_sub).unfold(0, per_frame, sub_step).contiguous().view(-1)
    result = signal.new_zeros(*outer, out_sub, sub)
    result.index_add_(-2, frame, subframe_signal)
    return result.view(*outer, -1)


def _pqmf_filters(dtype):
    from .np_oracle import pqmf_filters
    ana, syn = pqmf_filters()
    return torch.from_numpy(ana).to(dtype), torch.from_numpy(syn).to(dtype)


def pqmf_synthesis(x, subbands=4, taps=62):  # pqmf.py:121-135
    _, syn = _pqmf_filters(x.dtype)
    updown = torch.zeros(subbands, subbands, subbands, dtype=x.dtype)
    for k in range(subbands):
        updown[k, k, 0] = 1.0
    x = F.conv_transpose1d(x, updown * subbands, stride=subbands)
    return F.conv1d(F.pad(x, (taps // 2, taps // 2)), syn)


def pqmf_analysis(x, subbands=4, taps=62):  # pqmf.py:108-119
    ana, _ = _pqmf_filters(x.dtype)
    updown = torch.zeros(subbands, subbands, subbands, dtype=x.dtype)
    for k in range(subbands):
        updown[k, k, 0] = 1.0
    x = F.conv1d(F.pad(x, (taps // 2, taps // 2)), ana)
    return F.conv1d(x, updown, stride=subbands)


def _hifigan_trunk(params, cfg, x):  # hifigan.py:92-106 / multiband_hifigan.py:101-116
    rks, rds = cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]
    nk = len(rks)
    rb = resblock1 if cfg.get("resblock_type", "1") == '1' else resblock2   # hifigan.py:28
    x = F.conv1d(x, params["conv_pre.weight"], _p(params, "conv_pre.bias"), padding=3)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        if cfg.get("transposedconv", True) == False:   # noqa: E712  hifigan.py:31-38
            x = upsample_layer(x, params[f"ups.{i}.conv.weight"], _p(params, f"ups.{i}.conv.bias"), u, k // 2)
        else:
            x = F.conv_transpose1d(x, params[f"ups.{i}.weight"], _p(params, f"ups.{i}.bias"), stride=u,
                                   padding=u // 2 + u % 2, output_padding=u % 2)
        xs = None
        for j in range(nk):
            r = rb(x, params, f"resblocks.{i * nk + j}", rks[j], rds[j])
            if xs is None:
                xs = r
            else:
                xs += r
        x = xs / nk
    x = F.leaky_relu(x)
    x = F.conv1d(x, params["conv_post.weight"], _p(params, "conv_post.bias"), padding=3)
    return torch.tanh(x)


def hifigan_forward(params, cfg, x):
    return _hifigan_trunk(params, cfg, x)[:, 0, :]


def mb_hifigan_forward(params, cfg, x):
    return _hifigan_trunk(params, cfg, x)


def _melgan_body(params, cfg, c):  # melgan.py:66-112, basis_melgan.py:70-125
    k = cfg["kernel_size"]
    x = F.pad(c, ((k - 1) // 2,) * 2, mode="reflect")
    x = F.conv1d(x, params["melgan.1.weight"], _p(params, "melgan.1.bias"))
    idx = 2
    slope = mel_slope(cfg)
    for u in cfg["upsample_scales"]:
        x = F.leaky_relu(x, slope)
        if cfg.get("transposedconv", True) == False and "L" in cfg:   # noqa: E712  basis_melgan.py:82-88 only
            x = upsample_layer(x, params[f"melgan.{idx + 1}.conv.weight"], _p(params, f"melgan.{idx + 1}.conv.bias"), u, u)
        else:
            x = F.conv_transpose1d(x, params[f"melgan.{idx + 1}.weight"], _p(params, f"melgan.{idx + 1}.bias"),
                                   stride=u, padding=u // 2 + u % 2, output_padding=u % 2)
        idx += 2
        for j in range(cfg["stacks"]):
            x = residual_stack(x, params, f"melgan.{idx}", cfg["stack_kernel_size"], cfg["stack_kernel_size"] ** j,
                               causal=cfg.get("use_causal_conv", False), slope=slope)
            idx += 1
    return x, idx


def melgan_forward(params, cfg, c):  # melgan.py:125-136
    x, idx = _melgan_body(params, cfg, c)
    k = cfg["kernel_size"]
    x = F.leaky_relu(x, mel_slope(cfg))
    x = F.pad(x, ((k - 1) // 2,) * 2, mode="reflect")
    x = F.conv1d(x, params[f"melgan.{idx}.conv.weight"], _p(params, f"melgan.{idx}.conv.bias"))
    if cfg.get("use_final_nonlinear_activation", True):   # melgan.py:108-110
        x = torch.tanh(x)
    return x[:, 0, :]


def _basis_pass(params, cfg, c):
    x, idx = _melgan_body(params, cfg, c)
    if cfg.get("lastlinear", False):  # basis_melgan.py:117-118
        x = last_linear(x, params, f"melgan.{idx}")
    if cfg.get("use_final_nonlinear_activation", True):   # basis_melgan.py:120-121
        x = torch.relu(x)
    weight = x.contiguous().transpose(1, 2)
    est = overlap_and_add(F.linear(weight, params["basis_signal.layer.weight"]), cfg["L"] // 2)
    return est, weight


def basis_melgan_forward(params, cfg, c):  # basis_melgan.py:140-162
    L = cfg["L"]
    zero_est, zero_weight = _basis_pass(params, cfg, torch.zeros_like(c))
    zero_est = zero_est[:, : zero_weight.size(1) * (L // 2)]
    est, weight = _basis_pass(params, cfg, c)
    est = est[:, : weight.size(1) * (L // 2)]
    return est - zero_est, weight - zero_weight


def basis_melgan_inference(params, cfg, c):  # basis_melgan.py:196-208
    est, _ = _basis_pass(params, cfg, c.transpose(1, 0).unsqueeze(0))
    return est.squeeze()


FORWARD = {
    "hifigan": hifigan_forward,
    "multiband-hifigan": mb_hifigan_forward,
    "melgan": melgan_forward,
    "basis-melgan": basis_melgan_forward,
}


def inference(model_name, params, cfg, c):
    """`.inference([T,80])` of the four generators."""
    if model_name == "basis-melgan":
        return basis_melgan_inference(params, cfg, c)
    y = FORWARD[model_name](params, cfg, c.transpose(1, 0).unsqueeze(0))
    if model_name == "multiband-hifigan":
        y = pqmf_synthesis(y)
    return y.squeeze()


def to_torch(params_np, dtype=torch.float32):
    import numpy as np  # noqa: F401
    return {k: torch.from_numpy(v).to(dtype) for k, v in params_np.items()}
