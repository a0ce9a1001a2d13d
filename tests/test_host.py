"""Host-side logic (no GPU): C-ABI surface, parameter tables, weight-norm folding, config errors."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest
import torch

from conftest import ALL_KEYS, GOLDEN, MODEL_KEYS, REPO, folded_weights
from fastvocoder_b200 import _lib, build_generator
from fastvocoder_b200.pqmf import PQMF, design_filters


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "fastvocoder_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fv_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} not exported by {_lib.LIB_PATH}"
    assert _lib.lib().fv_abi_version() == _lib.FV_ABI_VERSION


@pytest.mark.parametrize("key", ALL_KEYS)
def test_state_dict_keys_match_reference(specs, key):
    m = build_generator(specs[key]["model_name"], specs[key]["config"])
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(s) for k, s in specs[key]["spec_wn"]}           # checkpoint (weight-norm) form
    m.eval()
    m.remove_weight_norm()
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(s) for k, s in specs[key]["spec_folded"]}       # after remove_weight_norm()


@pytest.mark.parametrize("key,name", [("hifigan-light", "hifigan"), ("melgan-original", "melgan")])
def test_weight_norm_fold_bit_exact(specs, key, name):
    g = dict(np.load(os.path.join(GOLDEN, f"fold_{key}.npz")))
    m = build_generator(name, specs[key]["config"])
    sd = m.state_dict()
    layers = sorted({k[:-len(".weight_g")] for k in g if k.endswith(".weight_g")})
    for p in layers:
        sd[p + ".weight_g"] = torch.from_numpy(g[p + ".weight_g"])
        sd[p + ".weight_v"] = torch.from_numpy(g[p + ".weight_v"])
    m.load_state_dict(sd)
    m.remove_weight_norm()
    out = m.state_dict()
    for p in layers:
        assert np.array_equal(out[p + ".weight"].numpy(), g[p + ".weight"]), p   # torch's own fold -> same bits


def test_state_dict_round_trip_both_forms(specs):
    key = "basis-melgan-light"
    m = build_generator("basis-melgan", specs[key]["config"])
    sd_wn = m.state_dict()
    m2 = build_generator("basis-melgan", specs[key]["config"])
    m2.load_state_dict(sd_wn)
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd_wn[k]), k
    m.remove_weight_norm()
    m2.remove_weight_norm()
    a, b = m.state_dict(), m2.state_dict()
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # folded dict loads too, and strictness is enforced
    m3 = build_generator("basis-melgan", specs[key]["config"])
    m3.load_state_dict(a)
    with pytest.raises(RuntimeError):
        bad = dict(a)
        bad.pop("melgan.1.weight")
        m3.load_state_dict(bad)
    with pytest.raises(RuntimeError):
        bad = dict(a)
        bad["melgan.1.weight"] = torch.zeros(3, 3, 3)
        m3.load_state_dict(bad)


def test_synthetic_weights_load(specs):
    key = "multiband-hifigan-light"
    m = build_generator("multiband-hifigan", specs[key]["config"])
    w = folded_weights(specs, key)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    sd = m.state_dict()
    for k, v in w.items():
        assert np.array_equal(sd[k].numpy(), v)
    assert "pqmf.synthesis_filter" in sd and "weight_g" not in "".join(sd.keys())


def test_error_conventions(specs):
    with pytest.raises(Exception, match="no model find!"):
        build_generator("wavernn", {})
    cfg = dict(specs["melgan-original"]["config"])
    cfg["kernel_size"] = 6
    with pytest.raises(AssertionError, match="Not support even number kernel size."):
        build_generator("melgan", cfg)
    m = build_generator("hifigan", specs["hifigan-light"]["config"])
    with pytest.raises(_lib.FvError, match="no CPU"):
        m(torch.zeros(1, 80, 8))                      # no CPU fallback: must fail loudly
    with pytest.raises(_lib.FvError, match="no CPU"):
        m.tensor_cores_usable                         # needs bound (device-resident) weights
    from fastvocoder_b200 import HostPipeline
    with pytest.raises(_lib.FvError, match="CUDA"):
        HostPipeline(m)                               # the serving pipeline has no CPU path either
    with pytest.raises(ValueError):
        HostPipeline(m, depth=1)
    bad = _lib.FvConfig()
    bad.kind = 17
    h = C.c_void_p()
    assert _lib.lib().fv_create(C.byref(bad), C.byref(h)) == -1
    assert b"no model find" in _lib.lib().fv_last_error()


def test_output_lengths(specs):
    want = specs["_demo_lengths_T585"]
    m = build_generator("hifigan", specs["hifigan-light"]["config"])
    assert m.out_length(585) == want["hifigan-light"]
    m = build_generator("multiband-hifigan", specs["multiband-hifigan-light"]["config"])
    assert 4 * m.out_length(585) == want["multiband-hifigan-light"]
    m = build_generator("multiband-hifigan", specs["multiband-hifigan-large"]["config"])
    assert 4 * m.out_length(585) == want["multiband-hifigan-large"]            # the 60T-20 quirk
    m = build_generator("basis-melgan", specs["basis-melgan-light"]["config"])
    assert m.out_length(585, _lib.FV_FWD_BASIS_INFERENCE) == want["basis-melgan-light"]
    assert m.out_length(585) == 240 * 585
    m = build_generator("melgan", specs["melgan-original"]["config"])
    assert m.out_length(200) == 48000


def test_flops_accounting_matches_survey(specs):
    # MAC per mel frame
    for key, name, macs in [("hifigan-light", "hifigan", 62.47e6), ("multiband-hifigan-light", "multiband-hifigan", 53.50e6),
                            ("melgan-original", "melgan", 45.48e6), ("hifigan-large", "hifigan", 249.5e6)]:
        m = build_generator(name, specs[key]["config"])
        per_frame = m.forward_flops(1, 1000) / 2 / 1000
        assert abs(per_frame - macs) / macs < 2e-3, (key, per_frame)
    m = build_generator("basis-melgan", specs["basis-melgan-light"]["config"])
    one_pass = m.forward_flops(1, 1000, _lib.FV_FWD_BASIS_INFERENCE) / 2 / 1000
    # SURVEY lists 22.44 M with the basis Linear as 0.014 M; it is 256*30*16 = 0.123 M/frame -> 22.55 M
    assert abs(one_pass - 22.548e6) / 22.548e6 < 1e-3


def test_pqmf_design_is_byte_identical(specs, ops_golden):
    ana, syn = design_filters()
    assert hashlib.sha256(ana.numpy().tobytes()).hexdigest() == specs["_pqmf"]["analysis_sha256"]
    assert hashlib.sha256(syn.numpy().tobytes()).hexdigest() == specs["_pqmf"]["synthesis_sha256"]
    p = PQMF()
    assert set(p.state_dict()) == {"analysis_filter", "synthesis_filter", "updown_filter"}
    assert np.array_equal(p.synthesis_filter.numpy(), ops_golden["pqmf_synthesis_filter"])
    with pytest.raises(_lib.FvError):
        p.synthesis(torch.zeros(1, 4, 8))             # CPU tensor: no fallback


def test_tile_planners_respect_hardware_limits(tmp_path):
    """Host-only sweep (tests/native/planner_check.cu): every plan tc2_plan / tc3_plan makes for the shipped layer shapes and a
    grid of odd ones fits shared memory / TMEM / barrier slots and is internally consistent — catches planner regressions
    before they reach a GPU box (a bad plan there means a trap or a hang, not a test failure)."""
    import shutil
    import subprocess
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc) and not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "planner_check")
    src = os.path.join(REPO, "tests", "native", "planner_check.cu")
    r = subprocess.run([nvcc if os.path.exists(nvcc) else "nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-diag-suppress", "177", "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    env = {k: v for k, v in os.environ.items() if not k.startswith("FV_")}
    variants = [{}, {"FV_TC3_PP": "0"}, {"FV_TC3_PP": "1"}, {"FV_TC3_PP": "3"}, {"FV_TC3_RING": "0"},
                {"FV_LOADER_FLAT": "2"}, {"FV_A_REUSE": "1"}, {"FV_TC3_RING_M": "1"}, {"FV_STACK_M": "1"},
                {"FV_STACK_M": "2"}]                                                      # every planner knob
    for extra in variants:
        r = subprocess.run([exe], capture_output=True, text=True, env=dict(env, **extra))
        assert r.returncode == 0, (extra, r.stdout[-3000:])
        assert "failures: 0" in r.stdout, extra


def test_switch_error_conventions(specs):
    """Non-default switches: accepted where the reference class accepts them, rejected (loudly) where it does not."""
    cfg = dict(specs["melgan-original"]["config"])
    cfg["use_causal_conv"] = True
    build_generator("melgan", cfg)                                   # causal MelGAN with an odd kernel is wired
    cfg["kernel_size"] = 6
    with pytest.raises(NotImplementedError):                         # melgan.py:68-71 changes the sequence length
        build_generator("melgan", cfg)
    h = dict(specs["hifigan-light"]["config"])
    h["transposedconv"] = False
    m = build_generator("hifigan", h)                                # UpsampleLayer: k = upsample_kernel_sizes[i], padding k // 2
    assert m.out_length(24) > 240 * 24                                # u*L + 1 per stage (even kernels)
    bad = _lib.FvConfig()
    bad.kind = _lib.FV_HIFIGAN
    bad.in_channels = 80
    bad.num_upsamples = 1
    bad.upsample_rates[0] = 2
    bad.upsample_kernel_sizes[0] = 4
    bad.channels[0] = 32
    bad.channels[1] = 16
    bad.pre_kernel_size = 7
    bad.use_causal_conv = 1                                          # a MelGAN-family switch on a HiFi-GAN config
    hd = C.c_void_p()
    assert _lib.lib().fv_create(C.byref(bad), C.byref(hd)) != 0
    assert b"use_causal_conv" in _lib.lib().fv_last_error()
    b = dict(specs["basis-melgan-light"]["config"])
    b["lastlinear"] = True
    b["out_channels"] = 64
    mb = build_generator("basis-melgan", b)
    keys = set(mb.state_dict())
    assert any(k.endswith("bn_1.running_var") for k in keys) and any(k.endswith("linear_2.weight_g") for k in keys)


def test_variant_output_lengths(specs):
    """UpsampleLayer changes the length law: Conv1d(k, padding k // 2) on the stretched signal gives u*L + 1 for even k."""
    g = np.load(os.path.join(GOLDEN, "model_hifigan-light-upsamplelayer.npz"))
    h = dict(specs["hifigan-light-upsamplelayer"]["config"])
    m = build_generator("hifigan", h)
    assert m.out_length(g["mel"].shape[-1]) == g["forward0_f32"].shape[-1]
    g = np.load(os.path.join(GOLDEN, "model_multiband-hifigan-light-upsamplelayer.npz"))
    m = build_generator("multiband-hifigan", dict(specs["multiband-hifigan-light-upsamplelayer"]["config"]))
    assert m.out_length(g["mel"].shape[-1]) == g["forward0_f32"].shape[-1]


# ---- round 2: advisor findings -----------------------------------------------------------------------------------
def test_resblock_conv_counts_and_type_comparison_follow_the_reference_constructors():
    """modules.py:190-251 hard-codes 3 (convs1, convs2) pairs for ResBlock1 and 2 convs for ResBlock2 whatever the length of
    the dilation list; hifigan.py:28 compares `resblock_type == '1'` (an int 1 selects ResBlock2)."""
    from fastvocoder_b200 import HiFiGANGenerator
    m2 = HiFiGANGenerator(resblock_type="2")                       # constructor-default dilations [[1, 3, 5]] * 3
    keys = {n for n, _, _ in m2._spec}
    assert "resblocks.0.convs.1.weight" in keys and "resblocks.0.convs.2.weight" not in keys
    assert not any(".convs1." in k for k in keys)
    m_int = HiFiGANGenerator(resblock_type=1)                      # unquoted YAML 1 -> ResBlock2, as in the reference
    assert any(".convs." in n for n, _, _ in m_int._spec) and not any(".convs1." in n for n, _, _ in m_int._spec)
    m1 = HiFiGANGenerator(resblock_dilation_sizes=[[1, 3, 5, 7]] * 3)   # the 4th entry is ignored
    assert "resblocks.0.convs1.2.weight" in {n for n, _, _ in m1._spec}
    assert "resblocks.0.convs1.3.weight" not in {n for n, _, _ in m1._spec}
    with pytest.raises(IndexError):
        HiFiGANGenerator(resblock_dilation_sizes=[[1, 3]] * 3)     # dilation[2] raises in ResBlock1.__init__


def test_nested_generator_state_dict_round_trips_and_dtype_casts_keep_float32():
    import torch
    from fastvocoder_b200 import HiFiGANGenerator

    class Wrapper(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gen = HiFiGANGenerator(upsample_initial_channel=32)

    w = Wrapper()
    sd = w.state_dict()
    assert "gen.conv_pre.weight_g" in sd and "gen.ups.0.weight_v" in sd
    w2 = Wrapper()
    res = w2.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    for (k, a), (_, b) in zip(w.gen.state_dict().items(), w2.gen.state_dict().items()):
        assert torch.equal(a, b), k
    w2.half()
    assert w2.gen.packed_weights.dtype == torch.float32
    assert "pre.conv_pre.weight_g" in w.gen.state_dict(prefix="pre.")
