"""Pin the CPU oracle (numpy restatement + ATen port) against outputs of the real reference.

The fixtures under tests/golden/ were produced by oracle/gen_golden.py, which imports the unmodified
reference from /root/reference (build container only).  No GPU needed.
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import ALL_KEYS, folded_weights, load_model_golden
from oracle import np_oracle as O
from oracle import torch_port as P

TOL32 = 2e-5   # fp32 summation-order noise through ~80 conv layers (measured fp32-vs-fp64: <= 4e-6)


@pytest.mark.parametrize("key", ALL_KEYS)
def test_numpy_oracle_forward_matches_reference(specs, key):
    g = load_model_golden(key)
    name, cfg = specs[key]["model_name"], specs[key]["config"]
    w = {k: v.astype(np.float64) for k, v in folded_weights(specs, key).items()}
    y = O.FORWARD[name](w, cfg, g["mel"].astype(np.float64))
    ys = y if isinstance(y, tuple) else (y,)
    assert ys[0].shape == g["forward0_f64"].shape
    # fp64 oracle vs fp64 reference: only association-order noise
    assert np.abs(ys[0] - g["forward0_f64"]).max() < 1e-10
    assert np.abs(ys[0] - g["forward0_f32"]).max() < TOL32
    if len(ys) > 1:
        assert np.abs(ys[1] - g["forward1_f32"]).max() < 5e-5


@pytest.mark.parametrize("key", ALL_KEYS)
def test_numpy_oracle_inference_matches_reference(specs, key):
    g = load_model_golden(key)
    name, cfg = specs[key]["model_name"], specs[key]["config"]
    w = {k: v.astype(np.float64) for k, v in folded_weights(specs, key).items()}
    y = O.INFERENCE[name](w, cfg, g["mel"][0].T.astype(np.float64))
    assert y.shape == g["inference_f64"].shape
    assert np.abs(y - g["inference_f64"]).max() < 1e-10
    if "realmel_T80" in g:
        y = O.INFERENCE[name](w, cfg, g["realmel_T80"].astype(np.float64))
        assert np.abs(y - g["realmel_inference_f64"]).max() < 1e-10
        assert np.abs(y - g["realmel_inference_f32"]).max() < TOL32


@pytest.mark.parametrize("key", ALL_KEYS)
def test_torch_port_matches_reference(specs, key):
    g = load_model_golden(key)
    name, cfg = specs[key]["model_name"], specs[key]["config"]
    w = P.to_torch(folded_weights(specs, key))
    with torch.no_grad():
        y = P.FORWARD[name](w, cfg, torch.from_numpy(g["mel"]))
        ys = y if isinstance(y, tuple) else (y,)
        # same ATen ops, same order -> bit-identical to the reference on the same machine; allow noise across CPUs
        assert np.abs(ys[0].numpy() - g["forward0_f32"]).max() < TOL32
        if len(ys) > 1:
            assert np.abs(ys[1].numpy() - g["forward1_f32"]).max() < 5e-5
        inf = P.inference(name, w, cfg, torch.from_numpy(g["mel"][0].T.copy()))
        assert np.abs(inf.numpy() - g["inference_f32"]).max() < TOL32


def test_output_lengths_match_demo_wavs(specs):
    """resource/demo/*.wav pin lengths for T=585."""
    T = 585
    want = specs["_demo_lengths_T585"]
    assert 240 * T == want["hifigan-light"] == want["multiband-hifigan-light"]
    assert (16 * T + 1) * 15 == want["basis-melgan-light"]
    assert 4 * (60 * T - 20) == want["multiband-hifigan-large"]
    # and the oracle's transposed conv reproduces the MB-large 10T-4 / 60T-20 quirk
    x = np.zeros((1, 2, 7))
    y = O.conv_transpose1d(x, np.zeros((2, 3, 16)), None, stride=10, padding=5, output_padding=0)
    assert y.shape[-1] == 10 * 7 - 4


def test_ops_convt(ops_golden):
    for (k, s) in [(16, 8), (10, 5), (6, 3), (4, 2), (20, 10), (12, 6), (8, 4), (16, 10), (16, 6)]:
        pre = f"convt_k{k}_s{s}_"
        y = O.conv_transpose1d(ops_golden[pre + "x"].astype(np.float64), ops_golden[pre + "w"].astype(np.float64),
                               ops_golden[pre + "b"].astype(np.float64), stride=s, padding=s // 2 + s % 2,
                               output_padding=s % 2)
        assert np.abs(y - ops_golden[pre + "y64"]).max() < 1e-12


def test_ops_resblock1_and_residual_stack(ops_golden):
    for k in (3, 7, 11):
        params = {"rb." + n[len(f"resblock1_k{k}_p_"):]: v.astype(np.float64)
                  for n, v in ops_golden.items() if n.startswith(f"resblock1_k{k}_p_")}
        y = O.resblock1(ops_golden[f"resblock1_k{k}_x"].astype(np.float64), params, "rb", k, (1, 3, 5))
        assert np.abs(y - ops_golden[f"resblock1_k{k}_y64"]).max() < 1e-12
    for d in (1, 3, 9):
        params = {"rs." + n[len(f"resstack_d{d}_p_"):]: v.astype(np.float64)
                  for n, v in ops_golden.items() if n.startswith(f"resstack_d{d}_p_")}
        y = O.residual_stack(ops_golden[f"resstack_d{d}_x"].astype(np.float64), params, "rs", 3, d)
        assert np.abs(y - ops_golden[f"resstack_d{d}_y64"]).max() < 1e-12
    params = {"ll." + n[len("lastlayer_p_"):]: v for n, v in ops_golden.items() if n.startswith("lastlayer_p_")}
    y = O.last_layer(ops_golden["lastlayer_x"], params, "ll", 7)
    assert np.abs(y - ops_golden["lastlayer_y"]).max() < 1e-5


def test_ops_overlap_add_bit_exact(ops_golden):
    """Index arithmetic: the two-addend sums are order independent -> bit-exact."""
    assert np.array_equal(O.overlap_and_add(ops_golden["ola_signal"], 15), ops_golden["ola_out_step15"])
    # gcd sub-frame path with 3 overlapping frames: summation order = index_add_ order
    assert np.abs(O.overlap_and_add(ops_golden["ola2_signal"], 4) - ops_golden["ola2_out_step4"]).max() < 1e-6
    y = O.basis_signal_layer(ops_golden["basis_in"], ops_golden["basis_w"], 30)
    assert np.abs(y - ops_golden["basis_out"]).max() < 1e-5


def test_pqmf_filters_bit_exact(specs, ops_golden):
    ana, syn = O.pqmf_filters()
    assert np.array_equal(ana, ops_golden["pqmf_analysis_filter"])
    assert np.array_equal(syn, ops_golden["pqmf_synthesis_filter"])
    assert hashlib.sha256(ana.tobytes()).hexdigest() == specs["_pqmf"]["analysis_sha256"]
    assert hashlib.sha256(syn.tobytes()).hexdigest() == specs["_pqmf"]["synthesis_sha256"]
    assert specs["_pqmf"]["synthesis_sha256"].startswith("012720fb7e469a12")
    assert specs["_pqmf"]["analysis_sha256"].startswith("4af45e104f11e053")


def test_pqmf_analysis_synthesis(ops_golden):
    assert np.abs(O.pqmf_analysis(ops_golden["pqmf_ana_x"]) - ops_golden["pqmf_ana_y"]).max() < 2e-6
    assert np.abs(O.pqmf_synthesis(ops_golden["pqmf_syn_x"]) - ops_golden["pqmf_syn_y"]).max() < 5e-6
    # impulses: one non-zero product per output sample -> bit-exact
    assert np.array_equal(O.pqmf_synthesis(ops_golden["pqmf_syn_impulse_x"]), ops_golden["pqmf_syn_impulse_y"])
    assert np.array_equal(O.pqmf_analysis(ops_golden["pqmf_ana_impulse_x"]), ops_golden["pqmf_ana_impulse_y"])
