#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference generators.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python oracle/gen_golden.py

It imports the reference's ``model.generator`` package (one shim:
``scipy.signal.kaiser`` was removed in SciPy >= 1.13 and the reference's
pqmf.py:12 still imports it), builds each generator from the reference's own
YAML files, loads deterministic synthetic weights (fastvocoder_b200.synthetic,
numpy PCG64 -> regenerated identically in tests, never shipped), runs
``forward`` / ``inference`` on seeded inputs in fp32 and fp64 and stores the
outputs.  TEST INFRASTRUCTURE: nothing here is imported by the product.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import warnings

import numpy as np
import scipy.signal
import scipy.signal.windows
import torch
import yaml

warnings.filterwarnings("ignore")
REF = os.environ.get("FV_REFERENCE", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "tests", "golden")
scipy.signal.kaiser = scipy.signal.windows.kaiser
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

from model.generator import (BasisMelGANGenerator, HiFiGANGenerator,  # noqa: E402
                             MelGANGenerator, MultiBandHiFiGANGenerator)
from model.generator import modules as ref_modules  # noqa: E402
from model.generator.pqmf import PQMF  # noqa: E402

from fastvocoder_b200.synthetic import synth_mel, synth_state_dict  # noqa: E402

torch.set_num_threads(8)

MODELS = {
    # key: (model_name as in bin/synthesize.py, reference yaml)
    "hifigan-light": ("hifigan", "conf/hifigan/light.yaml"),
    "hifigan-large": ("hifigan", "conf/hifigan/large.yaml"),
    "multiband-hifigan-light": ("multiband-hifigan", "conf/multiband-hifigan/light.yaml"),
    "multiband-hifigan-large": ("multiband-hifigan", "conf/multiband-hifigan/large.yaml"),
    "melgan-original": ("melgan", "conf/melgan/original.yaml"),
    "basis-melgan-light": ("basis-melgan", "conf/basis-melgan/light.yaml"),
}


def build(model_name, config):
    """Same constructor calls as bin/synthesize.py:25-66."""
    if model_name == "melgan":
        return MelGANGenerator(in_channels=config["in_channels"], out_channels=config["out_channels"],
                               kernel_size=config["kernel_size"], channels=config["channels"],
                               upsample_scales=config["upsample_scales"],
                               stack_kernel_size=config["stack_kernel_size"], stacks=config["stacks"],
                               use_weight_norm=config["use_weight_norm"],
                               use_causal_conv=config["use_causal_conv"])
    if model_name == "hifigan":
        cls = HiFiGANGenerator
    elif model_name == "multiband-hifigan":
        cls = MultiBandHiFiGANGenerator
    elif model_name == "basis-melgan":
        return BasisMelGANGenerator(basis_signal_weight=torch.zeros(config["L"], config["out_channels"]).float(),
                                    L=config["L"], in_channels=config["in_channels"],
                                    out_channels=config["out_channels"], kernel_size=config["kernel_size"],
                                    channels=config["channels"], upsample_scales=config["upsample_scales"],
                                    stack_kernel_size=config["stack_kernel_size"], stacks=config["stacks"],
                                    use_weight_norm=config["use_weight_norm"],
                                    use_causal_conv=config["use_causal_conv"],
                                    transposedconv=config["transposedconv"])
    else:
        raise Exception("no model find!")
    return cls(resblock_kernel_sizes=config["resblock_kernel_sizes"], upsample_rates=config["upsample_rates"],
               upsample_initial_channel=config["upsample_initial_channel"], resblock_type=config["resblock_type"],
               upsample_kernel_sizes=config["upsample_kernel_sizes"],
               resblock_dilation_sizes=config["resblock_dilation_sizes"],
               transposedconv=config["transposedconv"], bias=config["bias"])


# Non-default architecture switches: YAML keys the reference classes accept but no shipped conf
# turns on.  Built by calling the reference constructors directly with the shipped light config + the override.
VARIANTS = {
    # key: (model_name, base yaml, overrides, (B, T))
    "hifigan-light-upsamplelayer": ("hifigan", "conf/hifigan/light.yaml", {"transposedconv": False}, (2, 24)),
    "hifigan-light-resblock2": ("hifigan", "conf/hifigan/light.yaml",
                                {"resblock_type": "2", "resblock_dilation_sizes": [[1, 3], [1, 3], [1, 3]]}, (2, 24)),
    "multiband-hifigan-light-upsamplelayer": ("multiband-hifigan", "conf/multiband-hifigan/light.yaml",
                                              {"transposedconv": False}, (2, 24)),
    "melgan-causal": ("melgan", "conf/melgan/original.yaml", {"use_causal_conv": True}, (2, 12)),
    "basis-melgan-upsamplelayer": ("basis-melgan", "conf/basis-melgan/light.yaml", {"transposedconv": False}, (2, 24)),
    "basis-melgan-causal-lastlinear": ("basis-melgan", "conf/basis-melgan/light.yaml",
                                       {"use_causal_conv": True, "lastlinear": True, "out_channels": 128}, (2, 24)),
    # round 2: ResBlock2 with the constructor-default THREE-entry dilation lists (only dilation[0..1] are used,
    # modules.py:233-245), and the MelGAN-family kwargs bin/synthesize.py never passes (bias, negative_slope, final activation)
    "hifigan-light-resblock2-dil3": ("hifigan", "conf/hifigan/light.yaml", {"resblock_type": "2"}, (2, 24)),
    "melgan-nobias-slope01": ("melgan", "conf/melgan/original.yaml",
                              {"bias": False, "nonlinear_activation_params": {"negative_slope": 0.1},
                               "use_final_nonlinear_activation": False}, (2, 12)),
    "basis-melgan-nobias-nofinal": ("basis-melgan", "conf/basis-melgan/light.yaml",
                                    {"bias": False, "nonlinear_activation_params": {"negative_slope": 0.3},
                                     "use_final_nonlinear_activation": False}, (2, 24)),
}
EXTRA_KW = ("bias", "nonlinear_activation_params", "use_final_nonlinear_activation")


def build_variant(model_name, config):
    extra = {k: config[k] for k in EXTRA_KW if k in config}     # constructor kwargs the CLI never passes
    if model_name == "melgan" and extra:
        return MelGANGenerator(in_channels=config["in_channels"], out_channels=config["out_channels"],
                               kernel_size=config["kernel_size"], channels=config["channels"],
                               upsample_scales=config["upsample_scales"],
                               stack_kernel_size=config["stack_kernel_size"], stacks=config["stacks"],
                               use_weight_norm=config["use_weight_norm"],
                               use_causal_conv=config["use_causal_conv"], **extra)
    if model_name == "basis-melgan":    # bin/synthesize.py never passes `lastlinear`; the class does (basis_melgan.py:41)
        return BasisMelGANGenerator(basis_signal_weight=torch.zeros(config["L"], config["out_channels"]).float(), **extra,
                                    L=config["L"], in_channels=config["in_channels"],
                                    out_channels=config["out_channels"], kernel_size=config["kernel_size"],
                                    channels=config["channels"], upsample_scales=config["upsample_scales"],
                                    stack_kernel_size=config["stack_kernel_size"], stacks=config["stacks"],
                                    use_weight_norm=config["use_weight_norm"],
                                    use_causal_conv=config["use_causal_conv"],
                                    transposedconv=config["transposedconv"],
                                    lastlinear=config.get("lastlinear", False))
    return build(model_name, config)


def run_model_golden(key, model_name, config, B, T, specs, manifest, builder, realmel):
    torch.manual_seed(0)
    model = builder(model_name, config)
    wn_sd = model.state_dict()                       # weight-norm form (checkpoint form)
    model.eval()
    model.remove_weight_norm()                       # bin/synthesize.py:69-71
    folded_sd = model.state_dict()
    specs[key] = {"model_name": model_name, "config": config, "spec_wn": spec_of(wn_sd),
                  "spec_folded": spec_of(folded_sd)}
    spec = [(k, tuple(v.shape)) for k, v in folded_sd.items()
            if not k.startswith("pqmf.") and not k.endswith("num_batches_tracked")]
    weights = synth_state_dict(spec, seed=0)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=False)
    mel = synth_mel(B, T, seed=1)
    x32 = torch.from_numpy(mel)
    arrays = {"mel": mel}
    with torch.no_grad():
        y32 = np_out(model(x32))
        inf32 = np_out(model.inference(mel[0].T.copy()))
        m64 = model.double()
        y64 = np_out(m64(x32.double()))
        inf64 = np_out(m64.inference(torch.from_numpy(mel[0].T.copy()).double()))
        model.float()
    for i, (a, b) in enumerate(zip(y32, y64)):
        arrays[f"forward{i}_f32"] = a
        if i == 0:
            arrays[f"forward{i}_f64"] = b
    arrays["inference_f32"] = inf32[0]
    arrays["inference_f64"] = inf64[0]
    print(key, "forward", [a.shape for a in y32], "peak", [float(np.abs(a).max()) for a in y32],
          "fp32-vs-fp64", [float(np.abs(a - b).max()) for a, b in zip(y32, y64)], "inference", inf32[0].shape)
    np.savez_compressed(os.path.join(OUT, f"model_{key}.npz"), **arrays)
    manifest[key] = {k: list(v.shape) for k, v in arrays.items()}


def variants_main():
    """Add the VARIANTS goldens to an existing tests/golden (does not touch the shipped-config fixtures)."""
    with open(os.path.join(OUT, "specs.json")) as f:
        specs = json.load(f)
    with open(os.path.join(OUT, "manifest.json")) as f:
        manifest = json.load(f)
    only = None
    for a in sys.argv:
        if a.startswith("--only="):          # add new variants without rewriting the existing fixture files
            only = set(a[len("--only="):].split(","))
    for key, (model_name, ypath, over, (B, T)) in VARIANTS.items():
        if only is not None and key not in only:
            continue
        with open(os.path.join(REF, ypath)) as f:
            config = yaml.load(f, Loader=yaml.Loader)
        config.update(over)
        run_model_golden(key, model_name, config, B, T, specs, manifest, build_variant, False)
    with open(os.path.join(OUT, "specs.json"), "w") as f:
        json.dump(specs, f, indent=0, sort_keys=True)
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote variants into", OUT)


def spec_of(sd):
    return [[k, list(v.shape)] for k, v in sd.items()]


def np_out(y):
    if isinstance(y, tuple):
        return [t.detach().cpu().numpy() for t in y]
    return [y.detach().cpu().numpy()]


def main():
    os.makedirs(OUT, exist_ok=True)
    specs = {}
    manifest = {}

    for key, (model_name, ypath) in MODELS.items():
        with open(os.path.join(REF, ypath)) as f:
            config = yaml.load(f, Loader=yaml.Loader)
        torch.manual_seed(0)
        model = build(model_name, config)
        wn_sd = model.state_dict()                       # weight-norm form (checkpoint form)
        model.eval()
        model.remove_weight_norm()                       # bin/synthesize.py:69-71
        folded_sd = model.state_dict()
        specs[key] = {
            "model_name": model_name,
            "config": config,
            "spec_wn": spec_of(wn_sd),
            "spec_folded": spec_of(folded_sd),
        }

        # ---- weight-norm fold golden (small): checkpoint-form tensors -> folded weights
        if key in ("hifigan-light", "melgan-original"):
            names = ["conv_post", "ups.3", "resblocks.9.convs1.2"] if model_name == "hifigan" else \
                    ["melgan.18", "melgan.19.stack.2", "melgan.22.conv"]
            fold = {}
            for n in names:
                fold[n + ".weight_g"] = wn_sd[n + ".weight_g"].numpy()
                fold[n + ".weight_v"] = wn_sd[n + ".weight_v"].numpy()
                fold[n + ".weight"] = folded_sd[n + ".weight"].numpy()
            np.savez_compressed(os.path.join(OUT, f"fold_{key}.npz"), **fold)

        # ---- deterministic scaled weights
        spec = [(k, tuple(v.shape)) for k, v in folded_sd.items() if not k.startswith("pqmf.")]
        weights = synth_state_dict(spec, seed=0)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=False)

        large = key.endswith("large")
        B, T = (1, 8) if large else (2, 24)
        mel = synth_mel(B, T, seed=1)
        x32 = torch.from_numpy(mel)
        arrays = {"mel": mel}
        with torch.no_grad():
            y32 = np_out(model(x32))
            inf32 = np_out(model.inference(mel[0].T.copy()))
            m64 = model.double()
            y64 = np_out(m64(x32.double()))
            inf64 = np_out(m64.inference(torch.from_numpy(mel[0].T.copy()).double()))
            model.float()
        for i, (a, b) in enumerate(zip(y32, y64)):
            arrays[f"forward{i}_f32"] = a
            if i == 0:                       # fp64 "truth" kept for the waveform output only (size)
                arrays[f"forward{i}_f64"] = b
        arrays["inference_f32"] = inf32[0]
        arrays["inference_f64"] = inf64[0]
        print(key, "forward", [a.shape for a in y32], "peak", [float(np.abs(a).max()) for a in y32],
              "fp32-vs-fp64", [float(np.abs(a - b).max()) for a, b in zip(y32, y64)],
              "inference", inf32[0].shape)

        # ---- the real mel crop (resource/test.mel.npy, frames 100..164) through `inference`
        if not large:
            real = np.load(os.path.join(REF, "resource", "test.mel.npy"))     # (80, 585) float64
            crop = np.ascontiguousarray(real[:, 100:164].T).astype(np.float32)  # (64, 80) like mel.T
            with torch.no_grad():
                r32 = model.inference(crop.copy()).numpy()
                r64 = model.double().inference(torch.from_numpy(crop).double()).numpy()
                model.float()
            arrays["realmel_T80"] = crop
            arrays["realmel_inference_f32"] = r32
            arrays["realmel_inference_f64"] = r64
        np.savez_compressed(os.path.join(OUT, f"model_{key}.npz"), **arrays)
        manifest[key] = {k: list(v.shape) for k, v in arrays.items()}

    # ---- per-op goldens -------------------------------------------------------------
    ops = {}
    g = torch.Generator().manual_seed(1234)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale)

    # ConvTranspose1d for every shipped (k, s, p, op) plus the MB-large k != 2s cases
    for (k, s) in [(16, 8), (10, 5), (6, 3), (4, 2), (20, 10), (12, 6), (8, 4), (16, 10), (16, 6)]:
        p, op = s // 2 + s % 2, s % 2
        x = rnd(2, 6, 11)
        w = rnd(6, 5, k, scale=0.3)
        b = rnd(5, scale=0.1)
        y = torch.nn.functional.conv_transpose1d(x, w, b, stride=s, padding=p, output_padding=op)
        y64 = torch.nn.functional.conv_transpose1d(x.double(), w.double(), b.double(), stride=s, padding=p,
                                                   output_padding=op)
        ops[f"convt_k{k}_s{s}_x"] = x.numpy()
        ops[f"convt_k{k}_s{s}_w"] = w.numpy()
        ops[f"convt_k{k}_s{s}_b"] = b.numpy()
        ops[f"convt_k{k}_s{s}_y"] = y.numpy()
        ops[f"convt_k{k}_s{s}_y64"] = y64.numpy()

    # ResBlock1 (reference module) for k in 3,7,11, C=8, L=70 (shorter than the k=11,d=5 receptive field edge cases)
    for k in (3, 7, 11):
        rb = ref_modules.ResBlock1(8, k, (1, 3, 5)).eval()
        sd = {n: (rnd(*v.shape, scale=(0.5 / np.sqrt(8 * k)) if n.endswith("weight") else 0.1))
              for n, v in rb.state_dict().items()}
        rb.load_state_dict(sd)
        x = rnd(2, 8, 70)
        with torch.no_grad():
            y = rb(x)
            y64 = rb.double()(x.double())
        ops[f"resblock1_k{k}_x"] = x.numpy()
        ops[f"resblock1_k{k}_y"] = y.numpy()
        ops[f"resblock1_k{k}_y64"] = y64.numpy()
        for n, v in sd.items():
            ops[f"resblock1_k{k}_p_{n}"] = v.numpy()

    # ResidualStack (reflect padding) d in 1,3,9, C=8, L=24
    for d in (1, 3, 9):
        rs = ref_modules.ResidualStack(kernel_size=3, channels=8, dilation=d).eval()
        sd = {n: (rnd(*v.shape, scale=(0.6 / np.sqrt(v.shape[1] * v.shape[2])) if n.endswith("weight") else 0.1))
              for n, v in rs.state_dict().items()}
        rs.load_state_dict(sd)
        x = rnd(2, 8, 24)
        with torch.no_grad():
            y = rs(x)
            y64 = rs.double()(x.double())
        ops[f"resstack_d{d}_x"] = x.numpy()
        ops[f"resstack_d{d}_y"] = y.numpy()
        ops[f"resstack_d{d}_y64"] = y64.numpy()
        for n, v in sd.items():
            ops[f"resstack_d{d}_p_{n}"] = v.numpy()

    # LastLayer
    ll = ref_modules.LastLayer(8, 1, "LeakyReLU", {"negative_slope": 0.2}, "ReflectionPad1d", 7, {}, True).eval()
    sd = {n: rnd(*v.shape, scale=0.2) for n, v in ll.state_dict().items()}
    ll.load_state_dict(sd)
    x = rnd(2, 8, 19)
    with torch.no_grad():
        ops["lastlayer_y"] = ll(x).numpy()
    ops["lastlayer_x"] = x.numpy()
    for n, v in sd.items():
        ops[f"lastlayer_p_{n}"] = v.numpy()

    # overlap_and_add + BasisSignalLayer
    sig = rnd(3, 17, 30)
    ops["ola_signal"] = sig.numpy()
    ops["ola_out_step15"] = ref_modules.overlap_and_add(sig, 15).numpy()
    sig2 = rnd(2, 5, 12)
    ops["ola2_signal"] = sig2.numpy()
    ops["ola2_out_step4"] = ref_modules.overlap_and_add(sig2, 4).numpy()     # gcd path: 12/4 -> 3 sub-frames
    bw = rnd(30, 16, scale=0.3)
    bl = ref_modules.BasisSignalLayer(bw.clone(), L=30)
    wgt = rnd(2, 9, 16)
    with torch.no_grad():
        ops["basis_w"] = bw.numpy()
        ops["basis_in"] = wgt.numpy()
        ops["basis_out"] = bl(wgt).numpy()

    # PQMF: filters, analysis, synthesis, impulses
    pq = PQMF()
    ops["pqmf_analysis_filter"] = pq.analysis_filter.numpy()
    ops["pqmf_synthesis_filter"] = pq.synthesis_filter.numpy()
    xa = rnd(2, 1, 96, scale=0.5)
    xs = rnd(2, 4, 24, scale=0.5)
    with torch.no_grad():
        ops["pqmf_ana_x"] = xa.numpy()
        ops["pqmf_ana_y"] = pq.analysis(xa).numpy()
        ops["pqmf_syn_x"] = xs.numpy()
        ops["pqmf_syn_y"] = pq.synthesis(xs).numpy()
        imp = torch.zeros(4, 4, 20)
        for k in range(4):
            imp[k, k, 3 + 4 * k] = 0.75 + 0.125 * k          # one impulse per band, different positions
        ops["pqmf_syn_impulse_x"] = imp.numpy()
        ops["pqmf_syn_impulse_y"] = pq.synthesis(imp).numpy()
        impa = torch.zeros(3, 1, 80)
        impa[0, 0, 0] = 1.0
        impa[1, 0, 41] = -0.5
        impa[2, 0, 79] = 0.3
        ops["pqmf_ana_impulse_x"] = impa.numpy()
        ops["pqmf_ana_impulse_y"] = pq.analysis(impa).numpy()
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **ops)

    specs["_pqmf"] = {
        "analysis_sha256": hashlib.sha256(pq.analysis_filter.numpy().tobytes()).hexdigest(),
        "synthesis_sha256": hashlib.sha256(pq.synthesis_filter.numpy().tobytes()).hexdigest(),
    }
    # output-length facts pinned by resource/demo/*.wav for T=585
    specs["_demo_lengths_T585"] = {"hifigan-light": 140400, "multiband-hifigan-light": 140400,
                                   "basis-melgan-light": 140415, "multiband-hifigan-large": 140320}
    with open(os.path.join(OUT, "specs.json"), "w") as f:
        json.dump(specs, f, indent=0, sort_keys=True)
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", OUT, "torch", torch.__version__)


if __name__ == "__main__":
    if "--variants" in sys.argv:
        variants_main()
    else:
        main()
        variants_main()
