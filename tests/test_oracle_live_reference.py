"""Extra pinning of the oracle against the LIVE reference (build container only; skipped where /root/reference is absent,
e.g. on the GPU box): randomly drawn small architectures of all four generator families — including the non-default
switches — are instantiated from the unmodified reference classes, loaded with seeded synthetic weights and run in fp64;
the numpy restatement must agree to 1e-9 and the ATen port (fp32) to fp32 noise.  The committed golden vectors
(tests/golden) pin the shipped configs; this pins the restatement's generality."""
import os
import random
import sys
import zlib

import numpy as np
import pytest
import torch

REF = os.environ.get("FV_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model", "generator")),
                                reason="reference tree not present (GPU box): the golden vectors cover parity there")


def _reference():
    import scipy.signal
    import scipy.signal.windows
    scipy.signal.kaiser = scipy.signal.windows.kaiser          # pqmf.py:12 imports the name SciPy >= 1.13 removed
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import model.generator as G
    return G


def _draw(rng, family, trial):
    """Shapes are random; the switches cycle with the trial index so that every one of them is exercised."""
    if family in ("hifigan", "multiband-hifigan"):
        rates = [rng.choice([2, 3, 4, 5]) for _ in range(2)]
        return {"resblock_kernel_sizes": rng.choice([[3, 5], [3, 7], [5]]), "upsample_rates": rates,
                "upsample_initial_channel": rng.choice([16, 32]), "resblock_type": "2" if trial % 2 else "1",
                "upsample_kernel_sizes": [2 * r for r in rates],
                "resblock_dilation_sizes": None, "transposedconv": trial % 3 != 2, "bias": True}
    scales = [rng.choice([2, 3, 4]) for _ in range(2)]
    cfg = {"in_channels": 80, "out_channels": 1, "kernel_size": 7, "channels": rng.choice([[32, 16, 8], [24, 24, 24]]),
           "upsample_scales": scales, "stack_kernel_size": 3, "stacks": rng.choice([1, 2, 3]), "use_weight_norm": True,
           "use_causal_conv": trial % 2 == 1}
    if family == "basis-melgan":
        cfg.update({"L": 2 * rng.choice([3, 5, 6]), "transposedconv": trial % 3 != 2, "lastlinear": trial % 3 == 1})
        cfg["out_channels"] = rng.choice([8, 12]) if cfg["lastlinear"] else cfg["channels"][-1]
    return cfg


def _build(G, family, cfg):
    if family == "melgan":
        return G.MelGANGenerator(in_channels=80, out_channels=1, kernel_size=cfg["kernel_size"], channels=cfg["channels"],
                                 upsample_scales=cfg["upsample_scales"], stack_kernel_size=3, stacks=cfg["stacks"],
                                 use_weight_norm=True, use_causal_conv=cfg["use_causal_conv"])
    if family == "basis-melgan":
        return G.BasisMelGANGenerator(basis_signal_weight=torch.zeros(cfg["L"], cfg["out_channels"]), L=cfg["L"], in_channels=80,
                                      out_channels=cfg["out_channels"], kernel_size=7, channels=cfg["channels"],
                                      upsample_scales=cfg["upsample_scales"], stack_kernel_size=3, stacks=cfg["stacks"],
                                      use_weight_norm=True, use_causal_conv=cfg["use_causal_conv"],
                                      transposedconv=cfg["transposedconv"], lastlinear=cfg["lastlinear"])
    cls = G.HiFiGANGenerator if family == "hifigan" else G.MultiBandHiFiGANGenerator
    return cls(resblock_kernel_sizes=cfg["resblock_kernel_sizes"], upsample_rates=cfg["upsample_rates"],
               upsample_initial_channel=cfg["upsample_initial_channel"], resblock_type=cfg["resblock_type"],
               upsample_kernel_sizes=cfg["upsample_kernel_sizes"], resblock_dilation_sizes=cfg["resblock_dilation_sizes"],
               transposedconv=cfg["transposedconv"], bias=True)


@pytest.mark.parametrize("family", ["hifigan", "multiband-hifigan", "melgan", "basis-melgan"])
def test_oracle_matches_live_reference_on_random_architectures(family):
    from fastvocoder_b200.synthetic import synth_mel, synth_state_dict
    from oracle import np_oracle as O
    from oracle import torch_port as P
    G = _reference()
    rng = random.Random(zlib.crc32(family.encode()))          # stable across processes
    for trial in range(6):
        cfg = _draw(rng, family, trial)
        if cfg.get("resblock_dilation_sizes", 0) is None:
            nd = 3 if cfg["resblock_type"] == "1" else 2
            cfg["resblock_dilation_sizes"] = [[1, 3, 5][:nd] for _ in cfg["resblock_kernel_sizes"]]
        torch.manual_seed(trial)
        model = _build(G, family, cfg).eval()
        model.remove_weight_norm()
        spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()
                if not k.startswith("pqmf.") and not k.endswith("num_batches_tracked")]
        weights = synth_state_dict(spec, seed=100 + trial)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=False)
        # MelGAN family: ReflectionPad1d of the d = 9 (causal: 18-sample) stack needs a longer first stage than its pad
        T = rng.choice([5, 9, 14]) if family in ("hifigan", "multiband-hifigan") else rng.choice([12, 20, 27])
        mel = synth_mel(2, T, seed=trial)
        with torch.no_grad():
            ref = model.double()(torch.from_numpy(mel).double())
        ref = [t.numpy() for t in (ref if isinstance(ref, tuple) else (ref,))]
        w64 = {k: v.astype(np.float64) for k, v in weights.items()}
        got = O.FORWARD[family](w64, cfg, mel.astype(np.float64))
        got = list(got) if isinstance(got, tuple) else [got]
        for a, b in zip(got, ref):
            assert a.shape == b.shape, (family, cfg, a.shape, b.shape)
            assert np.abs(a - b).max() < 1e-9, (family, cfg, float(np.abs(a - b).max()))
        with torch.no_grad():
            port = P.FORWARD[family](P.to_torch(weights), cfg, torch.from_numpy(mel))
        port = [t.numpy() for t in (port if isinstance(port, tuple) else (port,))]
        for a, b in zip(port, ref):
            assert np.abs(a - b).max() < 5e-5 * max(1.0, float(np.abs(b).max())), (family, cfg)
