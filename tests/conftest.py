import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

MODEL_KEYS = ["hifigan-light", "hifigan-large", "multiband-hifigan-light", "multiband-hifigan-large",
              "melgan-original", "basis-melgan-light"]
# non-default architecture switches, goldens from the reference classes with the switch turned on
VARIANT_KEYS = ["hifigan-light-upsamplelayer", "hifigan-light-resblock2", "multiband-hifigan-light-upsamplelayer",
                "melgan-causal", "basis-melgan-upsamplelayer", "basis-melgan-causal-lastlinear",
                # round 2: ResBlock2 with 3-entry dilation lists; MelGAN family bias=False / negative_slope / no final activation
                "hifigan-light-resblock2-dil3", "melgan-nobias-slope01", "basis-melgan-nobias-nofinal"]
ALL_KEYS = MODEL_KEYS + VARIANT_KEYS


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def specs():
    with open(os.path.join(GOLDEN, "specs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ops_golden():
    return dict(np.load(os.path.join(GOLDEN, "ops.npz")))


def load_model_golden(key):
    return dict(np.load(os.path.join(GOLDEN, f"model_{key}.npz")))


def folded_weights(specs, key, seed=0):
    """Regenerate the deterministic weights gen_golden.py loaded into the reference."""
    from fastvocoder_b200.synthetic import synth_state_dict
    spec = [(n, tuple(s)) for n, s in specs[key]["spec_folded"] if not n.startswith("pqmf.") and not n.endswith("num_batches_tracked")]
    return synth_state_dict(spec, seed=seed)
