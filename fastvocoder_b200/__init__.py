"""fastvocoder_b200 — B200-native (sm_100a) generator forward path for FastVocoder-style vocoders.

Drop-in for the mel -> waveform inference path of xcmyz/FastVocoder: the four generator classes keep the
reference constructors / state_dict keys / forward() / inference(), the compute is hand-written CUDA behind
a C ABI (include/fastvocoder_b200.h).  No CPU fallback.
"""
from .generators import (BasisMelGANGenerator, HiFiGANGenerator, MelGANGenerator,  # noqa: F401
                         MultiBandHiFiGANGenerator, build_generator)
from .pipeline import HostPipeline  # noqa: F401
from .pqmf import PQMF  # noqa: F401

__version__ = "0.1.0"
