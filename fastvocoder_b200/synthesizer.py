"""`Synthesizer` + CLI: drop-in for bin/synthesize.py (and the RTF loop of bin/test.py).

Same constructor, same flags (--checkpoint_path --mel_path --wav_path --model_name --config), same four
output files minus the Griffin-Lim side output (data/audio.py needs librosa/tensorflow and is not part of the
generator path).  Unlike the reference's synthesize.sh (CUDA_VISIBLE_DEVICES=-1) this requires a GPU.
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch
import yaml

from . import _lib
from .generators import build_generator

SAMPLE_RATE = 24000   # hparams.py:10
HOP_SIZE = 240        # hparams.py:9
NUM_MELS = 80         # hparams.py:4
RESCALE_OUT = 0.4     # hparams.py:14


def encode_16bits(x: torch.Tensor, rescale_out: float = 1.0) -> torch.Tensor:
    """data/audio.py:12-14 on the GPU: peak-normalise to rescale_out * 32767 and truncate to int16."""
    x = x.contiguous().float()
    if not x.is_cuda:
        raise _lib.FvError("encode_16bits runs on CUDA only")
    out = torch.empty(x.numel(), dtype=torch.int16, device=x.device)
    scratch = torch.empty(1, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().fv_encode_16bits(_lib.ptr(x), x.numel(), float(rescale_out), _lib.ptr(out),
                                               _lib.ptr(scratch), _lib.current_stream_ptr()), "fv_encode_16bits")
    return out.view(x.shape)


def save_wav(y: torch.Tensor, filename: str, sample_rate: int, rescale_out: float = 1.0):
    """data/audio.py:24-26; quantisation happens on the device, only int16 crosses PCIe."""
    import scipy.io.wavfile
    pcm = encode_16bits(y, rescale_out).cpu().numpy()
    scipy.io.wavfile.write(filename, sample_rate, pcm.astype(np.int16))


class Synthesizer:
    """bin/synthesize.py:17-84."""

    def __init__(self, checkpoint_path, config_path, model_name, device="cuda") -> None:
        self.device = torch.device(device)
        self.pattern = None
        self._bias_cache = {}     # (frames, weights version) -> zero-input waveform on the device
        self.model = self.load_model(checkpoint_path, config_path, model_name)

    def load_model(self, checkpoint_path, config_path, model_name):
        with open(config_path) as f:
            config = yaml.load(f, Loader=yaml.Loader)
        print(f"Loading Model of {model_name}...")
        model = build_generator(model_name, config).to(self.device)
        if checkpoint_path:
            # published Basis-MelGAN checkpoints carry a numpy `pattern` (bin/publish.py:71-75) -> weights_only=False
            ckpt = torch.load(os.path.join(checkpoint_path), map_location="cpu", weights_only=False)
            if model_name == "basis-melgan" and "pattern" in ckpt:
                self.pattern = torch.as_tensor(ckpt["pattern"]).float().to(self.device)
            model.load_state_dict(ckpt["model"])
        self.L = config.get("L")
        model.eval()
        model.remove_weight_norm()
        # batch-1 calls are launch-bound (~50-100 dependent kernels): inference() replays a CUDA graph captured per length
        model.use_cuda_graphs = os.environ.get("FV_CUDA_GRAPHS", "1") != "0"
        return model

    def zero_input_bias(self, frames: int) -> torch.Tensor:
        """`model.inference(zeros(frames, 80))` — the input-independent "bias" waveform of bin/synthesize.py:76-78.
        It depends on the weights and on `frames` only, so it is computed once per length and kept on the device
        (the reference recomputes it on every call; bin/publish.py:67-75 caches the same thing as `pattern`)."""
        key = (int(frames), self.model.packed_weights._version)
        hit = self._bias_cache.get(key)
        if hit is None:
            with torch.no_grad():
                hit = self.model.inference(torch.zeros(int(frames), NUM_MELS))
            if len(self._bias_cache) >= 64:          # bounded: drop the oldest length
                self._bias_cache.pop(next(iter(self._bias_cache)))
            self._bias_cache[key] = hit
        return hit

    def synthesize(self, mel):
        """mel: (T, 80) ndarray -> (est_source, est_source - bias, bias)   (bin/synthesize.py:74-80)."""
        with torch.no_grad():
            frames = int(np.asarray(mel).shape[0]) if not isinstance(mel, torch.Tensor) else int(mel.shape[0])
            bias = self.zero_input_bias(frames)
            est_source = self.model.inference(mel)
            est_source_remove_bias = est_source - bias
        return est_source, est_source_remove_bias, bias

    def synthesize_with_pattern(self, mel):
        """bin/test.py:82-91 (Basis-MelGAN): trim L//2 and subtract the published zero-input pattern."""
        with torch.no_grad():
            est_source = self.model.inference(mel)[:-(self.L // 2)]
            if self.pattern is not None:
                est_source = est_source - self.pattern[:est_source.size(0)]
            else:
                z = torch.zeros_like(torch.from_numpy(np.asarray(mel)).float())
                est_source = est_source - self.model.inference(z)[:-(self.L // 2)]
        return est_source

    def test_rtf(self, mel):
        with torch.no_grad():
            self.model.inference(mel)


def publish_model(checkpoint_path, config_path, model_name, save_path, frames: int = 30000, device="cuda"):
    """bin/publish.py:18-77: for Basis-MelGAN store the zero-input "pattern" (inference on `frames` all-zero mel frames,
    i.e. up to 300 s of audio) next to the weights so synthesis can subtract it without a second pass."""
    with open(config_path) as f:
        config = yaml.load(f, Loader=yaml.Loader)
    model = build_generator(model_name, config).to(device)
    ckpt = torch.load(os.path.join(checkpoint_path), map_location="cpu", weights_only=False)
    model.load_state_dict(ckpt["model"])
    if model_name == "basis-melgan":
        with torch.no_grad():
            bias = model.inference(torch.zeros(frames, config["in_channels"]))   # (16*frames + 1) * 15 samples
        torch.save({"model": model.state_dict(), "pattern": bias.cpu().numpy()}, save_path)
    model.eval()
    model.remove_weight_norm()
    return model


def run_publisher(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--checkpoint_path", type=str)
    parser.add_argument("--model_name", type=str)
    parser.add_argument("--config", type=str)
    parser.add_argument("--save_path", type=str)
    args = parser.parse_args(argv)
    publish_model(args.checkpoint_path, args.config, args.model_name, args.save_path)


def load_mel(path):
    """.npy mel, (80, T) or (T, 80) -> (T, 80)   (bin/test.py:110-113 auto-transpose)."""
    mel = np.load(path)
    if mel.shape[0] == NUM_MELS and mel.shape[1] != NUM_MELS:
        mel = mel.T
    return mel


def run_synthesizer(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--checkpoint_path", type=str)
    parser.add_argument("--mel_path", type=str)
    parser.add_argument("--wav_path", type=str)
    parser.add_argument("--model_name", type=str, help="melgan, hifigan, multiband-hifigan and basis-melgan.")
    parser.add_argument("--config", type=str, help="path to model configuration file")
    args = parser.parse_args(argv)

    synthesizer = Synthesizer(args.checkpoint_path, args.config, args.model_name)
    mel = np.load(args.mel_path)                        # (80, T) like the reference's files
    est_source, est_source_remove_bias, bias = synthesizer.synthesize(mel.T)
    save_wav(est_source, args.wav_path, SAMPLE_RATE, rescale_out=RESCALE_OUT)
    save_wav(est_source_remove_bias, args.wav_path[:-3] + "remove.wav", SAMPLE_RATE, rescale_out=RESCALE_OUT)
    save_wav(bias, args.wav_path[:-3] + "bias.wav", SAMPLE_RATE, rescale_out=RESCALE_OUT)


def run_test(argv=None):
    """RTF loop of bin/test.py:98-132 (10 passes over a folder of mels, batch 1)."""
    parser = argparse.ArgumentParser()
    parser.add_argument("--checkpoint_path", type=str)
    parser.add_argument("--file_path", type=str)
    parser.add_argument("--model_name", type=str)
    parser.add_argument("--config", type=str)
    args = parser.parse_args(argv)
    synthesizer = Synthesizer(args.checkpoint_path, args.config, args.model_name)
    mels, duration = [], 0.0
    for file in sorted(os.listdir(args.file_path)):
        if not file.endswith(".npy"):
            continue
        mel = load_mel(os.path.join(args.file_path, file))
        mels.append(mel)
        duration += (mel.shape[0] * HOP_SIZE) / SAMPLE_RATE
    print(f"duration is {duration}s.")
    for mel in mels:
        synthesizer.test_rtf(mel)                       # warm-up (allocations, lazy binding)
    torch.cuda.synchronize()
    s = time.perf_counter()
    for _ in range(10):
        for mel in mels:
            synthesizer.test_rtf(mel)
    torch.cuda.synchronize()
    cost = time.perf_counter() - s
    print(f"cost time: {cost}s.")
    print(f"rtf is {cost / (10.0 * duration)}.")


if __name__ == "__main__":
    run_synthesizer()
