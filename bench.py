#!/usr/bin/env python
"""bench.py — headline benchmark of the generator forward path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--model hifigan|basis-melgan|...]

Headline line (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE.json configs[1] — HiFi-GAN light
(conf/hifigan/light.yaml) inference, batch = 32 synthetic 80 x 1000 mels per GPU, fp32 in / fp32 out, seeded synthetic
weights of the exact architecture.  Metric: audio samples / second (whole job).  For N > 1 the utterance batch is sharded:
every rank runs its own 32 utterances (weak scaling), one NCCL broadcast of the packed weights at init, no per-step
collective; time = max over ranks.

The same JSON line also carries (default --model only):
  `workloads`  (N = 1)  the other BASELINE configs the metric names, each with value / e2e / roofline / cpu_baseline:
               configs[2] Basis-MelGAN light B=64 forward(), configs[3] Multiband-HiFi-GAN light + PQMF B=64;
  `strong`     configs[3] at a FIXED global batch of 64 utterances sharded over the N ranks (shard_range), with the
               1-GPU time of the same batch measured in the same run on rank 0 -> speed-up and efficiency;
  `latency_b1` (N = 1)  batch-1, T = 1000 latency / RTF of all four generators: eager launch chain vs CUDA-graph replay
               (the reference's published metric is batch-1 RTF, bin/test.py:123-132).

One JSON line on stdout (rank 0).  `value` is device-resident throughput (CUDA events around each step, L2 flushed
between steps); `e2e` goes through the public API with pinned-host mel in and waveform out inside the timed region.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # model_name: (yaml, per-GPU batch, frames, BASELINE.json config it corresponds to, key in tests/golden/specs.json)
    "hifigan": ("conf/hifigan/light.yaml", 32, 1000, "configs[1] HiFi-GAN light B=32 T=1000", "hifigan-light"),
    "basis-melgan": ("conf/basis-melgan/light.yaml", 64, 1000, "configs[2] Basis-MelGAN light B=64 T=1000 forward()",
                     "basis-melgan-light"),
    "multiband-hifigan": ("conf/multiband-hifigan/light.yaml", 64, 1000,
                          "configs[3] Multiband-HiFi-GAN light + PQMF synthesis B=64 T=1000", "multiband-hifigan-light"),
    "melgan": ("conf/melgan/original.yaml", 32, 1000, "MelGAN original B=32 T=1000 (configs[0] is its CPU case)",
               "melgan-original"),
}
SAMPLES_PER_FRAME = 240
STRONG_MODEL, STRONG_BATCH = "multiband-hifigan", 64      # BASELINE.json configs[3]: fixed batch, 1 -> 8 GPUs


def load_yaml(path):
    import yaml
    with open(os.path.join(REPO, path)) as f:
        return yaml.safe_load(f)


def load_specs():
    with open(os.path.join(REPO, "tests", "golden", "specs.json")) as f:
        return json.load(f)


def _synthetic():
    """fastvocoder_b200/synthetic.py loaded by path: numpy only — the reference arm must not import the product package."""
    spec = importlib.util.spec_from_file_location("_fv_synthetic", os.path.join(REPO, "fastvocoder_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def spec_weights_np(name, seed=0):
    """Seeded synthetic weights of the architecture, keyed by the reference's folded state_dict names (specs.json was
    written from the live reference classes by oracle/gen_golden.py) — identical to what the native arm loads."""
    S = _synthetic()
    key = WORKLOADS[name][4]
    spec = [(n, tuple(s)) for n, s in load_specs()[key]["spec_folded"] if not n.startswith("pqmf.")]
    return S.synth_state_dict(spec, seed)


def synthetic_weights(model, seed=0):
    from fastvocoder_b200.synthetic import synth_state_dict
    spec = [(n, s) for n, s, _ in model._spec]
    return {k: torch.from_numpy(v) for k, v in synth_state_dict(spec, seed).items()}


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tensor_tflops": d["bf16_tflops"],
                "tensor_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tensor_tflops": 1590.0, "tensor_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own CPU path (real classes when a reference tree is present, else the ATen port)
# ---------------------------------------------------------------------------------------------------------------
def _host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank when N > 1; the CPU arm must still use all host threads."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail and os.environ.get("OMP_NUM_THREADS", "") in ("", "1"):
        torch.set_num_threads(avail)
    return torch.get_num_threads()


def _reference_classes():
    """The UNMODIFIED reference generator classes if a reference tree is reachable (baseline/_ref or /root/reference; the
    GPU box has neither), else None.  One shim: scipy.signal.kaiser moved to scipy.signal.windows (pqmf.py:12)."""
    for root in (os.path.join(REPO, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(root, "model", "generator")):
            try:
                import scipy.signal
                import scipy.signal.windows
                if not hasattr(scipy.signal, "kaiser"):
                    scipy.signal.kaiser = scipy.signal.windows.kaiser
                if root not in sys.path:
                    sys.path.insert(0, root)
                from model.generator import (BasisMelGANGenerator, HiFiGANGenerator, MelGANGenerator,
                                             MultiBandHiFiGANGenerator)
                return {"root": root, "hifigan": HiFiGANGenerator, "multiband-hifigan": MultiBandHiFiGANGenerator,
                        "melgan": MelGANGenerator, "basis-melgan": BasisMelGANGenerator}
            except Exception:
                return None
    return None


def _build_reference_model(G, name, config, weights_np):
    """Constructor calls of bin/synthesize.py:25-68, then load_state_dict / eval / remove_weight_norm (:69-71)."""
    if name == "melgan":
        m = G[name](in_channels=config["in_channels"], out_channels=config["out_channels"], kernel_size=config["kernel_size"],
                    channels=config["channels"], upsample_scales=config["upsample_scales"],
                    stack_kernel_size=config["stack_kernel_size"], stacks=config["stacks"],
                    use_weight_norm=config["use_weight_norm"], use_causal_conv=config["use_causal_conv"])
    elif name == "basis-melgan":
        m = G[name](basis_signal_weight=torch.zeros(config["L"], config["out_channels"]).float(), L=config["L"],
                    in_channels=config["in_channels"], out_channels=config["out_channels"], kernel_size=config["kernel_size"],
                    channels=config["channels"], upsample_scales=config["upsample_scales"],
                    stack_kernel_size=config["stack_kernel_size"], stacks=config["stacks"],
                    use_weight_norm=config["use_weight_norm"], use_causal_conv=config["use_causal_conv"],
                    transposedconv=config["transposedconv"])
    else:
        m = G[name](resblock_kernel_sizes=config["resblock_kernel_sizes"], upsample_rates=config["upsample_rates"],
                    upsample_initial_channel=config["upsample_initial_channel"], resblock_type=config["resblock_type"],
                    upsample_kernel_sizes=config["upsample_kernel_sizes"],
                    resblock_dilation_sizes=config["resblock_dilation_sizes"], transposedconv=config["transposedconv"],
                    bias=config["bias"])
    m.eval()
    m.remove_weight_norm()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights_np.items()}, strict=False)
    return m


def make_cpu_forward(name, cfg, weights_np):
    """-> (callable x[B,80,T] -> waveform tensor, kind): the reference's CPU implementation of the path."""
    G = _reference_classes()
    if G is not None:
        model = _build_reference_model(G, name, cfg, weights_np)

        def fwd(x):
            y = model(x)
            if name == "multiband-hifigan":
                y = model.pqmf.synthesis(y)                 # multiband_hifigan.py:136
            return y[0] if isinstance(y, tuple) else y
        return fwd, "reference"
    from oracle import torch_port as P
    w = P.to_torch(weights_np)
    f = P.FORWARD[name]

    def fwd(x):
        y = f(w, cfg, x)
        if name == "multiband-hifigan":
            y = P.pqmf_synthesis(y)
        return y[0] if isinstance(y, tuple) else y
    return fwd, "port"


def cpu_throughput(name, cfg, weights_np, frames, utterances, repeats):
    """Best-of-`repeats` throughput of the CPU path on `utterances` utterances (first call is an untimed warm-up)."""
    S = _synthetic()
    fwd, kind = make_cpu_forward(name, cfg, weights_np)
    x = torch.from_numpy(S.synth_mel(utterances, frames, seed=4242))
    best = float("inf")
    with torch.no_grad():
        for i in range(repeats + 1):
            t0 = time.perf_counter()
            fwd(x)
            dt = time.perf_counter() - t0
            if i > 0:
                best = min(best, dt)
    return utterances * frames * SAMPLES_PER_FRAME / best, best, kind


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, SAME config as the native
    arm (the full per-GPU batch every step).  Real reference classes when a reference tree is present, else the ATen port
    (oracle/torch_port.py — /root/reference does not exist on the GPU box).  Does not import the product package."""
    if rank != 0:
        return
    name = args.model
    ypath, B, T, label, _ = WORKLOADS[name]
    B = args.batch or B
    T = args.frames or T
    cfg = load_yaml(ypath)
    cores = _host_threads()
    weights = spec_weights_np(name)
    fwd, kind = make_cpu_forward(name, cfg, weights)
    x = torch.from_numpy(_synthetic().synth_mel(B, T, seed=4242))

    def step():
        with torch.no_grad():
            return fwd(x)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * B * T * SAMPLES_PER_FRAME / dt
    line = {
        "impl": "reference", "metric": "audio samples/sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": label, "generator": name, "yaml": ypath, "batch_per_gpu": B, "global_batch": B, "frames": T,
                   "note": ("reference CPU path on host cores: " +
                            ("the unmodified reference generator classes" if kind == "reference"
                             else "ATen conv ops in the reference's order (oracle/torch_port.py; no reference tree on this box)") +
                            f"; each step = the full batch of {B} utterances")},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{B} utterances x {T} frames per step, {args.steps} steps after {args.warmup} warm-up"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rtf": dt / args.steps / (B * T * 0.01),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def build_model(ctx, name):
    """Rank 0 'loads the checkpoint'; ONE broadcast of the packed weights (the only collective of the job per model)."""
    from fastvocoder_b200 import build_generator
    from fastvocoder_b200.sharding import broadcast_weights
    cfg = load_yaml(WORKLOADS[name][0])
    model = build_generator(name, cfg)
    if ctx.rank == 0:
        model.load_state_dict(synthetic_weights(model), strict=False)
    model.eval()
    model.remove_weight_norm()
    model.to(ctx.dev)
    broadcast_weights(model, src=0)
    model.use_tensor_cores = not ctx.args.no_tc
    return model, cfg


def make_fwd(name, model):
    def fwd(x):
        with torch.no_grad():
            if name == "multiband-hifigan":
                return model(x, synthesize=True)[1]
            y = model(x)
            return y[0] if isinstance(y, tuple) else y
    return fwd


def time_device(ctx, fwd, mel_dev, steps, warmup, active=True):
    """CUDA events around each step, L2 flushed between steps, barrier + synchronize on both sides, max over ranks.
    `active=False` ranks only take part in the barriers / reduction (1-GPU leg of the strong-scaling measurement)."""
    from fastvocoder_b200.sharding import max_over_ranks
    if active:
        for _ in range(warmup):
            fwd(mel_dev)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    torch.cuda.synchronize()
    if active:
        for e0, e1 in ev:
            ctx.flush.zero_()                            # evict L2 between timed iterations (outside the event pair)
            e0.record()
            fwd(mel_dev)
            e1.record()
    torch.cuda.synchronize()
    ctx.barrier()
    t = sum(e0.elapsed_time(e1) for e0, e1 in ev) / 1e3 if active else 0.0
    return max_over_ranks(t, ctx.dev)


def time_e2e(ctx, fwd, mel_host, mel_stage, out_host, steps):
    """Serial form: copy in, compute, copy out back to back on one stream."""
    from fastvocoder_b200.sharding import max_over_ranks
    ctx.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        mel_stage.copy_(mel_host, non_blocking=True)
        y = fwd(mel_stage)
        out_host.copy_(y, non_blocking=True)
    torch.cuda.synchronize()
    t = max_over_ranks(time.perf_counter() - t0, ctx.dev)
    ctx.barrier()
    return t


def time_e2e_pipelined(ctx, model, fwd, mel_host, out_hosts, steps):
    """The public serving API (fastvocoder_b200.pipeline.HostPipeline): every step still copies its mel batch host -> device and
    its waveform device -> host, but on separate streams, so the copies of steps i+1 / i-1 overlap the compute of step i."""
    from fastvocoder_b200.pipeline import HostPipeline
    from fastvocoder_b200.sharding import max_over_ranks
    pipe = HostPipeline(model, fwd=fwd)
    for i in range(2):                                   # warm-up: stream-private workspace, buffers
        pipe.submit(mel_host, out_hosts[i % len(out_hosts)])
    pipe.finish()
    ctx.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        pipe.submit(mel_host, out_hosts[i % len(out_hosts)])
    pipe.finish()
    torch.cuda.synchronize()
    t = max_over_ranks(time.perf_counter() - t0, ctx.dev)
    ctx.barrier()
    return t


def roofline_of(ctx, name, model, mel_dev, B, T, label, profile_out=""):
    """Per-layer CUDA-event profile (separate pass) -> roofline object of the dominant kernel class."""
    peaks = measured_peaks()
    prof = model.profile_forward(mel_dev)
    prof = model.profile_forward(mel_dev)                # second pass: warm instruction caches
    by_kernel = {}
    for r in prof:
        k = by_kernel.setdefault(r["kernel"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        k["ms"] += r["ms"]; k["flops"] += r["flops"]; k["bytes"] += r["bytes"]; k["launches"] += 1
    total_ms = sum(k["ms"] for k in by_kernel.values())
    dom = max(by_kernel, key=lambda k: by_kernel[k]["ms"])
    d = by_kernel[dom]
    kernel_names = {"tcgen05": "conv_tc2_kernel (tcgen05, split-fp16 x3, persistent warp-specialised)",
                    "tcgen05-fused-unit": "conv_tc3_fused_kernel (tcgen05 fused ResBlock1 unit: conv1+lrelu+conv2+residual, "
                                          "TMA-fed split fp16 hi/lo activations)",
                    "tcgen05-fused-stack": "conv_tc3_fused_kernel<IO_STACK> (tcgen05 fused ResidualStack: dilated conv + lrelu + "
                                           "[1x1 | skip 1x1] pair, h kept in shared memory)",
                    "ffma": "conv_ffma_kernel (fp32 CUDA cores)"}
    achieved_tf = d["flops"] / (d["ms"] * 1e-3) / 1e12
    peak_tf = peaks["tensor_tflops_sustained"]           # kernel timed inside a long step -> sustained figure
    is_tc = dom != "ffma"
    traffic = None
    tpath = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            t = json.load(open(tpath)).get(name, {}).get(dom)
            if t:   # captured at a smaller batch; DRAM traffic of these kernels is proportional to positions
                traffic = {"bytes_per_launch": t["bytes_per_launch"] * B / t["batch"], "unit": "B",
                           "measured_at_batch": t["batch"], "scaled_to_batch": B, "launch": t["launch"],
                           "algorithmic_bytes": t["algorithmic_bytes"] * B / t["batch"],
                           "source": "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
        except Exception:
            traffic = None
    tc_ms = sum(v["ms"] for k, v in by_kernel.items() if k != "ffma")
    tc_fl = sum(v["flops"] for k, v in by_kernel.items() if k != "ffma")
    roofline = {
        "bound": "tensor", "kernel": kernel_names.get(dom, dom),
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_source": f"{peaks['source']} dense bf16/fp16 GEMM, sustained (MEASURED_PEAKS.json bf16_tflops_sustained)",
        "flops_basis": "algorithmic 2*Cin*Cout*k per output sample (SURVEY 8a); executed UMMA FLOPs are 3x (split-fp16)",
        "passes": 3 if is_tc else 1,
        "executed_mma_frac": (3 * achieved_tf / peak_tf) if is_tc else None,
        "share_of_step": d["ms"] / total_ms, "launches_per_step": d["launches"],
        "avg_launch_ms": d["ms"] / d["launches"],
        "traffic": traffic["bytes_per_launch"] if traffic else None, "traffic_unit": "B/launch", "traffic_detail": traffic,
        "all_tensor_core_kernels": {"tflops": tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
                                    "frac": (tc_fl / (tc_ms * 1e-3) / 1e12 / peak_tf) if tc_ms else None,
                                    "share_of_step": tc_ms / total_ms},
        "hbm_peak_gbs": peaks["hbm_gbs"],
        "by_kernel": {k: {"ms": v["ms"], "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12,
                          "gbs_if_unfused": v["bytes"] / (v["ms"] * 1e-3) / 1e9, "launches": v["launches"]}
                      for k, v in by_kernel.items()},
    }
    if profile_out:
        with open(profile_out, "w") as f:
            json.dump({"workload": label, "B": B, "T": T, "layers": prof, "by_kernel": roofline["by_kernel"]}, f, indent=1)
    return roofline, prof


def cpu_baseline_of(name, cfg, B, T, utts, repeats):
    """The CPU path beside the GPU number, on a bounded sample of the same workload (rank 0, N = 1 only)."""
    cores = _host_threads()
    weights = spec_weights_np(name)
    v, dt, kind = cpu_throughput(name, cfg, weights, T, utts, repeats)
    what = "unmodified reference classes" if kind == "reference" else "ATen port of the reference CPU path (oracle/torch_port.py)"
    out = {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
           "sample": f"{utts} of the {B} utterances x {T} frames, best of {repeats} after warm-up ({dt:.2f} s each), {what}"}
    torch.set_num_threads(1)                              # bin/test.py times the reference single-threaded, one utterance
    v1, dt1, _ = cpu_throughput(name, cfg, weights, T, 1, 1)
    torch.set_num_threads(cores)
    out["value_1thread"] = v1
    out["sample_1thread"] = f"1 utterance x {T} frames, 1 thread ({dt1:.2f} s)"
    return out


def run_workload(ctx, name, B, T, steps, warmup, detail=True, cpu_utts=16, cpu_repeats=3, profile_out=""):
    """One workload: device-resident throughput, end-to-end throughput, roofline (rank 0), CPU baseline (rank 0, N = 1)."""
    from fastvocoder_b200 import _lib
    from fastvocoder_b200.synthetic import synth_mel
    ypath, _, _, label, _ = WORKLOADS[name]
    model, cfg = build_model(ctx, name)
    fwd = make_fwd(name, model)
    mel_host = torch.from_numpy(synth_mel(B, T, seed=100 + ctx.rank)).pin_memory()
    mel_dev = mel_host.to(ctx.dev)
    samples_per_step = B * T * SAMPLES_PER_FRAME          # per rank
    y = fwd(mel_dev)
    torch.cuda.synchronize()
    out_host = torch.empty(y.shape, dtype=torch.float32).pin_memory()

    for _ in range(warmup):
        fwd(mel_dev)
    sampler = ClockSampler(ctx.local_rank)
    launches0, tc0 = _lib.lib().fv_launch_count(), _lib.lib().fv_tc_launch_count()
    sampler.start()
    t_dev = time_device(ctx, fwd, mel_dev, steps, 0)
    launches = _lib.lib().fv_launch_count() - launches0           # kernels launched inside the timed region (all steps)
    tc_launches = _lib.lib().fv_tc_launch_count() - tc0
    value = ctx.world * samples_per_step * steps / t_dev
    mel_stage = torch.empty_like(mel_dev)
    t_e2e_serial = time_e2e(ctx, fwd, mel_host, mel_stage, out_host, steps)
    out_host2 = torch.empty(y.shape, dtype=torch.float32).pin_memory()
    t_e2e = time_e2e_pipelined(ctx, model, fwd, mel_host, [out_host, out_host2], steps)
    clocks = sampler.result()
    e2e_value = ctx.world * samples_per_step * steps / t_e2e

    res = {"workload": label, "generator": name, "value": value, "unit": "samples/s", "ms_per_step": 1e3 * t_dev / steps,
           "batch_per_gpu": B, "global_batch": B * ctx.world, "frames": T,
           "rtf": t_dev / steps / (B * T * 0.01),
           "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(mel_host.numel() * 4),
                   "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": 1e3 * t_e2e / steps,
                   "mode": "HostPipeline (public API): per-step pinned H2D + D2H on copy streams, overlapping the neighbouring "
                           "steps' compute; every step makes both PCIe trips inside the timed region",
                   "serial_value": ctx.world * samples_per_step * steps / t_e2e_serial,
                   "serial_ms_per_step": 1e3 * t_e2e_serial / steps},
           "gpu_launches": int(launches), "tc_launches": int(tc_launches), "clocks": clocks}
    if ctx.rank == 0:
        res["tflops_algorithmic"] = model.forward_flops(B, T) / (res["ms_per_step"] * 1e-3) / 1e12
    prof = []
    if ctx.rank == 0 and detail:
        res["roofline"], prof = roofline_of(ctx, name, model, mel_dev, B, T, label, profile_out)
    if ctx.rank == 0 and ctx.world == 1 and detail and not ctx.args.skip_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline_of(name, cfg, B, T, min(B, cpu_utts), cpu_repeats)
    res["_model"], res["_prof"], res["_mel_dev"] = model, prof, mel_dev
    return res


def run_strong(ctx, steps, warmup):
    """BASELINE.json configs[3] at a FIXED global batch (64 utterances, Multiband-HiFi-GAN light + PQMF synthesis):
    (a) rank 0 alone runs the whole batch (the 1-GPU time, same box, same run), (b) the batch is sharded contiguously
    over the N ranks (shard_range) — no per-step collective.  speed-up = t1 / tN, efficiency = speed-up / N."""
    from fastvocoder_b200.sharding import shard_range
    from fastvocoder_b200.synthetic import synth_mel
    name, Bg = STRONG_MODEL, STRONG_BATCH
    T = WORKLOADS[name][2]
    model, _ = build_model(ctx, name)
    fwd = make_fwd(name, model)
    mel_all = torch.from_numpy(synth_mel(Bg, T, seed=7))
    total = Bg * T * SAMPLES_PER_FRAME
    full = mel_all.to(ctx.dev) if ctx.rank == 0 else None
    t1 = time_device(ctx, fwd, full, steps, warmup, active=ctx.rank == 0)
    del full
    lo, hi = shard_range(Bg, ctx.rank, ctx.world)
    if ctx.world == 1:
        tn = t1
    else:
        shard = mel_all[lo:hi].contiguous().to(ctx.dev)
        tn = time_device(ctx, fwd, shard, steps, warmup, active=hi > lo)
    return {"workload": WORKLOADS[name][3] + f", fixed global batch {Bg} sharded over {ctx.world} GPU(s)",
            "scaling": "strong", "global_batch": Bg, "n_gpus": ctx.world, "utterances_per_gpu": [Bg // ctx.world, -(-Bg // ctx.world)],
            "value": total * steps / tn, "unit": "samples/s", "ms_per_step": 1e3 * tn / steps,
            "one_gpu_ms_per_step": 1e3 * t1 / steps, "one_gpu_value": total * steps / t1,
            "speedup_vs_1gpu": t1 / tn, "efficiency": t1 / tn / ctx.world,
            "note": "1-GPU leg timed on rank 0 of the same job (other ranks idle at the barrier); device events, max over ranks"}


def run_latency_b1(ctx, steps):
    """Batch-1 latency (the reference's published metric is batch-1 RTF, bin/test.py:123-132): T = 1000 frames = 10 s of
    audio, eager launch chain vs CUDA-graph replay of the same chain (model.graphed), per generator."""
    from fastvocoder_b200.synthetic import synth_mel
    T, out = 1000, {}
    for name in ("hifigan", "multiband-hifigan", "melgan", "basis-melgan"):
        try:
            model, _ = build_model(ctx, name)
            x = torch.from_numpy(synth_mel(1, T, seed=11)).to(ctx.dev)
            eager = make_fwd(name, model)

            def timed(fn, n):
                for _ in range(3):
                    fn(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    fn(x)
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / n
            ms_eager = timed(eager, steps)
            entry = {"frames": T, "audio_s": T * 0.01, "eager_ms": ms_eager, "eager_rtf": ms_eager * 1e-3 / (T * 0.01)}
            try:
                g = model.graphed(x, synthesize=True) if name == "multiband-hifigan" else model.graphed(x)
                y_e = eager(x)
                y_g = g(x)
                y_g = y_g[1] if name == "multiband-hifigan" else (y_g[0] if isinstance(y_g, tuple) else y_g)
                entry["graph_bit_identical"] = bool(torch.equal(y_e, y_g))
                ms_graph = timed(g, steps)
                entry.update({"graph_ms": ms_graph, "graph_rtf": ms_graph * 1e-3 / (T * 0.01)})
            except Exception as e:   # noqa: BLE001
                entry["graph_error"] = str(e)[:200]
            out[name] = entry
        except Exception as e:       # noqa: BLE001
            out[name] = {"error": str(e)[:200]}
    return out


def hbm_kernels_of(ctx, T, profile_rows):
    """Standalone HBM-bound pieces of the path: GB/s of algorithmic bytes vs the measured copy peak."""
    from fastvocoder_b200 import PQMF
    from fastvocoder_b200.synthesizer import encode_16bits
    peaks = measured_peaks()
    dev = ctx.dev
    pq = PQMF().to(dev)
    Bq, Lq = 64, 60 * T                               # configs[3]: sub-bands [64, 4, 60 T] -> wave [64, 1, 240 T]
    xs = torch.rand(Bq, 4, Lq, device=dev) - 0.5
    xw = torch.rand(Bq, 1, 4 * Lq, device=dev) - 0.5
    wav = torch.rand(Bq * 4 * Lq, device=dev) - 0.5

    def timed(fn, nbytes, reps=5):
        fn()
        best = float("inf")
        for _ in range(reps):
            ctx.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        gbs = nbytes / (best * 1e-3) / 1e9
        return {"ms": best, "GB/s": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "algorithmic_bytes": nbytes}
    with torch.no_grad():
        hk = {
            "pqmf_synthesis": timed(lambda: pq.synthesis(xs), 2 * xs.numel() * 4),        # 960 B in + 960 B out / frame
            "pqmf_analysis": timed(lambda: pq.analysis(xw), 2 * xw.numel() * 4),
            "encode_16bits": timed(lambda: encode_16bits(wav), wav.numel() * (4 + 4 + 2)),  # peak pass + scale pass + int16 out
            "peak": {"hbm_gbs": peaks["hbm_gbs"], "source": peaks["source"]},
            "shape": f"B={Bq}, sub-bands 4 x {Lq}, L2 flushed before each run, best of 5",
        }
    for r in profile_rows:                            # the generator's own narrow output conv (conv_post / LastLayer)
        if r["N"] <= 4 and r["ms"] > 0:
            gbs = r["bytes"] / (r["ms"] * 1e-3) / 1e9
            hk["output_conv"] = {"layer": r["name"], "kernel": r["kernel"], "Cin": r["Cin"], "N": r["N"], "K": r["K"],
                                 "ms": r["ms"], "GB/s": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"],
                                 "algorithmic_bytes": r["bytes"],
                                 "note": "timed inside the step's per-layer profile (no L2 flush: input is the previous layer's output)"}
    return hk


def _public(res):
    return {k: v for k, v in res.items() if not k.startswith("_")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--model", default="hifigan", choices=list(WORKLOADS))
    ap.add_argument("--no-tc", action="store_true", help="force the exact-fp32 CUDA-core path")
    ap.add_argument("--batch", type=int, default=0, help="override per-GPU batch")
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip the workloads / strong / latency_b1 extras")
    ap.add_argument("--profile-out", default="", help="write the per-layer profile JSON here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    from fastvocoder_b200.sharding import init_distributed
    ctx = Ctx()
    ctx.args = args
    ctx.rank, ctx.local_rank, ctx.world = init_distributed()
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(ctx.local_rank)
    ctx.dev = torch.device("cuda", ctx.local_rank)
    ctx.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=ctx.dev)   # > 126 MB L2

    def barrier():
        if ctx.world > 1:
            dist.barrier(device_ids=[ctx.local_rank])
    ctx.barrier = barrier

    name = args.model
    ypath, B, T, label, _ = WORKLOADS[name]
    B = args.batch or B
    T = args.frames or T
    extras = name == "hifigan" and not args.headline_only and not args.batch and not args.frames and not args.no_tc

    head = run_workload(ctx, name, B, T, args.steps, args.warmup, detail=True, profile_out=args.profile_out)
    hbm_kernels = hbm_kernels_of(ctx, T, head["_prof"]) if ctx.rank == 0 else None
    del head["_model"], head["_mel_dev"]
    torch.cuda.empty_cache()

    workloads, strong, latency = None, None, None
    if extras:
        x_steps = min(args.steps, 10)
        if ctx.world == 1:
            workloads = {}
            for other in ("basis-melgan", "multiband-hifigan"):
                r = run_workload(ctx, other, WORKLOADS[other][1], WORKLOADS[other][2], x_steps, args.warmup, detail=True,
                                 cpu_utts=8, cpu_repeats=2)
                workloads[other] = _public(r)
                del r
                torch.cuda.empty_cache()
        strong = run_strong(ctx, x_steps, args.warmup)
        torch.cuda.empty_cache()
        if ctx.world == 1 and ctx.rank == 0:
            latency = run_latency_b1(ctx, 20)

    if ctx.rank == 0:
        line = {
            "metric": "audio samples/sec", "value": head["value"], "unit": "samples/s", "n_gpus": ctx.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32 (tcgen05 layers: fp16 hi+lo split x3, fp32 accumulate)" if not args.no_tc else "f32",
            "data": "synthetic",
            "config": {"workload": label, "generator": name, "yaml": ypath, "batch_per_gpu": B, "global_batch": B * ctx.world,
                       "frames": T, "parallelism": f"batch-shard x{ctx.world}", "l2": "explicit 256 MiB flush between timed steps",
                       "timing": "cuda events per step, max over ranks"},
            "rtf": head["rtf"], "tflops_algorithmic": head.get("tflops_algorithmic"),
            "gpu_launches": head["gpu_launches"], "tc_launches": head["tc_launches"],
            "e2e": head["e2e"], "clocks": head["clocks"], "roofline": head.get("roofline"),
            "cpu_baseline": head.get("cpu_baseline"), "hbm_kernels": hbm_kernels,
            "workloads": workloads, "strong": strong, "latency_b1": latency,
        }
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
