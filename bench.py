#!/usr/bin/env python
"""bench.py — headline benchmark of the generator forward path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--model hifigan|basis-melgan|...]

Workload (N=1): BASELINE.json configs[1] — HiFi-GAN light (conf/hifigan/light.yaml) inference, batch = 32 synthetic
80 x 1000 mels per GPU, fp32 in / fp32 out, seeded synthetic weights of the exact architecture.  Metric: audio
samples / second (whole job).  For N > 1 the utterance batch is sharded: every rank runs its own 32 utterances
(weak scaling), one NCCL broadcast of the packed weights at init, no per-step collective; time = max over ranks.

One JSON line on stdout (rank 0).  `value` is device-resident throughput (CUDA events around each step, L2 flushed
between steps); `e2e` goes through the public API with pinned-host mel in and waveform out inside the timed region.
"""
from __future__ import annotations

import argparse
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
    # model_name: (yaml, per-GPU batch, frames, BASELINE.json config it corresponds to)
    "hifigan": ("conf/hifigan/light.yaml", 32, 1000, "configs[1] HiFi-GAN light B=32 T=1000"),
    "basis-melgan": ("conf/basis-melgan/light.yaml", 64, 1000, "configs[2] Basis-MelGAN light B=64 T=1000 forward()"),
    "multiband-hifigan": ("conf/multiband-hifigan/light.yaml", 64, 1000,
                          "configs[3] Multiband-HiFi-GAN light + PQMF synthesis B=64 T=1000"),
    "melgan": ("conf/melgan/original.yaml", 32, 1000, "MelGAN original B=32 T=1000 (configs[0] is its CPU case)"),
}
SAMPLES_PER_FRAME = 240


def load_yaml(path):
    import yaml
    with open(os.path.join(REPO, path)) as f:
        return yaml.safe_load(f)


def load_specs():
    with open(os.path.join(REPO, "tests", "golden", "specs.json")) as f:
        return json.load(f)


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


def cpu_port_throughput(model_name, cfg, weights_np, frames, utterances, repeats):
    """Time the ATen port (oracle/torch_port.py — the reference's own CPU ops) on the host cores."""
    from oracle import torch_port as P
    w = P.to_torch(weights_np)
    from fastvocoder_b200.synthetic import synth_mel
    x = torch.from_numpy(synth_mel(utterances, frames, seed=4242))
    fwd = P.FORWARD[model_name]
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats + 1):       # first call is the warm-up
            t0 = time.perf_counter()
            y = fwd(w, cfg, x)
            if model_name == "multiband-hifigan":
                y = P.pqmf_synthesis(y)
            dt = time.perf_counter() - t0
            if _ > 0:
                best = min(best, dt)
    return utterances * frames * SAMPLES_PER_FRAME / best, best


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (ATen port; /root/reference is not on the GPU box) on host cores."""
    if rank != 0:
        return
    name = args.model
    ypath, B, T, label = WORKLOADS[name]
    T = args.frames or T
    cfg = load_yaml(ypath)
    # torchrun exports OMP_NUM_THREADS=1 to every rank when N > 1; the reference arm must still use all host threads it can
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail and os.environ.get("OMP_NUM_THREADS", "") in ("", "1"):
        torch.set_num_threads(avail)
    from fastvocoder_b200 import build_generator
    model = build_generator(name, cfg)
    weights = {k: v.numpy() for k, v in synthetic_weights(model).items()}
    from oracle import torch_port as P
    from fastvocoder_b200.synthetic import synth_mel
    w = P.to_torch(weights)
    sample_utts = 2                                # bounded sample of the batch per step
    x = torch.from_numpy(synth_mel(sample_utts, T, seed=4242))
    fwd = P.FORWARD[name]

    def step():
        with torch.no_grad():
            y = fwd(w, cfg, x)
            if name == "multiband-hifigan":
                y = P.pqmf_synthesis(y)
        return y
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * sample_utts * T * SAMPLES_PER_FRAME / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "audio samples/sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": label, "generator": name, "frames": T,
                   "note": "reference CPU path = ATen conv ops on host cores (oracle/torch_port.py); "
                           f"each step = {sample_utts} of the {B} utterances"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_utts} utterances x {T} frames per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rtf": dt / args.steps / (sample_utts * T * 0.01),
    }
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--profile-out", default="", help="write the per-layer profile JSON here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from fastvocoder_b200.sharding import broadcast_weights, init_distributed, max_over_ranks
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), world_env)
        return
    rank, local_rank, world = init_distributed()
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    from fastvocoder_b200 import _lib, build_generator
    from fastvocoder_b200.synthetic import synth_mel
    name = args.model
    ypath, B, T, label = WORKLOADS[name]
    B = args.batch or B
    T = args.frames or T
    cfg = load_yaml(ypath)
    model = build_generator(name, cfg)
    if rank == 0:                                   # only rank 0 "loads the checkpoint"
        model.load_state_dict(synthetic_weights(model), strict=False)
    model.eval()
    model.remove_weight_norm()
    model.to(dev)
    broadcast_weights(model, src=0)                 # the single collective of the whole job
    model.use_tensor_cores = not args.no_tc

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    mel_host = torch.from_numpy(synth_mel(B, T, seed=100 + rank)).pin_memory()
    mel_dev = mel_host.to(dev)
    samples_per_step = B * T * SAMPLES_PER_FRAME     # per rank
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def fwd(x):
        with torch.no_grad():
            if name == "multiband-hifigan":
                return model(x, synthesize=True)[1]
            y = model(x)
            return y[0] if isinstance(y, tuple) else y

    for _ in range(args.warmup):
        y = fwd(mel_dev)
    torch.cuda.synchronize()
    out_host = torch.empty(y.shape, dtype=torch.float32).pin_memory()

    # ---- device-resident timing -------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = _lib.lib().fv_launch_count()
    tc0 = _lib.lib().fv_tc_launch_count()
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    for e0, e1 in ev:
        flush.zero_()                                # evict L2 between timed iterations (outside the event pair)
        e0.record()
        y = fwd(mel_dev)
        e1.record()
    torch.cuda.synchronize()
    barrier()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    launches = _lib.lib().fv_launch_count() - launches0
    tc_launches = _lib.lib().fv_tc_launch_count() - tc0
    t_dev = max_over_ranks(sum(step_ms) / 1e3, dev)
    value = world * samples_per_step * args.steps / t_dev

    # ---- end to end through the public API: pinned host mel -> waveform on the host --------------------
    mel_stage = torch.empty_like(mel_dev)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mel_stage.copy_(mel_host, non_blocking=True)
        y = fwd(mel_stage)
        out_host.copy_(y, non_blocking=True)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0, dev)
    barrier()
    clocks = sampler.result()
    e2e_value = world * samples_per_step * args.steps / t_e2e

    # ---- per-layer profile (rank 0; separate pass so the event pairs do not perturb the timed region) ------
    roofline, profile_rows = None, []
    if rank == 0:
        peaks = measured_peaks()
        prof = model.profile_forward(mel_dev)
        prof = model.profile_forward(mel_dev)        # second pass: warm instruction caches
        by_kernel = {}
        for r in prof:
            k = by_kernel.setdefault(r["kernel"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            k["ms"] += r["ms"]; k["flops"] += r["flops"]; k["bytes"] += r["bytes"]; k["launches"] += 1
        total_ms = sum(k["ms"] for k in by_kernel.values())
        dom = max(by_kernel, key=lambda k: by_kernel[k]["ms"])
        d = by_kernel[dom]
        kernel_names = {"tcgen05": "conv_tc2_kernel (tcgen05, split-fp16 x3, persistent warp-specialised)",
                        "tcgen05-fused-unit": "conv_tc3_fused_kernel (tcgen05 fused ResBlock1 unit: conv1+lrelu+conv2+residual)",
                        "ffma": "conv_ffma_kernel (fp32 CUDA cores)"}
        achieved_tf = d["flops"] / (d["ms"] * 1e-3) / 1e12
        peak_tf = peaks["tensor_tflops_sustained"]   # kernel timed inside a long step -> sustained figure
        is_tc = dom != "ffma"
        # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), if any
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
        # all tensor-core kernels together (they share the roofline): algorithmic FLOPs / their summed time
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
            # DRAM bytes of ONE launch of that kernel (ncu dram__bytes_read.sum + dram__bytes_write.sum), or null
            "traffic": traffic["bytes_per_launch"] if traffic else None, "traffic_unit": "B/launch", "traffic_detail": traffic,
            "all_tensor_core_kernels": {"tflops": tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
                                        "frac": (tc_fl / (tc_ms * 1e-3) / 1e12 / peak_tf) if tc_ms else None,
                                        "share_of_step": tc_ms / total_ms},
            "hbm_peak_gbs": peaks["hbm_gbs"],
            "by_kernel": {k: {"ms": v["ms"], "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12,
                              "gbs_if_unfused": v["bytes"] / (v["ms"] * 1e-3) / 1e9, "launches": v["launches"]}
                          for k, v in by_kernel.items()},
        }
        profile_rows = prof
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"workload": label, "B": B, "T": T, "layers": prof, "by_kernel": roofline["by_kernel"]}, f,
                          indent=1)

    # ---- standalone HBM-bound pieces of the path: GB/s of algorithmic bytes vs the measured copy peak ----
    hbm_kernels = None
    if rank == 0:
        from fastvocoder_b200 import PQMF
        from fastvocoder_b200.synthesizer import encode_16bits
        peaks = measured_peaks()
        pq = PQMF().to(dev)
        Bq, Lq = 64, 60 * T                           # configs[3]: sub-bands [64, 4, 60 T] -> wave [64, 1, 240 T]
        xs = torch.rand(Bq, 4, Lq, device=dev) - 0.5
        xw = torch.rand(Bq, 1, 4 * Lq, device=dev) - 0.5
        wav = torch.rand(Bq * 4 * Lq, device=dev) - 0.5

        def timed(fn, nbytes, reps=5):
            fn()
            best = float("inf")
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            gbs = nbytes / (best * 1e-3) / 1e9
            return {"ms": best, "GB/s": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "algorithmic_bytes": nbytes}
        with torch.no_grad():
            hbm_kernels = {
                "pqmf_synthesis": timed(lambda: pq.synthesis(xs), 2 * xs.numel() * 4),        # 960 B in + 960 B out / frame
                "pqmf_analysis": timed(lambda: pq.analysis(xw), 2 * xw.numel() * 4),
                "encode_16bits": timed(lambda: encode_16bits(wav), wav.numel() * (4 + 4 + 2)),  # peak pass + scale pass + int16 out
                "peak": {"hbm_gbs": peaks["hbm_gbs"], "source": peaks["source"]},
                "shape": f"B={Bq}, sub-bands 4 x {Lq}, L2 flushed before each run, best of 5",
            }
        del xs, xw, wav
        for r in profile_rows:                        # the generator's own narrow output conv (conv_post / LastLayer)
            if r["N"] <= 4 and r["ms"] > 0:
                gbs = r["bytes"] / (r["ms"] * 1e-3) / 1e9
                hbm_kernels["output_conv"] = {"layer": r["name"], "kernel": r["kernel"], "Cin": r["Cin"], "N": r["N"], "K": r["K"],
                                              "ms": r["ms"], "GB/s": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"],
                                              "algorithmic_bytes": r["bytes"],
                                              "note": "timed inside the step's per-layer profile (no L2 flush: input is the previous layer's output)"}

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        weights = {k: v.numpy() for k, v in synthetic_weights(model).items()}
        utts = min(B, 16)                            # bounded sample: ~10-20 s of CPU work in total
        v, dt = cpu_port_throughput(name, cfg, weights, T, utts, repeats=3)
        cpu_baseline = {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{utts} of the {B} utterances x {T} frames, best of 3 after warm-up ({dt:.2f} s each), "
                                  "ATen port of the reference CPU path (oracle/torch_port.py)"}
        nthreads = torch.get_num_threads()           # bin/test.py times the reference single-threaded, one utterance
        torch.set_num_threads(1)
        v1, dt1 = cpu_port_throughput(name, cfg, weights, T, 1, repeats=1)
        torch.set_num_threads(nthreads)
        cpu_baseline["value_1thread"] = v1
        cpu_baseline["sample_1thread"] = f"1 utterance x {T} frames, 1 thread ({dt1:.2f} s)"

    if rank == 0:
        flops_step = model.forward_flops(B, T)
        ms_per_step = 1e3 * t_dev / args.steps
        line = {
            "metric": "audio samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (tcgen05 layers: fp16 hi+lo split x3, fp32 accumulate)" if not args.no_tc else "f32",
            "data": "synthetic",
            "config": {"workload": label, "generator": name, "yaml": ypath, "batch_per_gpu": B, "global_batch": B * world,
                       "frames": T, "parallelism": f"batch-shard x{world}", "l2": "explicit 256 MiB flush between timed steps",
                       "timing": "cuda events per step, max over ranks"},
            "rtf": t_dev / args.steps / (B * T * 0.01),
            "tflops_algorithmic": flops_step / (ms_per_step * 1e-3) / 1e12,
            "gpu_launches": int(launches), "tc_launches": int(tc_launches),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(mel_host.numel() * 4),
                    "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": 1e3 * t_e2e / args.steps},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "hbm_kernels": hbm_kernels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
