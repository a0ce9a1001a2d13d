#!/bin/bash
# Multi-GPU legs of the scaling sweep on ONE 8-GPU box (gpurun --gpus 8), launched the way the driver does it.  The default bench
# line carries the weak-scaling headline (32 utterances per GPU) AND the `strong` object (configs[3], fixed global batch 64 sharded
# over the ranks, with the 1-GPU time of the same batch measured in the same job).  N = 1 comes from the 1-GPU evidence run.
TAG=${1:-sc2}
OUT=gpurun_out
mkdir -p $OUT
for N in ${NS:-8 4 2}; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline > $OUT/scale_n${N}_$TAG.json 2> $OUT/scale_n${N}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/scale_n${N}_$TAG.json").read().strip().splitlines()[-1])
    s=d.get("strong") or {}
    print("N=$N n_gpus", d["n_gpus"], "weak: ms/step %.2f value %.4e e2e %.4e | strong B=64: %.2f ms (1 GPU %.2f ms) speed-up %.2f efficiency %.3f | clocks %s"%(
        d["ms_per_step"], d["value"], d["e2e"]["value"], s.get("ms_per_step",0), s.get("one_gpu_ms_per_step",0), s.get("speedup_vs_1gpu",0), s.get("efficiency",0), d["clocks"]))
except Exception as e:
    print("N=$N failed", e); print(open("$OUT/scale_n${N}_$TAG.err").read()[-1500:])
PY
done
