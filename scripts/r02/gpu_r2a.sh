#!/bin/bash
# round 2, call A: TMA probe + fresh baseline numbers of the round-1 binary on this pod
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2a_smi.txt 2>&1
timeout 120 scripts/probes/tma_probe > $OUT/r2a_tma_probe.txt 2>&1; echo "tma_probe rc=$?" >> $OUT/r2a_tma_probe.txt
cat $OUT/r2a_tma_probe.txt
for m in hifigan basis-melgan multiband-hifigan melgan; do
  timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/r2a_layers_$m.json > $OUT/r2a_bench_$m.json 2> $OUT/r2a_bench_$m.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2a_bench_$m.json").read().strip().splitlines()[-1])
    print("$m", "ms/step %.2f  samples/s %.3e  e2e %.3e  algTF %.1f  frac %.4f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["tflops_algorithmic"], d["roofline"]["frac"]), d["clocks"])
except Exception as e:
    print("$m", "bench failed", e); print(open("$OUT/r2a_bench_$m.err").read()[-1500:])
PY
done
