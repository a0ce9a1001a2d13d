#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
for m in hifigan basis-melgan multiband-hifigan melgan; do
timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_${m}_$TAG.json > $OUT/bench_${m}_$TAG.json 2> $OUT/bench_${m}_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${m}_$TAG.json").read().strip().splitlines()[-1])
    print("$m", "ms/step %.2f  samples/s %.3e  e2e %.3e  algTF %.1f  frac %.4f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["tflops_algorithmic"], d["roofline"]["frac"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$m", "bench failed", e); print(open("$OUT/bench_${m}_$TAG.err").read()[-1500:])
PY
done
