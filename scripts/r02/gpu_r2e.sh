#!/bin/bash
# round 2, call E: full GPU suite + headline benches with per-layer profiles (new bench.py)
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -40 > $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
for m in hifigan multiband-hifigan basis-melgan melgan; do
  timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/${TAG}_layers_${m}.json > $OUT/${TAG}_bench_${m}.json 2> $OUT/${TAG}_bench_${m}.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_${m}.json").read().strip().splitlines()[-1])
    print("$m", "ms/step %.2f  samples/s %.3e  e2e %.3e  algTF %.1f  frac %.4f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["tflops_algorithmic"], d["roofline"]["frac"]), d["clocks"])
except Exception as e:
    print("$m", "bench failed", e); print(open("$OUT/${TAG}_bench_${m}.err").read()[-1500:])
PY
done
