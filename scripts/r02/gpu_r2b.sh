#!/bin/bash
# round 2, call B: split/TMA path — unit tests first (bounded), then the model tests, then A/B bench
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -x -k "fused_resblock1 or resblock1_golden" 2>&1 | tail -15 > $OUT/${TAG}_pytest_unit.log
cat $OUT/${TAG}_pytest_unit.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -40 > $OUT/${TAG}_pytest.log
tail -8 $OUT/${TAG}_pytest.log
for m in hifigan multiband-hifigan; do
 for sp in 1 0; do
  FV_SPLIT=$sp timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/${TAG}_layers_${m}_split$sp.json > $OUT/${TAG}_bench_${m}_split$sp.json 2> $OUT/${TAG}_bench_${m}_split$sp.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_${m}_split$sp.json").read().strip().splitlines()[-1])
    print("$m split=$sp", "ms/step %.2f  samples/s %.3e  e2e %.3e  algTF %.1f  frac %.4f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["tflops_algorithmic"], d["roofline"]["frac"]), d["clocks"])
except Exception as e:
    print("$m", "bench failed", e); print(open("$OUT/${TAG}_bench_${m}_split$sp.err").read()[-1500:])
PY
 done
done
