#!/bin/bash
TAG=${1:-l2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
for mbs in 4096 256 96 64 40 24; do
  FV_L2_BUDGET_MB=$mbs timeout 300 python bench.py --steps 6 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_hifigan_${TAG}_$mbs.json > $OUT/bench_hifigan_${TAG}_$mbs.json 2> $OUT/bench_hifigan_${TAG}_$mbs.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_hifigan_${TAG}_$mbs.json").read().strip().splitlines()[-1])
    print("L2 budget $mbs MB: ms/step %.2f  samples/s %.3e e2e %.3e launches %d"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"]))
except Exception as e:
    print("$mbs failed", e); print(open("$OUT/bench_hifigan_${TAG}_$mbs.err").read()[-800:])
PY
done
for mbs in 4096 64; do
  FV_L2_BUDGET_MB=$mbs timeout 300 python bench.py --model basis-melgan --steps 6 --warmup 3 --skip-cpu-baseline > $OUT/bench_basis_${TAG}_$mbs.json 2> $OUT/bench_basis_${TAG}_$mbs.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_basis_${TAG}_$mbs.json").read().strip().splitlines()[-1])
    print("basis L2 budget $mbs MB: ms/step %.2f  samples/s %.3e"%(d["ms_per_step"], d["value"]))
except Exception as e:
    print("$mbs failed", e)
PY
done
