#!/bin/bash
# A/B bench of env knobs: usage gpu_ab.sh TAG "ENV1=.. ENV2=.." "ENV=.." ...   (each quoted arg = one configuration)
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
i=0
for cfg in "$@"; do
  i=$((i+1))
  for m in hifigan basis-melgan; do
    env $cfg timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_${m}_${TAG}_$i.json > $OUT/bench_${m}_${TAG}_$i.json 2> $OUT/bench_${m}_${TAG}_$i.err
    python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${m}_${TAG}_$i.json").read().strip().splitlines()[-1])
    print("[$cfg] $m ms/step %.2f" % d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("[$cfg] $m failed", e); print(open("$OUT/bench_${m}_${TAG}_$i.err").read()[-800:])
PY
  done
done
