#!/bin/bash
# round 2, call R: conv_tc2 accumulators handed back per M tile (single-set plans: N = 256 layers) — parity, then A/B
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "not planner_knobs" 2>&1 | tail -4 > $OUT/r2r_pytest.log
cat $OUT/r2r_pytest.log
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2r_layers_${m}_$label.json > $OUT/r2r_bench_${m}_$label.json 2> $OUT/r2r_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2r_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2r_layers_${m}_$label.json"))["layers"]
    wide=sum(x["ms"] for x in L if x["N"]>=256)
    print("%-18s %-6s ms/step %.2f clk %s (ms*GHz %.2f) | N>=256 layers %.3f | sum %.2f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], d["ms_per_step"]*d["clocks"]["sm_mhz"]/1e3, wide, sum(x["ms"] for x in L)))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2r_bench_${m}_$label.err").read()[-1500:])
PY
}
ab mt1 basis-melgan FV_X=0
ab mt0 basis-melgan FV_TC2_ACC_PER_MT=0
ab mt1 melgan FV_X=0
ab mt0 melgan FV_TC2_ACC_PER_MT=0
ab mt1b basis-melgan FV_X=0
ab mt0b basis-melgan FV_TC2_ACC_PER_MT=0
ab mt1 hifigan FV_X=0
ab mt0 hifigan FV_TC2_ACC_PER_MT=0
