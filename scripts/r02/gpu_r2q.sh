#!/bin/bash
# round 2, call Q: conv1(i+1) / conv2(i) interleaved issue on the k = 3 fused units — parity, then A/B
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fused_resblock1 or resblock1_golden or (model_forward and hifigan) or batch_equals or bench_shapes" 2>&1 | tail -4 > $OUT/r2q_pytest.log
cat $OUT/r2q_pytest.log
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2q_layers_${m}_$label.json > $OUT/r2q_bench_${m}_$label.json 2> $OUT/r2q_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2q_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2q_layers_${m}_$label.json"))["layers"]
    def grp(c,k): return sum(x["ms"] for x in L if x["kernel"]=="tcgen05-fused-unit" and x["Cin"]==c and x["K"]==k)
    print("%-18s %-6s ms/step %.2f clk %s (ms*GHz %.2f) | k3 units: C64 %.3f C32 %.3f C16 %.3f | sum %.2f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], d["ms_per_step"]*d["clocks"]["sm_mhz"]/1e3, grp(64,3), grp(32,3), grp(16,3), sum(x["ms"] for x in L)))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2q_bench_${m}_$label.err").read()[-1500:])
PY
}
ab pair1 hifigan FV_X=0
ab pair0 hifigan FV_TC3_PAIR=0
ab pair1 multiband-hifigan FV_X=0
ab pair0 multiband-hifigan FV_TC3_PAIR=0
ab pair1b hifigan FV_X=0
ab pair0b hifigan FV_TC3_PAIR=0
