#!/bin/bash
# round 2, call F: batch=single test + per-layer A/B of the fused-unit roles (one box, one call)
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "batch_equals_single or fused_resblock1 or resblock1_golden or hifigan-light" 2>&1 | tail -5 > $OUT/r2f_pytest.log
cat $OUT/r2f_pytest.log
ab() { # label env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --model hifigan --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2f_layers_$label.json > $OUT/r2f_bench_$label.json 2> $OUT/r2f_bench_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2f_bench_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2f_layers_$label.json"))["layers"]
    def grp(c,k): return sum(x["ms"] for x in L if x["kernel"]=="tcgen05-fused-unit" and x["Cin"]==c and x["K"]==k)
    s=" ".join("C%dk%d=%.3f"%(c,k,grp(c,k)) for c in (64,32,16) for k in (3,7,11))
    print("%-10s ms/step %.2f clk %s | %s"%("$label", d["ms_per_step"], d["clocks"]["sm_mhz"], s))
except Exception as e:
    print("$label", "bench failed", e); print(open("$OUT/r2f_bench_$label.err").read()[-800:])
PY
}
ab base FV_X=0
ab oldld FV_LIB=$PWD/fastvocoder_b200/_C/libfv_oldld.so
ab epi1 FV_TC3_EPI=1
ab iss3 FV_TC3_ISSUERS=3
ab iss2 FV_TC3_ISSUERS=2
ab pp2only FV_TC3_PP=1
ab base2 FV_X=0
