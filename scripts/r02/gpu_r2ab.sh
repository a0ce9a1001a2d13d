#!/bin/bash
# round 2, call AB: re-sweep of existing planner knobs on the FINAL binary (conditions changed since they were first measured)
OUT=gpurun_out
ab() { # label env...
  label=$1; shift 1
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2ab_layers_$label.json > $OUT/r2ab_bench_$label.json 2> $OUT/r2ab_bench_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2ab_bench_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2ab_layers_$label.json"))["layers"]
    fu=[x for x in L if x["kernel"]=="tcgen05-fused-unit"]
    def s(C,K): return sum(x["ms"] for x in fu if x["Cin"]==C and x["K"]==K)
    ups=sum(x["ms"] for x in L if x["name"].startswith("ups"))
    c128=sum(x["ms"] for x in L if x["Cin"]==128 and x["N"]==128)
    print("%-10s ms/step %.2f clk %s | C16 k3/7/11 %.3f %.3f %.3f | C32 %.3f %.3f %.3f | C64 %.3f %.3f %.3f | C128 convs %.3f | ups %.3f (%s) | sum %.2f"%(
      "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], s(16,3), s(16,7), s(16,11), s(32,3), s(32,7), s(32,11), s(64,3), s(64,7), s(64,11), c128, ups,
      " ".join("%.3f"%x["ms"] for x in L if x["name"].startswith("ups")), sum(x["ms"] for x in L)))
except Exception as e:
    print("$label", "bench failed", e); print(open("$OUT/r2ab_bench_$label.err").read()[-800:])
PY
}
ab base FV_X=0
ab pp3 FV_TC3_PP=3
ab mtxs1 FV_TC2_MT_XS=1
ab wait0 FV_WAIT_HINT=0
ab wait2 FV_WAIT_HINT=2
ab reuse FV_A_REUSE=1
ab pdl1 FV_PDL=1
ab base2 FV_X=0
