#!/bin/bash
# round 2, call AC: planner rules adopted after the knob re-sweep (ping-pong tiles for split C = 32 k = 3; two accumulator sets for
# upsample layers) — full GPU suite on the new default plans, then in-call A/B of the upsample rule on all four generators
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $OUT/r2ac_pytest.log
cat $OUT/r2ac_pytest.log
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2ac_layers_${m}_$label.json > $OUT/r2ac_bench_${m}_$label.json 2> $OUT/r2ac_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2ac_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2ac_layers_${m}_$label.json"))["layers"]
    ups=[x for x in L if x["K"]==2 and x["N"]>x["Cin"]/2 and x["kernel"]=="tcgen05" and x["N"]>=32 and x["N"]!=x["Cin"]]
    c32k3=sum(x["ms"] for x in L if x["kernel"]=="tcgen05-fused-unit" and x["Cin"]==32 and x["K"]==3)
    print("%-18s %-6s ms/step %.2f clk %s | upsample layers %.3f (%s) | C32k3 units %.3f | sum %.2f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"],
          sum(x["ms"] for x in ups), " ".join("%d->%d:%.3f"%(x["Cin"], x["N"], x["ms"]) for x in ups), c32k3, sum(x["ms"] for x in L)))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2ac_bench_${m}_$label.err").read()[-800:])
PY
}
ab new basis-melgan FV_X=0
ab off basis-melgan FV_TC2_UPS_ACC2=0
ab new melgan FV_X=0
ab off melgan FV_TC2_UPS_ACC2=0
ab new multiband-hifigan FV_X=0
ab off multiband-hifigan FV_TC2_UPS_ACC2=0
ab new hifigan FV_X=0
ab off hifigan FV_TC2_UPS_ACC2=0 FV_TC3_PP=2
