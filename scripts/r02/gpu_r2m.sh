#!/bin/bash
# round 2, call M: MelGAN-family ResidualStacks on the hybrid split path (h via TMA into the pair layer) — parity, then A/B
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "melgan or basis or residual_stack or tc_conv or wide_layer or ragged or bench_shapes or cuda_graph" 2>&1 | tail -4 > $OUT/r2m_pytest.log
cat $OUT/r2m_pytest.log
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2m_layers_${m}_$label.json > $OUT/r2m_bench_${m}_$label.json 2> $OUT/r2m_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2m_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2m_layers_${m}_$label.json"))["layers"]
    pair=sum(x["ms"] for x in L if ".stack.4" in x["name"] or ".stack.3" in x["name"])
    dil=sum(x["ms"] for x in L if ".stack.2" in x["name"] or ".stack.1" in x["name"])
    print("%-14s %-8s ms/step %.2f clk %s (ms*GHz %.2f) | dilated convs %.3f | pair 1x1 %.3f | sum %.2f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], d["ms_per_step"]*d["clocks"]["sm_mhz"]/1e3, dil, pair, sum(x["ms"] for x in L)))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2m_bench_${m}_$label.err").read()[-1500:])
PY
}
ab st1 basis-melgan FV_X=0
ab st0 basis-melgan FV_STACK_SPLIT=0
ab st1 melgan FV_X=0
ab st0 melgan FV_STACK_SPLIT=0
ab st1b basis-melgan FV_X=0
ab st0b basis-melgan FV_STACK_SPLIT=0
