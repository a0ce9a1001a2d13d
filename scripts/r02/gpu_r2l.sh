#!/bin/bash
# round 2, call L: split chain across stages, running sum folded into the prefetched addend — parity, then A/B (two runs each)
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "model_forward or model_inference or batch_equals or bench_shapes or cuda_graph" 2>&1 | tail -4 > $OUT/r2l_pytest.log
cat $OUT/r2l_pytest.log
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2l_layers_${m}_$label.json > $OUT/r2l_bench_${m}_$label.json 2> $OUT/r2l_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2l_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2l_layers_${m}_$label.json"))["layers"]
    ups=sum(x["ms"] for x in L if x["name"].startswith("ups") or x["name"].startswith("conv_p"))
    last=sum(x["ms"] for x in L if x["K"]==11 and x["dil"]==5)
    print("%-18s %-8s ms/step %.2f clk %s (ms*GHz %.2f) | ups+pre+post %.3f | k11 d5 %.3f | sum %.2f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], d["ms_per_step"]*d["clocks"]["sm_mhz"]/1e3, ups, last, sum(x["ms"] for x in L)))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2l_bench_${m}_$label.err").read()[-1500:])
PY
}
ab fin1 hifigan FV_X=0
ab fin0 hifigan FV_SPLIT_FINAL=0
ab fin1b hifigan FV_X=0
ab fin0b hifigan FV_SPLIT_FINAL=0
ab fin1 multiband-hifigan FV_X=0
ab fin0 multiband-hifigan FV_SPLIT_FINAL=0
ab fin1b multiband-hifigan FV_X=0
ab fin0b multiband-hifigan FV_SPLIT_FINAL=0
