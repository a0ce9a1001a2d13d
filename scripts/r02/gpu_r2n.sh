#!/bin/bash
# round 2, call N: N-tile width for the 256-wide layers (Basis-MelGAN / MelGAN): NT=256 (one CTA per tile, mt=2) vs NT=128 x 2
OUT=gpurun_out
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2n_layers_${m}_$label.json > $OUT/r2n_bench_${m}_$label.json 2> $OUT/r2n_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2n_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2n_layers_${m}_$label.json"))["layers"]
    pair=sum(x["ms"] for x in L if ".stack.4" in x["name"] or ".stack.3" in x["name"])
    dil=sum(x["ms"] for x in L if ".stack.2" in x["name"] or ".stack.1" in x["name"])
    up=sum(x["ms"] for x in L if x["K"]==2 and x["N"]>=64)
    print("%-14s %-8s ms/step %.2f clk %s (ms*GHz %.2f) | dilated %.3f | pair %.3f | convT %.3f | sum %.2f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], d["ms_per_step"]*d["clocks"]["sm_mhz"]/1e3, dil, pair, up, sum(x["ms"] for x in L)))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2n_bench_${m}_$label.err").read()[-1500:])
PY
}
ab nt256 basis-melgan FV_X=0
ab nt128 basis-melgan FV_NT_MAX=128
ab nt256 melgan FV_X=0
ab nt128 melgan FV_NT_MAX=128
ab nt128 hifigan FV_NT_MAX=128
ab nt256 hifigan FV_X=0
timeout 600 env FV_NT_MAX=128 python -m pytest tests -m gpu -q -x -k "model_forward and (melgan or basis)" 2>&1 | tail -3
