#!/bin/bash
# round 2, call AA: packed fp32x2 epilogue math (-DFV_PACKED_F32=1 variant binary, FV_LIB) vs the default binary, in one call
OUT=gpurun_out
P=$PWD/fastvocoder_b200/_C/libfv_packed.so
FV_LIB=$P timeout 600 python -m pytest tests -m gpu -q -x -k "fused_resblock1 or (model_forward and hifigan) or batch_equals" 2>&1 | tail -3
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2aa_layers_${m}_$label.json > $OUT/r2aa_bench_${m}_$label.json 2> $OUT/r2aa_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2aa_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2aa_layers_${m}_$label.json"))["layers"]
    fu=[x for x in L if x["kernel"]=="tcgen05-fused-unit"]
    k3=sum(x["ms"] for x in fu if x["K"]==3); k7=sum(x["ms"] for x in fu if x["K"]==7); k11=sum(x["ms"] for x in fu if x["K"]==11)
    print("%-18s %-8s ms/step %.2f clk %s (ms*GHz %.2f) | fused units k3 %.3f k7 %.3f k11 %.3f | sum %.2f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], d["ms_per_step"]*d["clocks"]["sm_mhz"]/1e3, k3, k7, k11, sum(x["ms"] for x in L)))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2aa_bench_${m}_$label.err").read()[-1500:])
PY
}
ab base hifigan FV_X=0
ab packed hifigan FV_LIB=$P
ab base2 hifigan FV_X=0
ab packed2 hifigan FV_LIB=$P
ab base multiband-hifigan FV_X=0
ab packed multiband-hifigan FV_LIB=$P
