#!/bin/bash
# round 2, call I: wide (C=128) ResBlock stages on the split / TMA chain through conv_tc2 — parity, then A/B
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "model_forward or model_inference or batch_equals or bench_shapes or ragged or tc_conv or test_conv1d" 2>&1 | tail -8 > $OUT/r2i_pytest.log
cat $OUT/r2i_pytest.log
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2i_layers_${m}_$label.json > $OUT/r2i_bench_${m}_$label.json 2> $OUT/r2i_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2i_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2i_layers_${m}_$label.json"))["layers"]
    c128=sum(x["ms"] for x in L if x["Cin"]==128 and x["N"]==128)
    print("%-18s %-8s ms/step %.2f clk %s | C128 convs %.3f ms | launches %d"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], c128, d["gpu_launches"]))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2i_bench_${m}_$label.err").read()[-1500:])
PY
}
ab wide1 hifigan FV_X=0
ab wide0 hifigan FV_SPLIT_WIDE=0
ab wide1epi1 hifigan FV_TC2_EPI=1
ab wide1 multiband-hifigan FV_X=0
ab wide0 multiband-hifigan FV_SPLIT_WIDE=0
ab wide1b hifigan FV_X=0
