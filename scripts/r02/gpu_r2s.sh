#!/bin/bash
# round 2, call S: HBM-side kernels (conv_narrow7 with 8 outputs per thread, vectorised save_wav quantiser) — parity, then A/B
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "conv1d or encode_16bits or synthesizer or model_forward or model_inference or ragged" 2>&1 | tail -4 > $OUT/r2s_pytest.log
cat $OUT/r2s_pytest.log
for q in 2 1; do
 for m in hifigan melgan; do
  FV_NARROW7_Q=$q timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only > $OUT/r2s_bench_${m}_q$q.json 2> $OUT/r2s_bench_${m}_q$q.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2s_bench_${m}_q$q.json").read().strip().splitlines()[-1])
    hk=d["hbm_kernels"]
    print("$m Q=$q ms/step %.2f | output_conv %.3f ms %.0f GB/s (%.2f) | encode_16bits %.3f ms %.0f GB/s (%.2f) | pqmf_syn %.3f ms (%.2f)"%(d["ms_per_step"], hk["output_conv"]["ms"], hk["output_conv"]["GB/s"], hk["output_conv"]["frac_of_hbm_peak"], hk["encode_16bits"]["ms"], hk["encode_16bits"]["GB/s"], hk["encode_16bits"]["frac_of_hbm_peak"], hk["pqmf_synthesis"]["ms"], hk["pqmf_synthesis"]["frac_of_hbm_peak"]))
except Exception as e:
    print("$m q=$q failed", e); print(open("$OUT/r2s_bench_${m}_q$q.err").read()[-1500:])
PY
 done
done
