#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_hifigan_$TAG.json > $OUT/bench_hifigan_$TAG.json 2> $OUT/bench_hifigan_$TAG.err
timeout 300 python bench.py --model basis-melgan --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_basis_$TAG.json > $OUT/bench_basis_$TAG.json 2> $OUT/bench_basis_$TAG.err
python - <<PY
import json
for m in ("hifigan","basis"):
    try:
        d=json.loads(open("$OUT/bench_%s_$TAG.json"%m).read().strip().splitlines()[-1])
        print(m, "ms/step %.2f  samples/s %.3e  e2e %.3e  algTF %.1f  frac %.4f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["tflops_algorithmic"], d["roofline"]["frac"]), d["clocks"])
    except Exception as e:
        print(m, "bench failed", e); print(open("$OUT/bench_%s_$TAG.err"%m).read()[-1500:])
PY
# ncu full: HiFi B=8: tc launches in order: 0 conv_pre, 1 ups0, 2.. C=128 resblocks (18), 20 ups1, 21.. C=64 (18), 39 ups2, 40.. C=32, 58 ups3, 59.. C=16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 16 -c 2 -o $OUT/prof2_c128_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_c128_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 35 -c 2 -o $OUT/prof2_c64_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_c64_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 73 -c 2 -o $OUT/prof2_c16_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_c16_$TAG.log 2>&1
ls -la $OUT/*.ncu-rep | tail -5
