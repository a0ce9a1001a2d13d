#!/bin/bash
# quick iteration round: tc unit tests (bounded), model tests, HiFi + Basis bench with per-layer profile
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -x -k "tc_conv or tc_conv_transpose or test_conv1d or conv_transpose1d" 2>&1 | grep -E "^E  |FAILED|passed|failed|error|timeout|Error" | head -30 > $OUT/pytest_unit_$TAG.log
tail -3 $OUT/pytest_unit_$TAG.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -40 > $OUT/pytest_$TAG.log
tail -4 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_hifigan_$TAG.json > $OUT/bench_hifigan_$TAG.json 2> $OUT/bench_hifigan_$TAG.err
timeout 600 python bench.py --model basis-melgan --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_basis_$TAG.json > $OUT/bench_basis_$TAG.json 2> $OUT/bench_basis_$TAG.err
python - <<PY
import json
for m in ("hifigan","basis"):
    try:
        d=json.loads(open("$OUT/bench_%s_$TAG.json"%m).read().strip().splitlines()[-1])
        print(m, "ms/step %.2f  samples/s %.3e  e2e %.3e  algTF %.1f  frac %.4f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["tflops_algorithmic"], d["roofline"]["frac"]), d["clocks"])
    except Exception as e:
        print(m, "bench failed", e); print(open("$OUT/bench_%s_$TAG.err"%m).read()[-1500:])
PY
