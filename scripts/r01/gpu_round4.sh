#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -q -x -k "fused_resblock1" 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error|timeout" | head -20 > $OUT/pytest_fused_$TAG.log
tail -4 $OUT/pytest_fused_$TAG.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/pytest_$TAG.log
tail -4 $OUT/pytest_$TAG.log
for m in hifigan multiband-hifigan; do
timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_${m}_$TAG.json > $OUT/bench_${m}_$TAG.json 2> $OUT/bench_${m}_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${m}_$TAG.json").read().strip().splitlines()[-1])
    print("$m", "ms/step %.2f  samples/s %.3e  e2e %.3e  algTF %.1f  frac %.4f"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["tflops_algorithmic"], d["roofline"]["frac"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$m", "bench failed", e); print(open("$OUT/bench_${m}_$TAG.err").read()[-1500:])
PY
done
FV_NO_FUSE=1 timeout 300 python bench.py --steps 6 --warmup 3 --skip-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nofuse ms/step %.2f'%d['ms_per_step'])"
