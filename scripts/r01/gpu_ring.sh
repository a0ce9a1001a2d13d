#!/bin/bash
# streamed-weight fused units (C=64 k=7/11): unit tests first (bounded), then full tests, then A/B against FV_TC3_RING=0
OUT=gpurun_out
TAG=${1:-ring}
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -q -x -k "fused_resblock1_unit or resblock1_golden" 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error|timeout" | head -20 > $OUT/${TAG}_pytest_unit.log
cat $OUT/${TAG}_pytest_unit.log
if grep -q "failed\|error\|Error" $OUT/${TAG}_pytest_unit.log; then
  timeout 100 python -m pytest tests -m gpu -q -x -k "fused_resblock1_unit" 2>&1 | tail -40
  exit 0
fi
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/${TAG}_pytest.log
cat $OUT/${TAG}_pytest.log
run() {  # tag model env...
  local tag=$1 model=$2; shift 2
  env "$@" timeout 300 python bench.py --model $model --steps 8 --warmup 3 --skip-cpu-baseline \
      --profile-out $OUT/${TAG}_prof_${model}_$tag.json > $OUT/${TAG}_${model}_$tag.json 2> $OUT/${TAG}_${model}_$tag.err
}
run ring hifigan FV_X=0
run noring hifigan FV_TC3_RING=0
run ring multiband-hifigan FV_X=0
run noring multiband-hifigan FV_TC3_RING=0
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/${TAG}_*_*.json")):
    if "prof" in f: continue
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s ms/step %7.2f  samples/s %.3e  e2e %.3e  clk %s" % (os.path.basename(f), d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
        print("     pqmf", {k: round(v["ms"], 4) for k, v in d["hbm_kernels"].items() if isinstance(v, dict) and "ms" in v})
    except Exception as e:
        print(os.path.basename(f), "failed", e, open(f.replace(".json", ".err")).read()[-600:])
PY
