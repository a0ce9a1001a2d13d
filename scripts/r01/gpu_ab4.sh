#!/bin/bash
OUT=gpurun_out
TAG=${1:-ab4}
mkdir -p $OUT
run() {  # tag model env...
  local tag=$1 model=$2; shift 2
  env "$@" timeout 300 python bench.py --model $model --steps 8 --warmup 3 --skip-cpu-baseline \
      --profile-out $OUT/${TAG}_prof_${model}_$tag.json > $OUT/${TAG}_${model}_$tag.json 2> $OUT/${TAG}_${model}_$tag.err
}
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/${TAG}_pytest.log
run def hifigan FV_X=0
run def basis-melgan FV_X=0
run def multiband-hifigan FV_X=0
run def melgan FV_X=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -s 3 -c 1 -o $OUT/${TAG}_ncu_tc3 -f \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/${TAG}_ncu_tc3.log 2>&1
cat $OUT/${TAG}_pytest.log
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/${TAG}_*_*.json")):
    if "prof" in f: continue
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s ms/step %7.2f  samples/s %.3e  e2e %.3e  clk %s" % (os.path.basename(f), d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(os.path.basename(f), "failed", e, open(f.replace(".json", ".err")).read()[-600:])
PY
