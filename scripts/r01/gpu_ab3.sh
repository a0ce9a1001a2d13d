#!/bin/bash
# after the instruction-count pass (32-bit plane offsets, opaque base pointers, cheaper wait loop): tests + benches
OUT=gpurun_out
TAG=${1:-ab3}
mkdir -p $OUT
run() {  # tag model env...
  local tag=$1 model=$2; shift 2
  env "$@" timeout 300 python bench.py --model $model --steps 8 --warmup 3 --skip-cpu-baseline \
      --profile-out $OUT/${TAG}_prof_${model}_$tag.json > $OUT/${TAG}_${model}_$tag.json 2> $OUT/${TAG}_${model}_$tag.err
}
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/${TAG}_pytest.log
run def hifigan FV_X=0
run sleep hifigan FV_WAIT_HINT=2
run flat2 hifigan FV_LOADER_FLAT=2
run flat0 hifigan FV_LOADER_FLAT=0
run def basis-melgan FV_X=0
run flat2 basis-melgan FV_LOADER_FLAT=2
run def multiband-hifigan FV_X=0
run def melgan FV_X=0
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
