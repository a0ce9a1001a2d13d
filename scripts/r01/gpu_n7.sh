#!/bin/bash
OUT=gpurun_out
TAG=${1:-n7}
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/${TAG}_pytest.log
cat $OUT/${TAG}_pytest.log
run() {  # tag model env...
  local tag=$1 model=$2; shift 2
  env "$@" timeout 300 python bench.py --model $model --steps 8 --warmup 3 --skip-cpu-baseline \
      --profile-out $OUT/${TAG}_prof_${model}_$tag.json > $OUT/${TAG}_${model}_$tag.json 2> $OUT/${TAG}_${model}_$tag.err
}
run on hifigan FV_X=0
run off hifigan FV_NARROW7=0
run on multiband-hifigan FV_X=0
run off multiband-hifigan FV_NARROW7=0
run on melgan FV_X=0
run off melgan FV_NARROW7=0
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/${TAG}_*_o*.json")):
    if "prof" in f: continue
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        oc = d["hbm_kernels"].get("output_conv", {})
        print("%-40s ms/step %7.2f  samples/s %.3e  clk %s | output conv %s %.3f ms %.0f GB/s" % (os.path.basename(f), d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], oc.get("kernel"), oc.get("ms", 0), oc.get("GB/s", 0)))
    except Exception as e:
        print(os.path.basename(f), "failed", e, open(f.replace(".json", ".err")).read()[-600:])
PY
