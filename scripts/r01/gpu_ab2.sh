#!/bin/bash
# A/B of the round-1 late optimisations (one binary, env switches): flattened loaders, tc3 epiB look-ahead,
# programmatic dependent launch, A-collector reuse.  Writes gpurun_out/ab2_*.json and a summary table.
OUT=gpurun_out
mkdir -p $OUT
run() {  # tag model env...
  local tag=$1 model=$2; shift 2
  env "$@" timeout 300 python bench.py --model $model --steps 8 --warmup 3 --skip-cpu-baseline \
      --profile-out $OUT/ab2_prof_${model}_$tag.json > $OUT/ab2_${model}_$tag.json 2> $OUT/ab2_${model}_$tag.err
}
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/ab2_pytest_default.log
FV_PDL=1 FV_A_REUSE=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/ab2_pytest_pdl_reuse.log
run base hifigan FV_LOADER_FLAT=0 FV_EPIB_AHEAD=0
run flat hifigan FV_LOADER_FLAT=1 FV_EPIB_AHEAD=0
run flatahead hifigan FV_LOADER_FLAT=1 FV_EPIB_AHEAD=1
run pdl hifigan FV_PDL=1
run reuse hifigan FV_A_REUSE=1
run pdlreuse hifigan FV_PDL=1 FV_A_REUSE=1
run base basis-melgan FV_LOADER_FLAT=0
run flat basis-melgan FV_LOADER_FLAT=1
run pdlreuse basis-melgan FV_PDL=1 FV_A_REUSE=1
run base multiband-hifigan FV_LOADER_FLAT=0
run flat multiband-hifigan FV_LOADER_FLAT=1
run pdlreuse multiband-hifigan FV_PDL=1 FV_A_REUSE=1
echo "default:"; cat $OUT/ab2_pytest_default.log; echo "pdl+reuse:"; cat $OUT/ab2_pytest_pdl_reuse.log
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/ab2_*_*.json")):
    if "prof" in f: continue
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s ms/step %7.2f  samples/s %.3e  e2e %.3e  clk %s" % (os.path.basename(f), d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(os.path.basename(f), "failed", e, open(f.replace(".json", ".err")).read()[-600:])
PY
