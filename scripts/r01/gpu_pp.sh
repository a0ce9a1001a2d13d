#!/bin/bash
# ping-pong tiles for the streamed-weight fused units: unit tests (bounded), full tests, A/B against FV_TC3_PP=0, stall report
OUT=gpurun_out
TAG=${1:-pp}
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
run pp hifigan FV_X=0
run nopp hifigan FV_TC3_PP=0
run pp multiband-hifigan FV_X=0
run nopp multiband-hifigan FV_TC3_PP=0
FV_STALL_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > $OUT/${TAG}_stall.json 2> $OUT/${TAG}_stall.err
grep -A6 "tc3 C=64 K=7 dil=1 \|tc3 C=64 K=11 dil=3 " $OUT/${TAG}_stall.err | tail -14
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/${TAG}_*_*pp.json")):
    if "prof" in f: continue
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s ms/step %7.2f  samples/s %.3e  clk %s" % (os.path.basename(f), d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(os.path.basename(f), "failed", e, open(f.replace(".json", ".err")).read()[-600:])
for t in ("pp","nopp"):
    d=json.load(open("$OUT/${TAG}_prof_hifigan_%s.json"%t))
    print(t, [(r["K"], r["dil"], round(r["ms"],3)) for r in d["layers"] if r["Cin"]==64 and r["N"]==64 and r["kernel"].startswith("tcgen05-f")])
PY
