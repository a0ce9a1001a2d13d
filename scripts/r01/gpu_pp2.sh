#!/bin/bash
OUT=gpurun_out
TAG=${1:-pp2}
mkdir -p $OUT
FV_TC3_PP=2 timeout 300 python -m pytest tests -m gpu -q -x -k "fused_resblock1_unit or resblock1_golden or model_forward or ragged or batch_equals" 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/${TAG}_pytest.log
cat $OUT/${TAG}_pytest.log
run() {  # tag model env...
  local tag=$1 model=$2; shift 2
  env "$@" timeout 300 python bench.py --model $model --steps 8 --warmup 3 --skip-cpu-baseline \
      --profile-out $OUT/${TAG}_prof_${model}_$tag.json > $OUT/${TAG}_${model}_$tag.json 2> $OUT/${TAG}_${model}_$tag.err
}
run pp2 hifigan FV_TC3_PP=2
run pp1 hifigan FV_X=0
run pp2m4 hifigan FV_TC3_PP=2 FV_TC3_PP_M=4
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/${TAG}_hifigan_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-36s ms/step %7.2f  samples/s %.3e  clk %s" % (os.path.basename(f), d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(os.path.basename(f), "failed", e, open(f.replace(".json", ".err")).read()[-600:])
for t in ("pp1","pp2","pp2m4"):
    d=json.load(open("$OUT/${TAG}_prof_hifigan_%s.json"%t))
    print(t, [(r["Cin"], r["K"], r["dil"], round(r["ms"],3)) for r in d["layers"] if r["Cin"]<=32 and r["kernel"].startswith("tcgen05-f")], "total %.2f"%sum(r["ms"] for r in d["layers"]))
PY
