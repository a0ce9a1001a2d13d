#!/bin/bash
OUT=gpurun_out
TAG=${1:-ring2}
mkdir -p $OUT
FV_STALL_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > $OUT/${TAG}_stall.json 2> $OUT/${TAG}_stall.err
grep -A6 "tc3 C=64 K=7 dil=1 \|tc3 C=64 K=11 dil=3 \|tc3 C=64 K=7 dil=5 " $OUT/${TAG}_stall.err | tail -24
run() {  # tag model env...
  local tag=$1 model=$2; shift 2
  env "$@" timeout 300 python bench.py --model $model --steps 8 --warmup 3 --skip-cpu-baseline \
      --profile-out $OUT/${TAG}_prof_${model}_$tag.json > $OUT/${TAG}_${model}_$tag.json 2> $OUT/${TAG}_${model}_$tag.err
}
run m2 hifigan FV_X=0
run m1 hifigan FV_TC3_RING_M=1
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/${TAG}_hifigan_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s ms/step %7.2f  samples/s %.3e  clk %s" % (os.path.basename(f), d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"]))
        print("     hbm", {k: round(v["ms"], 4) for k, v in d["hbm_kernels"].items() if isinstance(v, dict) and "ms" in v})
    except Exception as e:
        print(os.path.basename(f), "failed", e, open(f.replace(".json", ".err")).read()[-600:])
for t in ("m2","m1"):
    d=json.load(open("$OUT/${TAG}_prof_hifigan_%s.json"%t))
    print(t, [(r["K"], r["dil"], round(r["ms"],3)) for r in d["layers"] if r["Cin"]==64 and r["N"]==64 and r["kernel"].startswith("tcgen05-f")])
PY
