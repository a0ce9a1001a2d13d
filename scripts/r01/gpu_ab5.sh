#!/bin/bash
OUT=gpurun_out
TAG=${1:-ab5}
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -20 > $OUT/${TAG}_pytest.log
timeout 600 python bench.py --model multiband-hifigan --steps 8 --warmup 3 > $OUT/${TAG}_mb.json 2> $OUT/${TAG}_mb.err
timeout 600 python bench.py --steps 8 --warmup 3 --skip-cpu-baseline > $OUT/${TAG}_hifigan.json 2> $OUT/${TAG}_hifigan.err
cat $OUT/${TAG}_pytest.log
python - <<PY
import json
for m in ("mb","hifigan"):
    try:
        d=json.loads(open("$OUT/${TAG}_%s.json"%m).read().strip().splitlines()[-1])
        print(m, "ms/step %.2f value %.3e e2e %.3e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), d["clocks"])
        print("   hbm", json.dumps(d.get("hbm_kernels")))
        print("   cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(m, "failed", e); print(open("$OUT/${TAG}_%s.err"%m).read()[-1500:])
PY
