#!/bin/bash
# round 2, call G: full GPU suite, then the DEFAULT bench line (workloads / strong / latency_b1) and the reference arm
OUT=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | head -40 > $OUT/r2g_pytest.log
tail -8 $OUT/r2g_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out $OUT/r2g_layers_hifigan.json > $OUT/r2g_bench_default.json 2> $OUT/r2g_bench_default.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2g_bench_default.json").read().strip().splitlines()[-1])
    print("hifigan ms/step %.2f e2e %.3e frac %.4f clocks %s"%(d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]))
    for k,w in (d.get("workloads") or {}).items():
        print(" ", k, "ms/step %.2f value %.3e e2e %.3e frac %.4f cpu %s"%(w["ms_per_step"], w["value"], w["e2e"]["value"], w["roofline"]["frac"], (w.get("cpu_baseline") or {}).get("value")))
    print("  strong", d.get("strong"))
    print("  latency_b1", json.dumps(d.get("latency_b1"))[:1500])
    print("  cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print("default bench failed", e); print(open("$OUT/r2g_bench_default.err").read()[-3000:])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2g_bench_reference.json 2> $OUT/r2g_bench_reference.err
tail -c 1500 $OUT/r2g_bench_reference.json
