#!/bin/bash
# round 2, call T: HostPipeline (overlapped PCIe copies) — correctness test + the default bench line
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "host_pipeline or cuda_graph or synthesizer" 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2t_bench_default.json 2> $OUT/r2t_bench_default.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2t_bench_default.json").read().strip().splitlines()[-1])
    print("hifigan ms/step %.2f value %.3e | e2e %.3e (%.2f ms) serial %.3e (%.2f ms) | clocks %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["serial_value"], d["e2e"]["serial_ms_per_step"], d["clocks"]))
    for k,w in (d.get("workloads") or {}).items():
        print(" ", k, "ms/step %.2f e2e %.2f ms serial %.2f ms"%(w["ms_per_step"], w["e2e"]["ms_per_step"], w["e2e"]["serial_ms_per_step"]))
except Exception as e:
    print("default bench failed", e); print(open("$OUT/r2t_bench_default.err").read()[-3000:])
PY
