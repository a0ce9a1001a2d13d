#!/bin/bash
# round 2, call X: full GPU suite on the current binary (fused ResidualStack, vectorised PQMF, weight range guard) + default bench
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $OUT/r2x_pytest.log
cat $OUT/r2x_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2x_bench_default.json 2> $OUT/r2x_bench_default.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2x_bench_default.json").read().strip().splitlines()[-1])
    print("hifigan ms/step %.2f value %.3e | e2e %.3e (%.2f ms) | roofline %s | clocks %s"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k:d["roofline"][k] for k in ("achieved","peak","frac")}, d["clocks"]))
    for k,w in (d.get("workloads") or {}).items():
        print(" ", k, "ms/step %.2f e2e %.2f ms frac %.3f"%(w["ms_per_step"], w["e2e"]["ms_per_step"], w["roofline"]["frac"]))
    print(" hbm", {k:(round(v["ms"],3), round(v["frac_of_hbm_peak"],2)) for k,v in d["hbm_kernels"].items() if isinstance(v,dict) and "ms" in v})
    print(" strong", d["strong"]); print(" latency", {k:(v.get("eager_ms"), v.get("graph_ms")) for k,v in (d["latency_b1"] or {}).items() if isinstance(v,dict)})
except Exception as e:
    print("default bench failed", e); print(open("$OUT/r2x_bench_default.err").read()[-3000:])
PY
