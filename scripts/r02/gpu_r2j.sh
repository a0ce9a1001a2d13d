#!/bin/bash
# round 2, call J: conv_tc2 split-input layers: M tiles per CTA tile (accumulator chains vs epilogue overlap); B=1 latency with PDL
OUT=gpurun_out
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2j_layers_${m}_$label.json > $OUT/r2j_bench_${m}_$label.json 2> $OUT/r2j_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2j_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2j_layers_${m}_$label.json"))["layers"]
    def g(k): return sum(x["ms"] for x in L if x["Cin"]==128 and x["N"]==128 and x["K"]==k)
    print("%-18s %-8s ms/step %.2f clk %s | C128 k3 %.3f k7 %.3f k11 %.3f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], g(3), g(7), g(11)))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2j_bench_${m}_$label.err").read()[-1500:])
PY
}
ab base hifigan FV_X=0
ab mt3 hifigan FV_TC2_MT_XS=3
ab mt4 hifigan FV_TC2_MT_XS=4
ab mt1 hifigan FV_TC2_MT_XS=1
ab base2 hifigan FV_X=0
python - <<'PY'
import os, subprocess, json, sys
code = r'''
import sys, json, torch
sys.path.insert(0, ".")
import bench
from types import SimpleNamespace
ctx = bench.Ctx(); ctx.args = SimpleNamespace(no_tc=False); ctx.rank = 0; ctx.local_rank = 0; ctx.world = 1
ctx.dev = torch.device("cuda", 0)
print(json.dumps(bench.run_latency_b1(ctx, 20)))
'''
for env in ({}, {"FV_PDL": "1"}):
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **env))
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print("latency_b1", env, {k: (round(v.get("eager_ms", 0), 3), round(v.get("graph_ms", 0), 3), v.get("graph_bit_identical")) for k, v in d.items()})
    except Exception as e:
        print("latency failed", env, e, out.stderr[-800:])
PY
