#!/bin/bash
# round 2, call U: fused ResidualStack kernel (IO_STACK) — parity, then A/B against the two-launch hybrid split path
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fused_residual_stack or fused_stack_model or residual_stack" 2>&1 | tail -6 > $OUT/r2u_pytest_a.log
cat $OUT/r2u_pytest_a.log
grep -q "passed" $OUT/r2u_pytest_a.log && ! grep -q "failed" $OUT/r2u_pytest_a.log || { echo "kernel tests failed: skipping the rest"; exit 0; }
timeout 1200 python -m pytest tests -m gpu -q -x -k "melgan or ragged or bench_shapes or cuda_graph or batch_equals or edge_lengths or config0 or fused_resblock1" 2>&1 | tail -6 > $OUT/r2u_pytest_b.log
cat $OUT/r2u_pytest_b.log
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2u_layers_${m}_$label.json > $OUT/r2u_bench_${m}_$label.json 2> $OUT/r2u_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2u_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2u_layers_${m}_$label.json"))["layers"]
    st=sum(x["ms"] for x in L if ".stack." in x["name"] or "skip_layer" in x["name"])
    fs=sum(x["ms"] for x in L if x["kernel"]=="tcgen05-fused-stack")
    print("%-14s %-8s ms/step %.2f clk %s (ms*GHz %.2f) | all stack layers %.3f | fused-stack launches %.3f | sum %.2f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], d["ms_per_step"]*d["clocks"]["sm_mhz"]/1e3, st, fs, sum(x["ms"] for x in L)))
    if "$label"=="f1":
        for x in L:
            if x["kernel"]=="tcgen05-fused-stack": print("    %-40s C=%d d=%d ms=%.3f algTF=%.1f"%(x["name"], x["Cin"], x["dil"], x["ms"], x["flops"]/x["ms"]/1e9))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2u_bench_${m}_$label.err").read()[-1500:])
PY
}
ab f1 melgan FV_X=0
ab f0 melgan FV_STACK_FUSED=0
ab f1b melgan FV_X=0
ab m2 melgan FV_STACK_M=2
ab f1 basis-melgan FV_X=0
FV_STALL_DEBUG=1 timeout 300 python bench.py --model melgan --steps 1 --warmup 1 --skip-cpu-baseline --headline-only --batch 8 2> $OUT/r2u_stall_melgan.txt > /dev/null
grep -A12 "tc3-stack" $OUT/r2u_stall_melgan.txt | head -80
