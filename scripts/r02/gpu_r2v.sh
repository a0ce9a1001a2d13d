#!/bin/bash
# round 2, call V: single-pass raw copy in the fused ResidualStack loaders; vectorised PQMF kernels; bulk-copy staged conv_narrow7
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "streaming_kernel or pqmf or conv1d or fused_residual_stack or fused_stack_model or melgan or multiband or encode" 2>&1 | tail -6 > $OUT/r2v_pytest.log
cat $OUT/r2v_pytest.log
grep -q "passed" $OUT/r2v_pytest.log && ! grep -q "failed" $OUT/r2v_pytest.log || { echo "tests failed: skipping the rest"; exit 0; }
ab() { # label model env...
  label=$1; m=$2; shift 2
  env "$@" timeout 300 python bench.py --model $m --steps 10 --warmup 3 --skip-cpu-baseline --headline-only --profile-out $OUT/r2v_layers_${m}_$label.json > $OUT/r2v_bench_${m}_$label.json 2> $OUT/r2v_bench_${m}_$label.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2v_bench_${m}_$label.json").read().strip().splitlines()[-1])
    L=json.load(open("$OUT/r2v_layers_${m}_$label.json"))["layers"]
    fs=sum(x["ms"] for x in L if x["kernel"]=="tcgen05-fused-stack")
    nar=[x for x in L if x["N"]<=4]
    hk=d.get("hbm_kernels") or {}
    print("%-18s %-8s ms/step %.2f clk %s | fused-stack %.3f | output conv %s | pqmf syn %.3f ana %.3f ms | sum %.2f"%("$m", "$label", d["ms_per_step"], d["clocks"]["sm_mhz"], fs,
          ", ".join("%s %.3f ms (%s)"%(x["name"], x["ms"], x["kernel"]) for x in nar), hk.get("pqmf_synthesis",{}).get("ms",0), hk.get("pqmf_analysis",{}).get("ms",0), sum(x["ms"] for x in L)))
    if "$label"=="d" and "$m"=="melgan":
        for x in L:
            if x["kernel"]=="tcgen05-fused-stack": print("    %-40s C=%d d=%d ms=%.3f algTF=%.1f"%(x["name"], x["Cin"], x["dil"], x["ms"], x["flops"]/x["ms"]/1e9))
except Exception as e:
    print("$m $label", "bench failed", e); print(open("$OUT/r2v_bench_${m}_$label.err").read()[-1500:])
PY
}
ab d melgan FV_X=0
ab f0 melgan FV_STACK_FUSED=0
ab d hifigan FV_X=0
ab old hifigan FV_NARROW7_STAGED=0 FV_PQMF_V4=0
ab d multiband-hifigan FV_X=0
ab mb1 multiband-hifigan FV_NARROW7_MB=1
ab old multiband-hifigan FV_NARROW7_STAGED=0 FV_PQMF_V4=0
ab d2 hifigan FV_X=0
FV_STALL_DEBUG=1 timeout 300 python bench.py --model melgan --steps 1 --warmup 1 --skip-cpu-baseline --headline-only --batch 8 2> $OUT/r2v_stall_melgan.txt > /dev/null
grep -A4 "tc3-stack" $OUT/r2v_stall_melgan.txt | tail -30
