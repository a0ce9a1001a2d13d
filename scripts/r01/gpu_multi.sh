#!/bin/bash
# multi-GPU round: N=1 and N=$1 benches launched exactly as the driver does, plus the reference arm
N=${1:-2}
TAG=${2:-mg}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error|Error" | tail -4
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT/bench_n1_$TAG.json 2> $OUT/bench_n1_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
python - <<PY
import json
for f in ("n1","n$N","ref"):
    try:
        d=json.loads(open("$OUT/bench_%s_$TAG.json"%f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d.get("n_gpus"), "ms/step %.2f value %.4e e2e %.4e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), "launches", d.get("gpu_launches"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e); print(open("$OUT/bench_%s_$TAG.err"%f).read()[-2000:])
PY
