#!/bin/bash
# round 2, call H (2 GPUs): the bench launched like the driver does for N > 1 (torchrun, NCCL) -> weak headline + strong object
OUT=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r2h_bench_n2.json 2> $OUT/r2h_bench_n2.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$OUT/r2h_bench_n2.json").read().splitlines() if l.startswith("{")][-1]
    print("N=2 weak: ms/step %.2f value %.3e e2e %.3e n_gpus %d"%(d["ms_per_step"], d["value"], d["e2e"]["value"], d["n_gpus"]))
    print("strong:", json.dumps(d["strong"]))
except Exception as e:
    print("failed", e); print(open("$OUT/r2h_bench_n2.err").read()[-3000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --frames 100 > $OUT/r2h_ref_n2.json 2> $OUT/r2h_ref_n2.err
grep -c '^{' $OUT/r2h_ref_n2.json
