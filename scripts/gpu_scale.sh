#!/bin/bash
# scaling sweep launched the way the driver does it: N = 1, 2, 4, 8 back to back (weak scaling, 32 utterances / GPU)
TAG=${1:-sc}
OUT=gpurun_out
mkdir -p $OUT
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --skip-cpu-baseline > $OUT/scale_n${N}_$TAG.json 2> $OUT/scale_n${N}_$TAG.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 10 --warmup 3 > $OUT/scale_n${N}_$TAG.json 2> $OUT/scale_n${N}_$TAG.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/scale_n${N}_$TAG.json").read().strip().splitlines()[-1])
    print("N=$N n_gpus", d["n_gpus"], "ms/step %.2f value %.4e e2e %.4e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("N=$N failed", e); print(open("$OUT/scale_n${N}_$TAG.err").read()[-1500:])
PY
done
