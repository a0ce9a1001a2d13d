#!/bin/bash
# round 2, call O: stall accounting of one Basis-MelGAN forward (what bounds the C=256 layers?)
OUT=gpurun_out
FV_STALL_DEBUG=1 timeout 300 python bench.py --model basis-melgan --steps 1 --warmup 3 --skip-cpu-baseline --headline-only > $OUT/r2o_stall.json 2> $OUT/r2o_stall.err
python - <<PY
lines = open("$OUT/r2o_stall.err").read().splitlines()
heads = [i for i, l in enumerate(lines) if l.startswith("[stall] tc")]
start = heads[-16] if len(heads) >= 16 else 0
open("$OUT/r2o_stall.txt", "w").write("\n".join(lines[start:]) + "\n")
PY
rm -f $OUT/r2o_stall.err
cat $OUT/r2o_stall.txt | cut -c1-330
