#!/bin/bash
# Stall-accounting run (round 2): per-role wait breakdown of every launch of one HiFi-GAN forward, split path on / off
TAG=${1:-s2}
OUT=gpurun_out
mkdir -p $OUT
for sp in 1 0; do
  FV_SPLIT=$sp FV_STALL_DEBUG=1 timeout 300 python bench.py --model hifigan --steps 1 --warmup 3 --skip-cpu-baseline --headline-only \
      > $OUT/${TAG}_stall_split$sp.json 2> $OUT/${TAG}_stall_split$sp.err
  python - <<PY
lines = open("$OUT/${TAG}_stall_split$sp.err").read().splitlines()
heads = [i for i, l in enumerate(lines) if l.startswith("[stall] tc")]
per_fwd = 50
start = heads[-per_fwd] if len(heads) >= per_fwd else 0
open("$OUT/${TAG}_stall_split$sp.txt", "w").write("\n".join(lines[start:]) + "\n")
print("split=$sp", len(heads), "launch reports; kept", len(lines) - start, "lines")
PY
  rm -f $OUT/${TAG}_stall_split$sp.err
  python scripts/stall_summary.py $OUT/${TAG}_stall_split$sp.txt > $OUT/${TAG}_stall_split${sp}_summary.txt
done
