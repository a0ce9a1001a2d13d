#!/bin/bash
# Stall-accounting run: where does each role of the persistent tcgen05 kernels wait?  (FV_STALL_DEBUG=1 launches the
# instrumented instantiation and synchronises after every launch — timings of this run are NOT bench values.)
TAG=${1:-s}
OUT=gpurun_out
mkdir -p $OUT
for m in hifigan basis-melgan; do
  FV_STALL_DEBUG=1 timeout 300 python bench.py --model $m --steps 1 --warmup 3 --skip-cpu-baseline \
      > $OUT/stall_${m}_$TAG.json 2> $OUT/stall_${m}_$TAG.err
  # keep the last forward only (bench runs several)
  python - <<PY
import re
lines = open("$OUT/stall_${m}_$TAG.err").read().splitlines()
heads = [i for i, l in enumerate(lines) if l.startswith("[stall] tc")]
per_fwd = {"hifigan": 57, "basis-melgan": 16}["$m"]
start = heads[-per_fwd] if len(heads) >= per_fwd else 0
open("$OUT/stall_${m}_$TAG.txt", "w").write("\n".join(lines[start:]) + "\n")
print("$m", len(heads), "launch reports; kept", len(lines) - start, "lines")
PY
  rm -f $OUT/stall_${m}_$TAG.err
done
[ -z "$SWEEP" ] && exit 0
# knob sweeps on the real bench (no instrumentation)
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_hifigan_base_$TAG.json > $OUT/bench_hifigan_base_$TAG.json 2> $OUT/bench_hifigan_base_$TAG.err
for M in 1 3; do
  FV_TC3_M=$M timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --profile-out $OUT/profile_hifigan_m${M}_$TAG.json > $OUT/bench_hifigan_m${M}_$TAG.json 2> $OUT/bench_hifigan_m${M}_$TAG.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_hifigan_*_$TAG.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.2f" % d["ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
