#!/bin/bash
# round 2, call P: (1) stall accounting of one Basis-MelGAN forward; (2) strong-scaling emulation on one GPU: the configs[3]
# step at the per-GPU batch sizes of N = 1/2/4/8 (64/32/16/8 utterances) -> efficiency estimate t(64) / (N * t(64/N))
OUT=gpurun_out
FV_STALL_DEBUG=1 timeout 300 python bench.py --model basis-melgan --steps 1 --warmup 3 --skip-cpu-baseline --headline-only > $OUT/r2p_stall.json 2> $OUT/r2p_stall.err
python - <<PY
lines = open("$OUT/r2p_stall.err").read().splitlines()
heads = [i for i, l in enumerate(lines) if l.startswith("[stall] tc")]
start = heads[-16] if len(heads) >= 16 else 0
open("$OUT/r2p_stall.txt", "w").write("\n".join(lines[start:]) + "\n")
PY
rm -f $OUT/r2p_stall.err
cut -c1-330 $OUT/r2p_stall.txt
for pdl in auto 1 0; do
 for b in 64 32 16 8; do
  if [ $pdl = auto ]; then E="FV_X=0"; else E="FV_PDL=$pdl"; fi
  env $E timeout 300 python bench.py --model multiband-hifigan --batch $b --steps 10 --warmup 3 --skip-cpu-baseline --headline-only > $OUT/r2p_mb_b${b}_pdl$pdl.json 2> $OUT/r2p_mb_b${b}_pdl$pdl.err
 done
 python - <<PY
import json
t={}
for b in (64,32,16,8):
    try:
        d=json.loads(open("$OUT/r2p_mb_b%d_pdl$pdl.json"%b).read().strip().splitlines()[-1]); t[b]=(d["ms_per_step"], d["clocks"]["sm_mhz"])
    except Exception as e:
        t[b]=(float("nan"),0)
print("pdl=$pdl", {b: "%.2f ms @%s"%t[b] for b in t}, "efficiency vs B=64:", {64//b: round(t[64][0]/(64/b*t[b][0]),3) for b in t})
PY
done
