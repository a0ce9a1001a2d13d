#!/bin/bash
# Round-2 evidence run (1 GPU): tests, smoke, the DEFAULT bench line (headline + workloads + strong + latency_b1), per-model
# benches with per-layer profiles, the reference arm, clocks during the run, the ncu launch list of one bench command and
# `ncu --set full` captures of the kernels DESIGN.md discusses.  Numbers printed under ncu are never bench values.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks_$TAG.csv &
SMI=$!
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $OUT/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out $OUT/profile_hifigan_$TAG.json > $OUT/bench_hifigan_$TAG.json 2> $OUT/bench_hifigan_$TAG.err
for m in basis-melgan multiband-hifigan melgan; do
  timeout 600 python bench.py --model $m --steps 10 --warmup 3 --profile-out $OUT/profile_${m}_$TAG.json > $OUT/bench_${m}_$TAG.json 2> $OUT/bench_${m}_$TAG.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-tc --skip-cpu-baseline > $OUT/bench_hifigan_fp32path_$TAG.json 2>/dev/null
FV_SPLIT=0 timeout 600 python bench.py --steps 10 --warmup 3 --headline-only --skip-cpu-baseline --profile-out $OUT/profile_hifigan_nosplit_$TAG.json > $OUT/bench_hifigan_nosplit_$TAG.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_$TAG.json 2>/dev/null
kill $SMI
B="python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --headline-only"
# ncu: launch list of the bench command (cold-cache, serialised: compare shares), then full captures at --batch 8
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $B > $OUT/ncu_launch_$TAG.log 2>&1
# gpurun merges at most 64 MiB back: every capture is reduced ON THE BOX to its raw-page csv + the text summary; the .ncu-rep
# itself is kept for the three kernels DESIGN.md discusses (no --import-source: the SASS listings are committed separately)
KEEP="fused_c32k7"
cap() { # name, kernel regex, skip, extra bench args...
  n=$1; k=$2; s=$3; shift 3
  timeout 600 ncu --set full --clock-control none -k regex:$k -s $s -c 1 -o $OUT/ncu_${n}_$TAG $B "$@" > $OUT/ncu_${n}_$TAG.log 2>&1
  ncu -i $OUT/ncu_${n}_$TAG.ncu-rep --page raw --csv > $OUT/ncu_${n}_$TAG.raw.csv 2>/dev/null
  python scripts/ncu_summary.py $OUT/ncu_${n}_$TAG.ncu-rep > $OUT/ncu_${n}_$TAG.txt 2>&1
  case " $KEEP " in *" $n "*) ;; *) rm -f $OUT/ncu_${n}_$TAG.ncu-rep ;; esac
  tail -c 300 $OUT/ncu_${n}_$TAG.log > $OUT/ncu_${n}_$TAG.log.tail; rm -f $OUT/ncu_${n}_$TAG.log
}
cap fused_c32k7 conv_tc3 12 --batch 8          # TMA-fed split fused unit, C=32 k=7 (resident weights, ping-pong tiles)
cap fused_c16k3 conv_tc3 18 --batch 8          # C=16 k=3
cap fused_c64k11 conv_tc3 6 --batch 8          # C=64 k=11 (streamed weights)
cap tc2_c128k11 conv_tc2 14 --batch 8          # conv_tc2, C=128 k=11
cap tc2_ups2_split conv_tc2 21 --batch 8       # ConvTranspose 64->32 x3 with the split-format epilogue
cap narrow7 conv_narrow7 0 --batch 8           # conv_post 16->1 k7 (HBM-bound)
cap pqmf_syn pqmf_synthesis 0 --batch 8        # hbm_kernels leg of the bench: B=64, 4 x 60000
cap encode16 encode16 0 --batch 8
cap absmax absmax 0 --batch 8
cap tc2_basis_k3 conv_tc2 6 --model basis-melgan --batch 8
cap pqmf_ana pqmf_analysis 0 --batch 8         # hbm_kernels leg: 64 x 240000 -> 4 x 60000
cap stack_c32 conv_tc3 3 --model melgan --batch 8   # fused ResidualStack C=32 (the 4th fused launch of a MelGAN forward)
cap stack_c64 conv_tc3 0 --model melgan --batch 8   # fused ResidualStack C=64
cat $OUT/pytest_$TAG.log; tail -2 $OUT/smoke_$TAG.log
python - <<PY
import json
for m in ("hifigan","basis-melgan","multiband-hifigan","melgan","hifigan_fp32path","hifigan_nosplit","reference"):
    try:
        d=json.loads(open("$OUT/bench_%s_$TAG.json"%m).read().strip().splitlines()[-1])
        print(m, "ms/step %.2f value %.4e e2e %.4e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("clocks"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(m, "failed", e)
PY
# gpurun merges at most 64 MiB back: never exceed it
if [ $(du -sm $OUT | cut -f1) -gt 55 ]; then rm -f $OUT/*.ncu-rep; fi
ls -la $OUT/*_$TAG.ncu-rep $OUT/*_$TAG.raw.csv; du -sh $OUT
