#!/bin/bash
# Round-end evidence run (1 GPU): tests, benches of all four models with per-layer profiles, clocks during the run,
# ncu launch list of the bench command, ncu --set full captures of the two dominant kernels.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks_$TAG.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $OUT/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out $OUT/profile_hifigan_$TAG.json > $OUT/bench_hifigan_$TAG.json 2> $OUT/bench_hifigan_$TAG.err
for m in basis-melgan multiband-hifigan melgan; do
  timeout 600 python bench.py --model $m --steps 10 --warmup 3 --profile-out $OUT/profile_${m}_$TAG.json > $OUT/bench_${m}_$TAG.json 2> $OUT/bench_${m}_$TAG.err
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-tc --skip-cpu-baseline > $OUT/bench_hifigan_fp32path_$TAG.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_reference_$TAG.json 2>/dev/null
kill $SMI
# ncu: launch list of the bench command (cold-cache, serialised: compare shares), then full captures
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -s 9 -c 2 -o $OUT/ncu_fused_unit_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_f_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 14 -c 2 -o $OUT/ncu_tc2_c128k11_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_t_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 6 -c 2 -o $OUT/ncu_tc2_basis_$TAG \
    python bench.py --model basis-melgan --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_b_$TAG.log 2>&1
cat $OUT/pytest_$TAG.log; tail -2 $OUT/smoke_$TAG.log
python - <<PY
import json
for m in ("hifigan","basis-melgan","multiband-hifigan","melgan","hifigan_fp32path","reference"):
    try:
        d=json.loads(open("$OUT/bench_%s_$TAG.json"%m).read().strip().splitlines()[-1])
        print(m, "ms/step %.2f value %.4e e2e %.4e"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("clocks"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(m, "failed", e)
PY
