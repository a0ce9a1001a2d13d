#!/bin/bash
# One gpurun round: parity tests, benches (tensor-core and fp32 paths), per-layer profile, ncu launch list + full capture.
# usage (from the repo root, on the GPU box):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|error" | head -60 > $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --profile-out $OUT/profile_hifigan_$TAG.json > $OUT/bench_hifigan_$TAG.json 2> $OUT/bench_hifigan_$TAG.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-tc --skip-cpu-baseline > $OUT/bench_hifigan_notc_$TAG.json 2> $OUT/bench_hifigan_notc_$TAG.err
timeout 600 python bench.py --model basis-melgan --steps 10 --warmup 3 --profile-out $OUT/profile_basis_$TAG.json > $OUT/bench_basis_$TAG.json 2> $OUT/bench_basis_$TAG.err
timeout 600 python bench.py --model multiband-hifigan --steps 5 --warmup 3 --skip-cpu-baseline > $OUT/bench_mb_$TAG.json 2> $OUT/bench_mb_$TAG.err
timeout 600 python bench.py --model melgan --steps 5 --warmup 3 --skip-cpu-baseline > $OUT/bench_melgan_$TAG.json 2> $OUT/bench_melgan_$TAG.err
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1
# full capture of the dominant kernel: a few wide (C=128) and a few narrow (C=16) launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 4 -c 4 -o $OUT/prof_wide_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_wide_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 64 -c 3 -o $OUT/prof_narrow_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_narrow_$TAG.log 2>&1
tail -3 $OUT/pytest_$TAG.log
cat $OUT/bench_hifigan_$TAG.json | head -c 3000
echo
cat $OUT/bench_hifigan_notc_$TAG.json | head -c 600
echo
cat $OUT/bench_basis_$TAG.json | head -c 1500
