#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -s 12 -c 1 -o $OUT/prof3_c16k3_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu3a_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -s 9 -c 1 -o $OUT/prof3_c32k11_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu3b_$TAG.log 2>&1
ls -la $OUT/prof3_*
