#!/bin/bash
TAG=${1:-x}
OUT=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 21 -c 1 -o $OUT/prof2_c64k7_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu2a_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 2 -c 1 -o $OUT/prof2_c128k3_$TAG \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu2b_$TAG.log 2>&1
ls -la $OUT/prof2_c64k7_$TAG.ncu-rep $OUT/prof2_c128k3_$TAG.ncu-rep
