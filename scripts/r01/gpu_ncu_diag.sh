#!/bin/bash
# ncu --set full of the narrow fused units (C=32 all nine + first C=16) and a C=64 k=7 unit (conv1, conv2+residual)
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -s 3 -c 10 -o $OUT/ncu_diag_tc3 -f \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_diag_tc3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 21 -c 2 -o $OUT/ncu_diag_tc2 -f \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --batch 8 > $OUT/ncu_diag_tc2.log 2>&1
ls -la $OUT/ncu_diag_*
