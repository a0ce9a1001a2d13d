#!/bin/bash
# round 2, call C: what bounds the k = 3 fused units?  stall reports of one ResBlock branch under planner knobs
OUT=gpurun_out
run() { # label, env..., then C K
  label=$1; shift
  env "$@" FV_STALL_DEBUG=1 python scripts/unit_bench.py $CC $KK 240000 16 3 2>&1 | grep -A5 "dil=1 " | sed "s/^/[$label] /"
}
for shape in "16 3" "32 3" "16 7"; do
  set -- $shape; CC=$1; KK=$2
  echo "=== C=$CC K=$KK"
  run base FV_X=0
  run iss1 FV_TC3_ISSUERS=1
  run m1 FV_TC3_M=1
  run m2 FV_TC3_M=2
  run pp3 FV_TC3_PP=3
  run pp3m2 FV_TC3_PP=3 FV_TC3_PP_M=2
done > $OUT/r2c_unit_stall.txt 2>&1
tail -100 $OUT/r2c_unit_stall.txt
