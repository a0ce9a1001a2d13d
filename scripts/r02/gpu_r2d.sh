#!/bin/bash
# round 2, call D: correctness of the reworked roles + knob sweep on single ResBlock branches (stall-report totals)
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fused_resblock1 or resblock1_golden or tc_conv or conv_transpose1d or test_conv1d" 2>&1 | tail -6 > $OUT/r2d_pytest_unit.log
cat $OUT/r2d_pytest_unit.log
run() { # label env...
  label=$1; shift
  env "$@" FV_STALL_DEBUG=1 python scripts/unit_bench.py $CC $KK $LL 16 3 2>&1 | grep -A5 "dil=1 " | python scripts/stall_oneline.py "$label"
}
for shape in "16 3 240000" "16 7 240000" "16 11 240000" "32 3 120000" "32 7 120000" "32 11 120000" "64 3 40000" "64 7 40000"; do
  set -- $shape; CC=$1; KK=$2; LL=$3
  echo "=== C=$CC K=$KK L=$LL"
  run base FV_X=0
  run epi1 FV_TC3_EPI=1
  run iss1 FV_TC3_ISSUERS=1
  run iss2 FV_TC3_ISSUERS=2
  run pp3 FV_TC3_PP=3
  run pp3iss1 FV_TC3_PP=3 FV_TC3_ISSUERS=1
  run pp3iss2 FV_TC3_PP=3 FV_TC3_ISSUERS=2
  run pp0 FV_TC3_PP=0
done > $OUT/r2d_sweep.txt 2>&1
cat $OUT/r2d_sweep.txt
