#!/bin/bash
# round 2, call Y: per-role stall report of the final binary's fused units (HiFi-GAN, B = 8 and B = 32)
OUT=gpurun_out
FV_STALL_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --headline-only --batch 8 2> $OUT/r2y_stall_hifigan_b8.txt > /dev/null
FV_STALL_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --headline-only 2> $OUT/r2y_stall_hifigan_b32.txt > /dev/null
grep -c "tc3" $OUT/r2y_stall_hifigan_b32.txt
