#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu captures of the two hot kernels.
# Outputs land in gpurun_out/ (scratch); summaries are copied to profiles/ by tools/summarise_profiles.py.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 24 -c 4 -f -o $OUT/conv_$TAG $BENCH > $OUT/conv_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lm_step -s 15 -c 3 -f -o $OUT/lm_$TAG $BENCH > $OUT/lm_$TAG.log 2>&1
ls -la $OUT
