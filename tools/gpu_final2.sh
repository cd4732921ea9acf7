#!/bin/bash
TAG=${1:-r01d}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests -q -m gpu --timeout 120 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -1 $OUT/pytest_gpu_$TAG.log | cut -c1-200
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 400 $OUT/bench_$TAG.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:lm_step -s 3 -c 3 -f -o $OUT/lm_$TAG python tools/ncu_lm.py 256 > $OUT/lm_$TAG.log 2>&1
tail -1 $OUT/lm_$TAG.log
