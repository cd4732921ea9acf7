#!/bin/bash
# One GPU session (run under gpurun): parity suite, LM variant sweep, bench, ncu launch list.  Logs -> gpurun_out/
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
timeout 700 python -m pytest tests -q -m gpu -x --timeout 120 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_$TAG.log | cut -c1-300
HA_LM_VARIANT=0 timeout 600 python -m pytest tests -q -m gpu -k "lm_ or lazy or full_size or end_to_end" > $OUT/pytest_gpu_v0_$TAG.log 2>&1
tail -3 $OUT/pytest_gpu_v0_$TAG.log | cut -c1-300
timeout 600 python tools/bench_lm.py 32 20 3 0,1,2,3,4,5 > $OUT/bench_lm_b32_$TAG.log 2>&1
timeout 600 python tools/bench_lm.py 256 10 3 0,1,2,3,4,5 > $OUT/bench_lm_b256_$TAG.log 2>&1
grep -h "HA_LM_VARIANT\|level\|whole" $OUT/bench_lm_b32_$TAG.log $OUT/bench_lm_b256_$TAG.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 1500 $OUT/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_$TAG.log 2>&1
ls -la $OUT | tail -12
