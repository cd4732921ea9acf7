#!/bin/bash
# Round-end GPU session: parity suite, smoke, bench, ncu launch list, full captures of the conv and LM kernels.
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -q -m gpu --timeout 120 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -2 $OUT/pytest_gpu_$TAG.log | cut -c1-200
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 600 $OUT/bench_$TAG.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 40 -c 10 -f -o $OUT/conv_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/conv_$TAG.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lm_step -s 3 -c 3 -f -o $OUT/lm_$TAG python tools/ncu_lm.py 256 > $OUT/lm_$TAG.log 2>&1
ls -la $OUT | tail -8
