#!/bin/bash
# LM kernel A/B on the GPU box: parity tests for the default variant, then tools/bench_lm.py over the variants.
TAG=${1:-lm}
VARS=${2:-0,1,2,3,4}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu -x --timeout 120 -k "lm_ or lazy or full_size or end_to_end or ford or level4" > $OUT/pytest_lm_$TAG.log 2>&1
tail -3 $OUT/pytest_lm_$TAG.log | cut -c1-300
timeout 600 python tools/bench_lm.py 256 10 3 $VARS > $OUT/bench_lm_b256_$TAG.log 2>&1
timeout 600 python tools/bench_lm.py 32 20 3 $VARS > $OUT/bench_lm_b32_$TAG.log 2>&1
grep -h "HA_LM_VARIANT\|level\|whole" $OUT/bench_lm_b256_$TAG.log $OUT/bench_lm_b32_$TAG.log | cut -c1-60,98-200
