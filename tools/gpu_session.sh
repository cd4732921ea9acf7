#!/bin/bash
# One GPU-box session (run under gpurun): tools/gpu_session.sh TAG stage [stage ...]
# Every stage has its own `timeout`; logs and raw ncu reports go to gpurun_out/ (scratch), summaries are made
# afterwards with tools/summarise_profiles.py TAG.
#   tests [-k expr]   parity suite (pytest -m gpu)          smoke      __graft_entry__.smoke()
#   bench [config]    bench.py (default kitti32)            launches   ncu launch list of a short bench run
#   ncu_conv          ncu --set full of the tcgen05 convs   ncu_lm     ncu --set full of the LM step (B=256)
#   lm_ab             tools/bench_lm.py at B=256 and B=32   sanitize   compute-sanitizer memcheck on smoke()
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
summarise_bench() {
python - "$1" <<'PY'
import json, sys
p = sys.argv[1]
try:
    d = json.loads(open(p).read().strip().splitlines()[-1])
    print('%s: value %.1f pairs/s  %.2f ms/step  e2e %.1f  launches %d' % (d['config']['name'], d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
    r = d['roofline']; print('  vgg %.2f ms  alg %.0f TF/s frac %.3f issued %.0f' % (r['ms_per_step'], r['achieved'], r['frac'], r['tensor_pipe_tflops_issued']))
    r = d['roofline_lm']; print('  lm %.3f ms  %.0f GB/s frac %.3f | B=256: %.3f ms frac %.3f' % (r['ms_per_step'], r['achieved'], r['frac'], r['at_batch_256']['ms_per_step'], r['at_batch_256']['frac']))
    print('  clocks', d['clocks']); print('  cpu', d.get('cpu_baseline')); print('  delta', d.get('pose_delta_vs_ref')); print('  info', d.get('informational'))
except Exception as e:
    print('bench parse failed', e); print(open(p.replace('.json', '.err')).read()[-1500:])
PY
}
while [ $# -gt 0 ]; do
  stage=$1; shift
  case $stage in
    tests)
      K=""; if [ "$1" == "-k" ]; then K="$2"; shift 2; fi
      timeout 1200 python -m pytest tests -q -m gpu --timeout 240 ${K:+-k "$K"} > $OUT/pytest_gpu_$TAG.log 2>&1
      grep -E "^E  |^FAILED|^ERROR|passed|failed" $OUT/pytest_gpu_$TAG.log | cut -c1-300 | tail -25 ;;
    smoke)
      timeout 180 python __graft_entry__.py smoke 2>&1 | tail -2 ;;
    bench)
      CFG=kitti32; case "$1" in kitti32|ford64|kitti1024x8|stress) CFG=$1; shift ;; esac
      timeout 900 python bench.py --config $CFG --steps 10 --warmup 3 > $OUT/bench_${CFG}_$TAG.json 2> $OUT/bench_${CFG}_$TAG.err
      summarise_bench $OUT/bench_${CFG}_$TAG.json ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_$TAG.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $OUT/launches_$TAG.log 2>&1
      tail -1 $OUT/launches_$TAG.log | cut -c1-200 ;;
    ncu_conv)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 40 -c 10 -f -o $OUT/conv_$TAG \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/conv_$TAG.log 2>&1
      tail -1 $OUT/conv_$TAG.log | cut -c1-200 ;;
    ncu_lm)
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:lm_step -s 3 -c 3 -f -o $OUT/lm_$TAG \
        python tools/ncu_lm.py 256 > $OUT/lm_$TAG.log 2>&1
      tail -1 $OUT/lm_$TAG.log | cut -c1-200 ;;
    lm_ab)
      timeout 400 python tools/bench_lm.py 256 10 3 0,1 > $OUT/bench_lm_b256_$TAG.log 2>&1
      timeout 400 python tools/bench_lm.py 32 20 3 0 > $OUT/bench_lm_b32_$TAG.log 2>&1
      grep -h "variant\|level\|whole" $OUT/bench_lm_b256_$TAG.log $OUT/bench_lm_b32_$TAG.log | cut -c1-60,98-200 ;;
    sanitize)
      timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > $OUT/memcheck_smoke_$TAG.log 2>&1
      tail -3 $OUT/memcheck_smoke_$TAG.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
