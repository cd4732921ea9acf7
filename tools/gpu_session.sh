#!/bin/bash
# One GPU-box session (run under gpurun): tools/gpu_session.sh TAG stage [stage ...]
# Every stage has its own `timeout`; logs and raw ncu reports go to gpurun_out/ (scratch), summaries are made
# afterwards with tools/summarise_profiles.py TAG.
#   tests [-k expr]   parity suite (pytest -m gpu)          smoke      __graft_entry__.smoke()
#   bench [config]    bench.py (default kitti32)            launches   ncu launch list of a short bench run
#   ncu_conv          ncu --set full of the tcgen05 convs   ncu_lm     ncu --set full of the LM step (B=256)
#   lm_ab             tools/bench_lm.py at B=256 and B=32   sanitize   compute-sanitizer memcheck on smoke()
#   ncu_lm_stress / ncu_conv_stress   BASELINE config 5 captures    convab     per-layer CTA-pair vs single-CTA conv table
#   pipeline          input pipeline (f-4) bench + launch list
TAG=$1; shift
OUT=gpurun_out
REP=/tmp/ha_ncu          # raw .ncu-rep files stay on the box (gpurun copies back at most 64 MiB): their raw pages come back as CSV
mkdir -p $OUT $REP
export_rep() {           # export_rep <name>: raw metrics page (+ SASS source page of the first launch) of $REP/<name>.ncu-rep
  ncu -i $REP/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  for k in ${2:-0}; do   # SASS source page (per-instruction samples / executed counts) of the listed launches
    ncu -i $REP/$1.ncu-rep --page source --csv --print-source sass --launch-skip $k --launch-count 1 2>/dev/null | awk 'NR==1||!/^"Kernel Name"/||!s++' > $OUT/$1_sass$k.csv
  done
  ls -la $REP/$1.ncu-rep $OUT/$1_raw.csv | cut -c20-120
}
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
      timeout 1200 python -m pytest tests -q -m gpu -s --timeout 240 ${K:+-k "$K"} > $OUT/pytest_gpu_$TAG.log 2>&1
      grep -E "^E  |^FAILED|^ERROR|passed|failed" $OUT/pytest_gpu_$TAG.log | cut -c1-300 | tail -25
      # the achieved deviations the parity tests print (DESIGN.md section 4 quotes them)
      grep -E "final \|d\| per pair|planted B=32|level [0-9]: max\|d\|" $OUT/pytest_gpu_$TAG.log | cut -c1-400 > $OUT/parity_$TAG.txt ;;
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
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 40 -c 10 -f -o $REP/conv_$TAG \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/conv_$TAG.log 2>&1
      export_rep conv_$TAG "0 7 8" ;;
    ncu_lm)
      timeout 400 ncu --set full --clock-control none --import-source on -k regex:lm_step -s 3 -c 3 -f -o $REP/lm_$TAG \
        python tools/ncu_lm.py 256 > $OUT/lm_$TAG.log 2>&1
      export_rep lm_$TAG 2 ;;
    ncu_lm_stress)   # BASELINE config 5: 4-level pyramid, 512 pairs per GPU
      timeout 600 ncu --set full --clock-control none -k regex:lm_step -s 4 -c 4 -f -o $REP/lmstress_$TAG \
        python tools/ncu_lm.py 512 0 4 > $OUT/lmstress_$TAG.log 2>&1
      export_rep lmstress_$TAG 3 ;;
    ncu_conv_stress) # the level-4 U-Net (adds conv_dec3) at a batch ncu can replay: tensor-pipe metrics per conv launch
      timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum \
        --clock-control none -k regex:conv -s 52 -c 26 --csv --log-file $OUT/convstress_$TAG.csv \
        python bench.py --config stress --batch 32 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $OUT/convstress_$TAG.log 2>&1
      tail -1 $OUT/convstress_$TAG.log | cut -c1-200 ;;
    convab)
      timeout 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,lts__t_bytes.sum \
        --clock-control none -k regex:conv3x3 --csv --log-file $OUT/convab_$TAG.csv python tools/bench_conv.py 32 1 > $OUT/convab_$TAG.log 2>&1
      python tools/convab_report.py $OUT/convab_$TAG.csv ;;
    lm_ab)
      timeout 400 python tools/bench_lm.py 256 10 3 0,2 > $OUT/bench_lm_b256_$TAG.log 2>&1
      timeout 400 python tools/bench_lm.py 32 20 3 0,2 > $OUT/bench_lm_b32_$TAG.log 2>&1
      timeout 300 python tools/lm_profile_batches.py 32 64 128 256 512 > $OUT/lm_batches_$TAG.txt 2>&1; cat $OUT/lm_batches_$TAG.txt
      grep -h "variant\|level\|whole" $OUT/bench_lm_b256_$TAG.log $OUT/bench_lm_b32_$TAG.log | cut -c1-200 ;;
    pipeline)        # input pipeline on the GPU vs PIL on the host, plus its ncu launch list
      timeout 300 python tools/bench_input_pipeline.py 32 20 8 > $OUT/pipeline_$TAG.json 2> $OUT/pipeline_$TAG.err
      cat $OUT/pipeline_$TAG.json | cut -c1-900
      timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:img_ -s 24 -c 8 --csv \
        --log-file $OUT/pipeline_launches_$TAG.csv python tools/bench_input_pipeline.py 32 2 1 > /dev/null 2>&1
      grep -v "^==" $OUT/pipeline_launches_$TAG.csv | cut -d, -f5,12- | head -30 ;;
    train)           # native train step vs the torch / cuDNN U-Net path; launch list of one native step
      timeout 400 python tools/bench_train.py 3 5 > $OUT/train_$TAG.json 2> $OUT/train_$TAG.err; cat $OUT/train_$TAG.json | cut -c1-1200
      timeout 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none -c 3000 --csv --log-file $OUT/train_launches_$TAG.csv python tools/bench_train.py 3 1 native > /dev/null 2>&1
      python tools/train_launch_report.py $OUT/train_launches_$TAG.csv > $OUT/train_launches_$TAG.md; head -45 $OUT/train_launches_$TAG.md ;;
    sanitize)
      timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > $OUT/memcheck_smoke_$TAG.log 2>&1
      tail -3 $OUT/memcheck_smoke_$TAG.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
