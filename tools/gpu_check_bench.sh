#!/bin/bash
# GPU session: full parity suite, bench.py, ncu launch list.  Logs -> gpurun_out/
TAG=${1:-chk}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu --timeout 180 > $OUT/pytest_gpu_$TAG.log 2>&1
tail -4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
timeout 600 python bench.py --steps ${2:-10} --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_$TAG.json').read().strip().splitlines()[-1])
    print('value %.1f pairs/s  %.2f ms/step  e2e %.1f  launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']))
    r=d['roofline']; print('vgg %.2f ms  alg %.0f TF/s frac %.3f issued %.0f'%(r['ms_per_step'],r['achieved'],r['frac'],r['tensor_pipe_tflops_issued']))
    r=d['roofline_lm']; print('lm %.3f ms  %.0f GB/s frac %.3f | B=256:'%(r['ms_per_step'],r['achieved'],r['frac']), r.get('at_batch_256'))
    print(d['clocks']); print(d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e); print(open('$OUT/bench_$TAG.err').read()[-2000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/launches_$TAG.log 2>&1
