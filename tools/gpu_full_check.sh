#!/bin/bash
# full GPU parity suite + bench (run under gpurun); logs in gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 600 python bench.py --steps ${1:-5} --warmup 3 ${2:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value %.1f pairs/s  %.2f ms/step  e2e %.1f  launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']))
    r=d['roofline']; print('vgg %.2f ms  alg %.0f TF/s frac %.3f issued %.0f'%(r['ms_per_step'],r['achieved'],r['frac'],r['tensor_pipe_tflops_issued']))
    r=d['roofline_lm']; print('lm %.3f ms  %.0f GB/s frac %.3f'%(r['ms_per_step'],r['achieved'],r['frac']))
    print(d['clocks']); print(d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
