#!/bin/bash
# LM kernel: parity tests + per-level ncu metrics (run under gpurun)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "lm_ or full_size or lazy" > gpurun_out/pytest_lm.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:lm_step --csv python tools/bench_lm.py ${1:-32} 1 ${2:-3} 2>&1 | grep -E "lm_step" | python -c "
import sys,csv,collections
rows=list(csv.reader(sys.stdin))
d=collections.OrderedDict()
for r in rows:
    d.setdefault((r[0],r[4][:36]),{})[r[12][:28]]=r[14]
last={}
for k,v in d.items(): last[k[1]]=v
for k,v in last.items(): print(k, v)
"
tail -2 gpurun_out/pytest_lm.log | cut -c1-200
