#!/bin/bash
# source-level (SASS) counters and the summary row of one k_filter launch of the v3 step
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-other --api-sample 0 --sustained-s 0"
ncu --set full --clock-control none --import-source on -k regex:'k_filter' -s 3 -c 1 -o /tmp/flt $B > gpurun_out/ncu_flt.log 2>&1
ncu -i /tmp/flt.ncu-rep --page source --csv > gpurun_out/src_filter.csv 2>/dev/null
python tools/ncu_summary.py /tmp/flt.ncu-rep gpurun_out/ncu_flt_summary.csv
ncu -i /tmp/flt.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[-1]
for k,x in zip(h,v):
    if any(t in k for t in ('warp_issue_stalled','smsp__average_warp','smsp__warps_eligible','smsp__issue_active','inst_executed.sum','achieved_occupancy','sm__warps_active')): print(k,x)
" > gpurun_out/ncu_flt_stalls.txt
ls -la gpurun_out/src_filter.csv
