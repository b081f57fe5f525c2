#!/bin/bash
# stall-reason breakdown of one launch of kernel $1 (regex) of the v3 step, skipping $2 launches
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-other --api-sample 0 --sustained-s 0"
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-3} -c 1 -o /tmp/st $B > gpurun_out/ncu_st.log 2>&1
python tools/ncu_summary.py /tmp/st.ncu-rep gpurun_out/ncu_st_summary.csv
ncu -i /tmp/st.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[-1]
for k,x in zip(h,v):
    if any(t in k for t in ('issue_stalled','smsp__warps_eligible.avg','smsp__issue_active.avg','inst_executed.sum','sm__warps_active.avg')): print(k,x)
" > gpurun_out/ncu_st_stalls.txt
ncu -i /tmp/st.ncu-rep --page source --csv > gpurun_out/src_st.csv 2>/dev/null
