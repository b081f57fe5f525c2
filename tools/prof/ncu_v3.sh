#!/bin/bash
# ncu captures of the v3 step (run on the GPU box through gpurun): plain run first, then the launch list, then one
# `--set full` capture of the four kernels of a timed step.  Reports go to /tmp (large); CSV summaries to gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-other --api-sample 0 --sustained-s 0"
$B > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_count_spec|k_filter|k_count_lenonly|k_resolve' -s 12 -c 4 -o /tmp/v3 $B > gpurun_out/ncu_v3.log 2>&1
python tools/ncu_summary.py /tmp/v3.ncu-rep gpurun_out/r02_ncu_v3_summary.csv
tail -3 gpurun_out/ncu_v3.log | cut -c1-200
