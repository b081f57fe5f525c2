#!/bin/bash
# ncu `--set full` captures of the count / filter / resolve kernels of the other BASELINE configurations (20 M read
# groups each, tools/bench_configs.py): plain run first, then one capture per configuration after its three warm-up steps.
mkdir -p gpurun_out
python tools/bench_configs.py --n 20000000 --steps 1 > gpurun_out/other_plain.jsonl 2> gpurun_out/other_plain.err || exit 1
for c in atac multiome sci3; do
  ncu --set full --clock-control none --import-source on -k regex:'k_count_spec|k_count_plan|k_filter|k_resolve' -s 9 -c 3 -o /tmp/o_$c \
      python tools/bench_configs.py --only $c --n 20000000 --steps 1 > gpurun_out/ncu_o_$c.log 2>&1
  python tools/ncu_summary.py /tmp/o_$c.ncu-rep gpurun_out/ncu_o_$c.csv
done
head -1 gpurun_out/ncu_o_atac.csv > gpurun_out/r02_ncu_other_summary.csv
for c in atac multiome sci3; do tail -n +2 gpurun_out/ncu_o_$c.csv >> gpurun_out/r02_ncu_other_summary.csv; done
ncu -i /tmp/o_sci3.ncu-rep --page source --csv > gpurun_out/src_sci3.csv 2>/dev/null
cat gpurun_out/r02_ncu_other_summary.csv | cut -c1-160
