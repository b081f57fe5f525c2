#!/bin/bash
# ncu `--set full` capture of the device writer's kernels (k_emit_*, k_zh_*) on the second chunk of a make_fastq run
# (scATAC bench input, 1 M read triples = 4 chunks), after a plain run of the same command.
mkdir -p gpurun_out
python tools/writer_prof.py 1000000 2 > gpurun_out/writer_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_emit_plan|k_emit_fill|k_zh_blocks|k_zh_compact' -s 4 -c 4 -o /tmp/writer \
    python tools/writer_prof.py 1000000 1 > gpurun_out/ncu_writer.log 2>&1
python tools/ncu_summary.py /tmp/writer.ncu-rep gpurun_out/r02_ncu_writer_summary.csv
ncu -i /tmp/writer.ncu-rep --page source --csv -k regex:k_zh_blocks > gpurun_out/src_writer.csv 2>/dev/null
cut -c1-200 gpurun_out/r02_ncu_writer_summary.csv
