( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3_bench_1gpu.json 2> gpurun_out/r3_bench_1gpu.err ) 2> gpurun_out/r3_bench_time.txt
tail -c 600 gpurun_out/r3_bench_1gpu.json; tail -3 gpurun_out/r3_bench_1gpu.err; cat gpurun_out/r3_bench_time.txt
