timeout 300 python -m pytest tests/test_gpu_zhuff.py tests/test_api_e2e.py -x -q -m gpu > gpurun_out/r3_t7.log 2>&1
tail -3 gpurun_out/r3_t7.log
timeout 300 python tools/api_prof.py 2000000 device:device > gpurun_out/r3_apiprof7.log 2>&1
grep -o '"fmt": "[a-z]*"\|"triples_per_s": [0-9]*\|"out_bytes": [^]]*]\|"kernel_ms": {[^}]*}' gpurun_out/r3_apiprof7.log
