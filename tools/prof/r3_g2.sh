timeout 400 python -m pytest tests/test_gpu_zhuff.py tests/test_api_e2e.py tests/test_gpu_textparse.py -x -q -m gpu > gpurun_out/r3_t2.log 2>&1
tail -5 gpurun_out/r3_t2.log
timeout 300 python tools/api_prof.py 2000000 > gpurun_out/r3_apiprof2.log 2>&1
tail -12 gpurun_out/r3_apiprof2.log
