timeout 600 python -m pytest tests/ -x -q -m gpu > gpurun_out/r3_t4.log 2>&1
tail -5 gpurun_out/r3_t4.log
timeout 200 python tools/api_prof.py 2000000 device:device > gpurun_out/r3_apiprof4.log 2>&1
tail -5 gpurun_out/r3_apiprof4.log
