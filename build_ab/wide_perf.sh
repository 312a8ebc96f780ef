timeout 300 python tools/gpu_perf.py 2000 -1 1500 2>&1 | grep "run 3"
timeout 300 python tools/gpu_perf.py 600 -1 3000 2>&1 | grep "run 3"
timeout 300 python tools/gpu_perf.py 148 -1 6000 2>&1 | grep "run 3"
