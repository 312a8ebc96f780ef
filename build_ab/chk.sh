echo "== packed rand 31337"; timeout 1500 python tools/gpu_soak.py 1200 31337 rand 2>&1 | grep -E "SOAK|^cfg" | tail -3
echo "== one-slot rand 31337"; KSW_B200_PACKED=0 timeout 1500 python tools/gpu_soak.py 600 31337 rand 2>&1 | grep -E "SOAK|^cfg" | tail -3
echo "== packed matrix"; timeout 900 python tools/gpu_soak.py 500 1234567 matrix 2>&1 | grep -E "SOAK|^cfg" | tail -3
echo "== suites"; timeout 800 python tools/gpu_debug.py 2>&1 | grep -E "TOTAL|mismatches [1-9]"; python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "== throughput"; timeout 300 python tools/gpu_perf.py 100000 100 1000 2>&1 | grep -E "run 3"
