echo "== packed, rand scoring, seed 31337 (was 118, then 6)"; timeout 1500 python tools/gpu_soak.py 1200 31337 rand 2>&1 | grep -E "SOAK|^cfg" | tail -4
echo "== one-slot, rand scoring, seed 31337 (was 6)"; KSW_B200_PACKED=0 timeout 1500 python tools/gpu_soak.py 1200 31337 rand 2>&1 | grep -E "SOAK|^cfg" | tail -4
echo "== packed, rand scoring, fresh seeds"; timeout 1500 python tools/gpu_soak.py 1500 424242 rand 2>&1 | grep -E "SOAK|^cfg" | tail -4
echo "== packed, SEDEF scoring"; timeout 900 python tools/gpu_soak.py 600 5150 2>&1 | grep -E "SOAK|^cfg" | tail -3
echo "== differential suite"; timeout 800 python tools/gpu_debug.py 2>&1 | grep -E "TOTAL|mismatches [1-9]"
echo "== test suite"; python -m pytest tests -x -q -m gpu 2>&1 | tail -2
echo "== throughput"; timeout 300 python tools/gpu_perf.py 100000 100 1000 2>&1 | grep -E "run 3"
