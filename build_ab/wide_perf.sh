for v in 1 0; do
echo "PACKED=$v"
KSW_B200_PACKED=$v timeout 300 python tools/gpu_perf.py 600 -1 3000 2>&1 | grep "run 3"
KSW_B200_PACKED=$v timeout 300 python tools/gpu_perf.py 148 -1 6000 2>&1 | grep "run 3"
done
