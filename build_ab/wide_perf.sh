for v in 1 0; do
echo "PACKED=$v"
KSW_B200_PACKED=$v timeout 300 python tools/gpu_perf.py 64 -1 10000 2>&1 | grep "run 3"
done
