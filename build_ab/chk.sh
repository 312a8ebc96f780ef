echo "=== BUGGY build (stale slot bases in the key pass): the new cases must FAIL here"
SEDEF_B200_LIB=build_ab/lib_bug.so timeout 600 python tools/gpu_debug.py 23 24 2>&1 | grep -E "^config|TOTAL"
SEDEF_B200_LIB=build_ab/lib_bug.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "leaves_the_band" 2>&1 | tail -2
echo "=== FIXED build"
timeout 800 python tools/gpu_debug.py 2>&1 | grep -E "TOTAL|mismatches [1-9]"
timeout 300 python tools/repro_fields.py 2>&1 | grep "^cfg"
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
