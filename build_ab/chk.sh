echo "== build WITHOUT the clamp: does the new test fail?"
SEDEF_B200_LIB=build_ab/lib_bug.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "degenerate" 2>&1 | tail -3
echo "== search for a reproducer on that build"
SEDEF_B200_LIB=build_ab/lib_bug.so timeout 600 python tools/find_tie_repro.py 2>&1 | tail -8
