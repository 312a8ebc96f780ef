#!/bin/bash
# A/B on the GPU box: tools/gpu_ab.sh <lib A> <lib B> -- device-resident config-2 throughput of two builds, interleaved
for rep in 1 2; do
  for lib in "$@"; do
    echo "== $lib"; SEDEF_B200_LIB=$lib python tools/gpu_perf.py 100000 100 1000 2>&1 | grep "^run [23]"
  done
done
