timeout 900 python tools/gpu_soak.py 1500 777001 extreme 2>&1 | grep -E "SOAK|^cfg" | tail -3
timeout 900 python tools/gpu_soak.py 1500 777002 matrix 2>&1 | grep -E "SOAK|^cfg" | tail -3
timeout 900 python tools/gpu_soak.py 1500 777003 rand 2>&1 | grep -E "SOAK|^cfg" | tail -3
timeout 900 python tools/gpu_soak.py 1000 777004 2>&1 | grep -E "SOAK|^cfg" | tail -3
KSW_B200_TB_WARP_MIN=0 timeout 900 python tools/gpu_soak.py 600 777005 extreme 2>&1 | grep -E "SOAK|^cfg" | tail -3
KSW_B200_PACKED=0 timeout 900 python tools/gpu_soak.py 800 777006 matrix 2>&1 | grep -E "SOAK|^cfg" | tail -3
