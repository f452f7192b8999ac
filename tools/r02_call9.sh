#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not whole_suite" > gpurun_out/r02_c9_parity.log 2>&1; echo "parity rc $?"; tail -3 gpurun_out/r02_c9_parity.log
VKHRT_POOL_MIN_RATIO=0 VKHRT_NESTED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_c9_parity_pool.log 2>&1; echo "pool-everywhere rc $?"; tail -3 gpurun_out/r02_c9_parity_pool.log
STEPS=40 tools/variants.sh r02_c9 c2 "VKHRT_POOL_CFG=0" "VKHRT_POOL_CFG=1" "VKHRT_LINEWISE=0" 2>&1 | tee gpurun_out/r02_c9_variants.txt
python tools/micro/host_frame_kinds.py 2>&1 | grep -E "ms/frame" | head -3
