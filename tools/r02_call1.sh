#!/bin/bash
# round 2, GPU call 1: parity of the new pool kernel + A/B of the pool variants on C2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_c1_gpu.txt 2>&1
nproc >> gpurun_out/r02_c1_gpu.txt; free -g | head -2 >> gpurun_out/r02_c1_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not whole_suite" > gpurun_out/r02_c1_parity.log 2>&1; echo "parity rc $?" 
tail -3 gpurun_out/r02_c1_parity.log
VKHRT_POOL_MIN_RATIO=0 VKHRT_NESTED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_c1_parity_pool.log 2>&1; echo "pool-everywhere rc $?"
tail -3 gpurun_out/r02_c1_parity_pool.log
STEPS=40 tools/variants.sh r02_c1 c2 "VKHRT_POOL_V=1" "VKHRT_POOL_V=2" "VKHRT_POOL_V=2 VKHRT_POOL_CFG=1" "VKHRT_POOL_V=2 VKHRT_POOL_CFG=2" "VKHRT_POOL_V=2 VKHRT_POOL_EXIT=5" "VKHRT_POOL_V=2 VKHRT_POOL_EXIT=7" "VKHRT_POOL_V=2 VKHRT_POOL_NODE_LANES=20" "VKHRT_POOL_V=2 VKHRT_POOL_NODE_LANES=28" 2>&1 | tee gpurun_out/r02_c1_variants.txt
