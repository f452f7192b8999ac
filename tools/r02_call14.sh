#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_at_size.py -m gpu -x -q -k "not whole_suite" > gpurun_out/r02_c14_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r02_c14_tests.log
VKHRT_POOL_MIN_RATIO=0 VKHRT_NESTED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_c14_pool.log 2>&1; echo "pool-everywhere rc $?"; tail -2 gpurun_out/r02_c14_pool.log
for v in "VKHRT_SAMPLE_BATCH=0" "VKHRT_SAMPLE_BATCH=1"; do env $v python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-strong-c5 > gpurun_out/r02_c14_c3.json 2>/dev/null; python tools/variant_line.py "$v c3" gpurun_out/r02_c14_c3.json; done
