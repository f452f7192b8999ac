#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not whole_suite" > gpurun_out/r02_c7_parity.log 2>&1; echo "parity rc $?"; tail -3 gpurun_out/r02_c7_parity.log
timeout 900 python -m pytest tests/test_at_size.py tests/test_assets_and_lod.py -m gpu -x -q > gpurun_out/r02_c7_atsize.log 2>&1; echo "atsize+lod rc $?"; tail -3 gpurun_out/r02_c7_atsize.log
python tools/build_breakdown.py 2>&1 | tail -12
python tools/build_breakdown.py 1000000 64 2>&1 | grep "^0" | tail -4
