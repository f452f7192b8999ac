#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pinned or line_wise or zero_copy" > gpurun_out/r02_c12_lines.log 2>&1; echo "lines rc $?"; tail -2 gpurun_out/r02_c12_lines.log
STEPS=60 tools/lib_variants.sh r02_c12 "c2" var_lp0 var_lp1 var_lp0 var_lp1 2>&1 | tee gpurun_out/r02_c12_variants.txt
