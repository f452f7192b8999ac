#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "per_vertex or lss_per or parity_small or bit_exact or dots" > gpurun_out/r02_c5_taper.log 2>&1; echo "taper rc $?"; tail -15 gpurun_out/r02_c5_taper.log
