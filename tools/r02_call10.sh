#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c10_tests.log 2>&1; echo "tests rc $?"; tail -4 gpurun_out/r02_c10_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c10_bench.json 2> gpurun_out/r02_c10_bench.err; echo "bench rc $?"; cut -c1-300 gpurun_out/r02_c10_bench.json
bash tools/r02_profiles.sh
