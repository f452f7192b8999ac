#!/bin/bash
mkdir -p gpurun_out
STEPS=40 tools/variants.sh r02_c4 c2 "VKHRT_POOL_CFG=0" "VKHRT_POOL_CFG=3" "VKHRT_POOL_CFG=4" "VKHRT_POOL_CFG=5" 2>&1 | tee gpurun_out/r02_c4_variants.txt
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_c4_bench.json 2> gpurun_out/r02_c4_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r02_c4_bench.err; cat gpurun_out/r02_c4_bench.json | cut -c1-3000
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_c4_ref.json 2> gpurun_out/r02_c4_ref.err; tail -3 gpurun_out/r02_c4_ref.err
