#!/bin/bash
mkdir -p gpurun_out
python tools/bench_rows.py > gpurun_out/r02_rows.json 2> gpurun_out/r02_rows.err; echo "rows rc $?"; grep -E '"row"|"ms"|frac' gpurun_out/r02_rows.json | paste - - - | head -20
EXTRA="" bash tools/workloads.sh r02_bench c1 c3 c4 c5
python bench.py --steps 200 --warmup 5 > gpurun_out/r02_bench_c2_200.json 2> gpurun_out/r02_bench_c2_200.err; echo "c2 rc $?"; python tools/variant_line.py c2 gpurun_out/r02_bench_c2_200.json
