#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list, one ncu --set full capture of the
# traversal kernel.  usage (from the repo root, on the GPU box): bash tools/gpu_check.sh [tag] [what...]
#   what: tests bench launches prof (default: all)
tag=${1:-run}; shift
what=${*:-tests bench launches prof}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for w in $what; do
case $w in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" ;;
bench)
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json ;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1; echo "launches rc=$?" ;;
prof)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_pool_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_prof \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_prof.log 2>&1; echo "prof rc=$?" ;;
esac
done
