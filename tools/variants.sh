#!/bin/bash
# bench.py under several environment variants (one line each): tools/variants.sh tag workload "VAR=1 VAR2=3" "..." ...
tag=$1; wl=$2; shift 2
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --workload $wl --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline --no-parity --no-strong-c5 > gpurun_out/${tag}_v$i.json 2> gpurun_out/${tag}_v$i.err
  python tools/variant_line.py "$v" gpurun_out/${tag}_v$i.json
done
