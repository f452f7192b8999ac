#!/bin/bash
# usage: bash tools/sweep.sh tag "VAR=val VAR2=val" "VAR=val" ...   -> one bench line per setting
tag=$1; shift
mkdir -p gpurun_out
: > gpurun_out/${tag}_sweep.txt
for cfg in "$@"; do
  out=$(env $cfg python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; s=r['warp_scheduler_rank0']; print('%.1f Mrays/s  kernel_ms %.3f  nodes/ray %.2f prims/ray %.2f iters/ray %.2f | ' % (d['value'], r['kernel_ms'], r['n_int_per_ray'], r['n_prim_per_ray'], r['phantom_iterations_per_ray']) + ' '.join('%s %.2fM x %.1f' % (k, v['steps']/1e6, v['lanes_per_step']) for k,v in s.items()))")
  echo "$cfg => $out" | tee -a gpurun_out/${tag}_sweep.txt
done
