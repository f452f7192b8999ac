#!/bin/bash
# all BASELINE.json configs on one GPU (short runs), one bench line each -> gpurun_out/<tag>_<workload>.json
tag=$1; shift
mkdir -p gpurun_out
for w in ${*:-c1 c3 c4 c5}; do
  steps=10; [ $w = c5 ] && steps=2
  timeout 900 python bench.py --workload $w --steps $steps --warmup 3 --no-strong-c5 $EXTRA > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err
  echo "$w rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/${tag}_$w.json')); r=d['roofline']
print('  %.1f Mrays/s  %.3f ms/step  e2e %.1f  frac %.3f  B_ray %.0f  N_int %.1f N_prim %.2f build_ms %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['bytes_per_ray'], r['n_int_per_ray'], r['n_prim_per_ray'], d['config']['build_ms']))
print('  sched', {k:(v['steps'], round(v['lanes_per_step'],1)) for k,v in r['warp_scheduler_rank0'].items()})
" 2>&1 | tail -3
done
