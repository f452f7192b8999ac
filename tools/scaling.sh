#!/bin/bash
# bench.py at N = 1,2,4,8 on one box (weak scaling) -> gpurun_out/<tag>_scale.txt
tag=${1:-scale}; mkdir -p gpurun_out; : > gpurun_out/${tag}_scale.txt
for n in ${2:-1 2 4 8}; do
  if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py"; fi
  timeout 600 $cmd --gpus $n --steps 30 --warmup 5 --no-cpu-baseline 2> gpurun_out/${tag}_n$n.err | tail -1 > gpurun_out/${tag}_n$n.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_n$n.json'))
print('N=%d value %.1f Mrays/s  ms/step %.3f  e2e %.1f  frame %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['workload'].split(',')[2]))" | tee -a gpurun_out/${tag}_scale.txt
done
