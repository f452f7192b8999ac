#!/bin/bash
mkdir -p gpurun_out
n=${1:-4}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+n)) bench.py --gpus $n --steps 20 --warmup 5 ) > gpurun_out/r02_c15_bench_n$n.json 2> gpurun_out/r02_c15_bench_n$n.err
echo "bench n$n rc $?"; grep real gpurun_out/r02_c15_bench_n$n.err; grep "^{" gpurun_out/r02_c15_bench_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d.get('strong_c5') or {}
print('N=%d value %.1f e2e %.1f (pcie/rank %.1f GB/s) frac %.3f parity %s assembled %s' % (d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['pcie_gbs_per_rank'], d['roofline']['frac'], d.get('parity',{}).get('bit_identical'), (d.get('parity_assembled') or {}).get('hits_identical')))
print('   strong_c5 tile %s: %.1f Mrays/s  ms/frame %.2f  eff %.3f  own shard ms %s  e2e %.1f  assembled %s' % (c.get('tile'), c.get('value',0), c.get('ms_per_frame',0), c.get('efficiency_vs_single_gpu_same_run',0), c.get('own_shard_ms_min_max_over_ranks'), c.get('e2e',{}).get('value',0), {k:v for k,v in (c.get('parity_assembled') or {}).items() if 'identical' in k}))
"
