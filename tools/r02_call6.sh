#!/bin/bash
# 2 GPUs: multi-GPU tests (peer/gather/moving camera/shared host frame/render_multi), headless --gpus, bench N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_c6_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/r02_c6_multi_tests.log 2>&1; echo "multi tests rc $?"; tail -5 gpurun_out/r02_c6_multi_tests.log
timeout 600 python -m pytest tests/test_cpp_host.py tests/test_gpu_parity.py -m gpu -x -q -k "cpp or headless or render_multi" > gpurun_out/r02_c6_host.log 2>&1; echo "host rc $?"; tail -3 gpurun_out/r02_c6_host.log
L=vkhrt_b200/_lib
for g in 1 2; do
  $L/vkhrt_headless --model synthetic:curly:100000:32 --technique phantom --size 1920x1080 --frames 8 --gpus $g --hits gpurun_out/r02_c6_hits_g$g.bin 2>&1 | tail -4 | sed "s/^/gpus=$g: /"
done
cmp gpurun_out/r02_c6_hits_g1.bin gpurun_out/r02_c6_hits_g2.bin && echo "headless hits identical 1 vs 2 GPUs"; rm -f gpurun_out/r02_c6_hits_g*.bin
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_c6_bench_n2.json 2> gpurun_out/r02_c6_bench_n2.err; echo "bench n2 rc $?"; tail -3 gpurun_out/r02_c6_bench_n2.err; cut -c1-600 gpurun_out/r02_c6_bench_n2.json
