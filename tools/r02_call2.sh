#!/bin/bash
# round 2, GPU call 2: ncu of the new pool kernel, render_multi test, at-size parity
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_pool2_kernel -s 3 -c 1 -f -o gpurun_out/r02_pool2_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_pool2_prof.log 2>&1; echo "prof rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render_multi or untile or device_outputs" > gpurun_out/r02_c2_multi.log 2>&1; echo "multi rc $?"; tail -3 gpurun_out/r02_c2_multi.log
( time timeout 1500 python -m pytest tests/test_at_size.py -m gpu -x -q -s ) > gpurun_out/r02_c2_atsize.log 2>&1; echo "atsize rc $?"; grep -E "parity at size|passed|failed|real" gpurun_out/r02_c2_atsize.log
