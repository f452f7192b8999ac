#!/bin/bash
# round 2 profiles: launch list of the default bench command, ncu --set full of the dominant kernel per workload
mkdir -p gpurun_out
B="--no-cpu-baseline --no-parity --no-strong-c5"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 $B > gpurun_out/r02_launches.log 2>&1; echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_pool_kernel -s 3 -c 1 -f -o gpurun_out/r02_c2_pool \
    python bench.py --steps 2 --warmup 3 $B > gpurun_out/r02_c2_pool.log 2>&1; echo "c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trace_pool_kernel -s 70 -c 1 -f -o gpurun_out/r02_c5_pool \
    python bench.py --workload c5 --steps 1 --warmup 3 $B > gpurun_out/r02_c5_pool.log 2>&1; echo "c5 rc=$?"
# one launch per sample (VKHRT_SAMPLE_BATCH=0), so that the captured launch is the one the bench line's kernel_ms times (sample 0: 2.07 M rays)
VKHRT_SAMPLE_BATCH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 10 -c 1 -f -o gpurun_out/r02_c3_lss \
    python bench.py --workload c3 --steps 2 --warmup 3 $B > gpurun_out/r02_c3_lss.log 2>&1; echo "c3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/r02_c4_dots \
    python bench.py --workload c4 --steps 2 --warmup 3 $B > gpurun_out/r02_c4_dots.log 2>&1; echo "c4 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/r02_c1_lane \
    python bench.py --workload c1 --steps 2 --warmup 3 $B > gpurun_out/r02_c1_lane.log 2>&1; echo "c1 rc=$?"
ls -la gpurun_out/*.ncu-rep
