#!/bin/bash
# regenerate every kept round-2 measurement with the current build (one GPU): ncu captures, rows, one bench line per BASELINE config
bash tools/r02_profiles.sh
python tools/bench_rows.py > gpurun_out/r02_rows.json 2> gpurun_out/r02_rows.err; echo "rows rc=$?"
EXTRA="" bash tools/workloads.sh r02_bench c1 c3 c4 c5
timeout 1500 python bench.py --steps 200 --warmup 5 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; echo "c2 rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c2_reference.json 2> gpurun_out/r02_bench_c2_reference.err; echo "ref rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c2.json')); print('C2', d['value'], d['e2e']['value'], d['e2e'].get('two_frames_in_flight'), d['roofline']['frac'], d['parity'], d.get('strong_c5',{}).get('value'))
print(open('gpurun_out/r02_bench_c2_reference.json').read()[:300])"
