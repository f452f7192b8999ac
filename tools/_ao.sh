VKHRT_POOL_MIN_RATIO=0 python -m pytest tests/test_gpu_parity.py tests/test_random_scenes.py -m gpu -x -q -k "occlusion or random_small" 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "occlusion" 2>&1 | tail -2
for e in 1 0 1; do VKHRT_POOL_AO=$e python tools/bench_rows.py 2>/dev/null | python -c "
import json,sys
d={r['row']:r['ms'] for r in json.load(sys.stdin)['rows']}
print('pool_ao=$e ao_4_rays ms', d['ambient_occlusion_4_rays'])"; done
