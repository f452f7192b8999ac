#!/bin/bash
# ncu --set full of the LBVH build / refit kernels on the C2 groom (6.4 M leaves) -> gpurun_out/r02_build.ncu-rep
mkdir -p gpurun_out
cat > build_tmp.py <<'PY'
import numpy as np, vkhrt_b200 as V
pos, idx = V.generate_groom(100000, 32, V.GROOM_CURLY)
with V.Scene(pos, idx) as sc:
    sc.build(); sc.build()
    sc.refit(pos + np.float32(0.001))
    print(sc.timing())
PY
# second build: kernels 13.. (validate, centroid, morton, histogram, scan, 8 x pass, karras, materialise, upper), then the refit pair
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"os_pass_kernel|karras_kernel|os_histogram|materialise_refit|upper_refit|centroid_kernel|morton_kernel" -s 14 -c 16 -f -o gpurun_out/r02_build python build_tmp.py > gpurun_out/r02_build.log 2>&1; echo "build rc=$?"
rm -f build_tmp.py
