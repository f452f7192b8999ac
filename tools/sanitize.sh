#!/bin/bash
# compute-sanitizer memcheck + racecheck over a small render of every technique (build + trace + shade kernels)
mkdir -p gpurun_out
cat > san_tmp.py <<'PY'
import numpy as np, vkhrt_b200 as V
pos, idx = V.generate_groom(1500, 12, V.GROOM_CURLY)
vi, pi = V.camera_matrices(aspect=float(np.float32(160) / np.float32(96)))
for tech in (V.PHANTOM, V.LSS, V.DOTS):
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        h, img, st = sc.render(V.make_frame(vi, pi, 160, 96, spp=2), stats=False)
        sc.refit(pos + np.float32(0.01))
        h2, _, _ = sc.render(V.make_frame(vi, pi, 160, 96, tile_size=32, tile_first=1, tile_stride=2))
        print(tech, int((h["flags"] & 1).sum()), int((h2["flags"] & 1).sum()))
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python san_tmp.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_$tool.log | tail -3
done
