#!/bin/bash
# ncu --set full of (a) the ambient-occlusion (any-hit, SRC_AO) launch of a C2 frame with 4 AO rays per pixel, (b) the pool kernel with the
# taper terms on the C2 groom with per-vertex radii 0.02 -> 0.005
mkdir -p gpurun_out
cat > extra_tmp.py <<'PY'
import sys, numpy as np, torch, vkhrt_b200 as V
pos, idx = V.generate_groom(100000, 32, V.GROOM_CURLY)
W, H = 1920, 1080
vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
d = torch.empty((W * H, 32), dtype=torch.uint8, device="cuda"); im = torch.empty((W * H, 4), dtype=torch.uint8, device="cuda")
if sys.argv[1] == "ao":
    with V.Scene(pos, idx) as sc:
        sc.build()
        f = V.make_frame(vi, pi, W, H, ao_samples=4, output_memory=V.MEM_DEVICE)
        for k in range(3): sc.render_into(f, d.data_ptr(), im.data_ptr())
        torch.cuda.synchronize()
else:
    rad = np.tile(np.linspace(0.02, 0.005, 33, dtype=np.float32), 100000)
    with V.Scene(pos, idx, radius_per_vertex=rad) as sc:
        sc.build()
        f = V.make_frame(vi, pi, W, H, output_memory=V.MEM_DEVICE)
        for k in range(4): sc.render_into(f, d.data_ptr(), None)
        torch.cuda.synchronize()
PY
# the AO passes of a Phantom frame run in the pool kernel's AO variant (64 slots); VKHRT_POOL_AO=0 gives the lane-bound kernel for comparison
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"trace_pool_kernel.*int.64" -s 2 -c 1 -f -o gpurun_out/r02_ao_pool python extra_tmp.py ao > gpurun_out/r02_ao_pool.log 2>&1; echo "ao pool rc=$?"
VKHRT_POOL_AO=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trace_kernel" -s 2 -c 1 -f -o gpurun_out/r02_ao python extra_tmp.py ao > gpurun_out/r02_ao.log 2>&1; echo "ao rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trace_pool_kernel" -s 2 -c 1 -f -o gpurun_out/r02_taper python extra_tmp.py taper > gpurun_out/r02_taper.log 2>&1; echo "taper rc=$?"
rm -f extra_tmp.py
