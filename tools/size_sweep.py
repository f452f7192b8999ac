"""device trace time of the C2 groom at several frame sizes (pool vs lane-bound kernel crossover); usage: python tools/size_sweep.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkhrt_b200 as V
pos, idx = V.generate_groom(100000, 32, V.GROOM_CURLY)
sc = V.Scene(pos, idx); sc.build()
for (W, H) in ((512, 512), (960, 540), (1280, 720), (1600, 900), (1920, 1080), (2560, 1440)):
    vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
    d = torch.empty((W * H, 32), dtype=torch.uint8, device="cuda")
    f = V.make_frame(vi, pi, W, H, output_memory=V.MEM_DEVICE)
    ts = []
    for i in range(8):
        sc.render_into(f, d.data_ptr(), None); torch.cuda.synchronize()
        ts.append(sc.timing()["trace_ms"])
    t = min(ts[2:])
    print("%dx%d  %.3f ms  %.1f Mrays/s" % (W, H, t, W * H / t / 1e3))
