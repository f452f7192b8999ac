"""LBVH build stages (CUDA events inside vkhrt_scene_build) for the C2 groom, a few repetitions."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhrt_b200 as V
strands, segs = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100000, 32)
pos, idx = V.generate_groom(strands, segs, V.GROOM_CURLY)
for tech in (V.PHANTOM, V.LSS, V.DOTS):
    for rep in range(4):
        with V.Scene(pos, idx, technique=tech) as sc:
            sc.build()
            t = sc.timing()
            r = []
            for _ in range(3):
                sc.refit(pos); r.append(sc.timing()["refit_ms"])
            print(tech, rep, sc.n_leaves, {k: round(t[k], 3) for k in ("geometry_ms", "morton_ms", "sort_ms", "hierarchy_ms", "refit_ms", "build_total_ms")}, "refit-only", [round(x, 3) for x in r], flush=True)
