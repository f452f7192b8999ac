#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/vkhrt_oracle.cpp).

The reference (mmzala/vkhrt) holds no golden vectors and cannot run here, so these
fixtures are produced by the oracle AFTER it passed its pins (tests/test_oracle_pins.py: survey KATs,
the literal numpy GLSL transcription, analytic cylinder/capsule/triangle answers).  They freeze that
state: tests/test_golden.py checks the oracle against them bit-for-bit on CPU and the CUDA path against
them on the GPU box (where neither /root/reference nor a second opinion is available).

    python tests/golden/make_golden.py      # rewrites the fixtures; commit the result
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import vkhrt_b200 as V                      # host helpers only (groom generator, camera): no GPU needed
from oracle import oracle as O

GROOM = dict(n_strands=96, segments=8, style=1, seed=0x5EED0001)
FRAME = dict(width=64, height=40)


def kat_inputs():
    rng = np.random.default_rng(0xC0FFEE)
    n = 96
    rays_o, rays_d, curves = [], [], []
    for _ in range(n):
        p0 = rng.normal(0, 1, 3) + (0, 150, 0)
        d = rng.normal(0, 1, 3)
        d /= np.linalg.norm(d)
        L = rng.uniform(0.1, 1.0)
        cv = np.array([p0, p0 + d * L / 3 + rng.normal(0, 0.02, 3), p0 + d * 2 * L / 3 + rng.normal(0, 0.02, 3), p0 + d * L], np.float32)
        tgt = 0.5 * (cv[0] + cv[3]) + rng.normal(0, 0.02, 3)
        ro = (tgt + rng.normal(0, 1, 3) * 8).astype(np.float32)
        rd = tgt - ro
        rays_o.append(ro)
        rays_d.append((rd / np.linalg.norm(rd)).astype(np.float32))
        curves.append(cv.reshape(12))
    return np.array(rays_o, np.float32), np.array(rays_d, np.float32), np.array(curves, np.float32)


def main():
    ro, rd, cv = kat_inputs()
    prhi = np.array([[*(lambda r: (r[0], r[1], *r[2], r[3]))(O.prhi(ro[i], rd[i], cv[i]))] for i in range(len(ro))], np.float32)
    lss_in = np.concatenate([cv[:, 0:3], np.full((len(cv), 1), 0.03, np.float32), cv[:, 9:12], np.full((len(cv), 1), 0.012, np.float32)], axis=1).astype(np.float32)
    lss = np.array([[*(lambda r: (float(r[0]), r[1], r[2], *r[3]))(O.lss(ro[i], rd[i], lss_in[i]))] for i in range(len(ro))], np.float32)
    tri_in = np.concatenate([cv[:, 0:3], cv[:, 9:12], cv[:, 0:3] + np.float32([0.0, 0.3, 0.1])], axis=1).astype(np.float32)
    tri = np.array([[*(lambda r: (float(r[0]), r[1], r[2], *r[3]))(O.tri(ro[i], rd[i], tri_in[i], i & 1))] for i in range(len(ro))], np.float32)
    np.savez_compressed(os.path.join(HERE, "intersector_kats.npz"), ray_o=ro, ray_d=rd, curves=cv, prhi=prhi, lss_in=lss_in, lss=lss, tri_in=tri_in, tri=tri)

    pos, idx = V.generate_groom(GROOM["n_strands"], GROOM["segments"], GROOM["style"], GROOM["seed"])
    W, H = FRAME["width"], FRAME["height"]
    vi, pi = V.camera_matrices(position=(0.0, 152.0, 16.0), aspect=float(np.float32(W) / np.float32(H)), fov=50.0)
    out = dict(positions=pos, indices=idx, view_inverse=vi, proj_inverse=pi, size=np.array([W, H]))
    rays = np.array([np.concatenate(O.raygen(vi, pi, W, H, px, py, s)) for (px, py, s) in ((0, 0, 0), (W - 1, 0, 0), (W // 2, H // 2, 0), (3, 5, 1), (10, 20, 7))], np.float32)
    out["raygen"] = rays
    for tech, name in ((0, "phantom"), (1, "lss"), (2, "dots")):
        sc = O.OracleScene(pos, idx, technique=tech, radius=0.05)
        nodes, ids, morton, lohi = sc.bvh()
        out[f"{name}_prims"] = sc.primitives()
        out[f"{name}_nodes"] = nodes.view(np.uint8).reshape(-1, 64)
        out[f"{name}_ids"] = ids
        out[f"{name}_morton"] = morton
        for mode in (0, 1):
            h, img, st = sc.render(O.make_frame(vi, pi, W, H, spp=2, shade_mode=mode, miss_rgb=(0.05, 0.1, 0.2)), stats=True)
            out[f"{name}_hits"] = h.view(np.uint8).reshape(-1, 32)
            out[f"{name}_rgba{mode}"] = img
        out[f"{name}_stats"] = np.array([st["rays"], st["nodes_visited"], st["prims_tested"], st["hits"]], np.uint64)
    np.savez_compressed(os.path.join(HERE, "small_scene.npz"), **out)
    for f in ("intersector_kats.npz", "small_scene.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
