"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the pinned oracle).
CPU: the oracle reproduces them bit-for-bit.  GPU: the CUDA path, called through the C ABI, reproduces the
scene-level ones bit-for-bit without the oracle in the loop."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TECHS = ((0, "phantom"), (1, "lss"), (2, "dots"))


@pytest.fixture(scope="module")
def kats():
    return np.load(os.path.join(GOLD, "intersector_kats.npz"))


@pytest.fixture(scope="module")
def scene():
    return np.load(os.path.join(GOLD, "small_scene.npz"))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_golden_has_signal(kats, scene):
    assert (kats["prhi"][:, 0] > 0).sum() > 30 and (kats["prhi"][:, 0] == 0).sum() > 10
    assert (kats["lss"][:, 0] > 0).sum() > 30 and (kats["tri"][:, 0] > 0).sum() > 5
    for _, name in TECHS:
        hits = scene[f"{name}_hits"].reshape(-1).view(np.dtype([("t", "<f4"), ("segment", "<u4"), ("u", "<f4"), ("n", "<f4", 3), ("primitive", "<u4"), ("flags", "<u4")]))
        assert 200 < (hits["flags"] & 1).sum() < hits.shape[0]


def test_oracle_reproduces_intersector_kats(O, kats):
    ro, rd = kats["ray_o"], kats["ray_d"]
    for i in range(ro.shape[0]):
        t, u, n, it = O.prhi(ro[i], rd[i], kats["curves"][i])
        assert np.array_equal(_bits([t, u, *n, it]), _bits(kats["prhi"][i])), f"prhi case {i}"
        hit, t, u, n = O.lss(ro[i], rd[i], kats["lss_in"][i])
        assert np.array_equal(_bits([float(hit), t, u, *n]), _bits(kats["lss"][i])), f"lss case {i}"
        hit, t, u, n = O.tri(ro[i], rd[i], kats["tri_in"][i], i & 1)
        assert np.array_equal(_bits([float(hit), t, u, *n]), _bits(kats["tri"][i])), f"tri case {i}"


def test_oracle_reproduces_small_scene(O, V, scene):
    pos, idx = scene["positions"], scene["indices"]
    W, H = [int(x) for x in scene["size"]]
    vi, pi = scene["view_inverse"], scene["proj_inverse"]
    # the host helpers that produced the inputs are part of the fixture too
    p2, i2 = V.generate_groom(96, 8, V.GROOM_CURLY, 0x5EED0001)
    assert p2.tobytes() == pos.tobytes() and i2.tobytes() == idx.tobytes()
    v2, q2 = V.camera_matrices(position=(0.0, 152.0, 16.0), aspect=float(np.float32(W) / np.float32(H)), fov=50.0)
    assert np.array_equal(_bits(v2), _bits(vi)) and np.array_equal(_bits(q2), _bits(pi))
    for k, (px, py, s) in enumerate(((0, 0, 0), (W - 1, 0, 0), (W // 2, H // 2, 0), (3, 5, 1), (10, 20, 7))):
        assert np.array_equal(_bits(np.concatenate(O.raygen(vi, pi, W, H, px, py, s))), _bits(scene["raygen"][k]))
    for tech, name in TECHS:
        sc = O.OracleScene(pos, idx, technique=tech, radius=0.05)
        nodes, ids, morton, _ = sc.bvh()
        assert np.array_equal(_bits(sc.primitives()), _bits(scene[f"{name}_prims"]))
        assert nodes.tobytes() == scene[f"{name}_nodes"].tobytes()
        assert np.array_equal(ids, scene[f"{name}_ids"]) and np.array_equal(morton, scene[f"{name}_morton"])
        for mode in (0, 1):
            h, img, st = sc.render(O.make_frame(vi, pi, W, H, spp=2, shade_mode=mode, miss_rgb=(0.05, 0.1, 0.2)), stats=True)
            assert h.tobytes() == scene[f"{name}_hits"].tobytes()
            assert np.array_equal(img, scene[f"{name}_rgba{mode}"])
        assert [st["rays"], st["nodes_visited"], st["prims_tested"], st["hits"]] == [int(x) for x in scene[f"{name}_stats"]]


@pytest.mark.gpu
@pytest.mark.parametrize("tech,name", TECHS)
def test_cuda_path_reproduces_small_scene(V, scene, tech, name):
    pos, idx = scene["positions"], scene["indices"]
    W, H = [int(x) for x in scene["size"]]
    vi, pi = scene["view_inverse"], scene["proj_inverse"]
    with V.Scene(pos, idx, technique=tech, radius=0.05) as sc:
        sc.build()
        nodes, ids, morton, _ = sc.bvh()
        assert np.array_equal(_bits(sc.primitives()), _bits(scene[f"{name}_prims"]))
        assert nodes.tobytes() == scene[f"{name}_nodes"].tobytes()
        assert np.array_equal(ids, scene[f"{name}_ids"]) and np.array_equal(morton, scene[f"{name}_morton"])
        for mode in (0, 1):
            h, img, st = sc.render(V.make_frame(vi, pi, W, H, spp=2, shade_mode=mode, miss_rgb=(0.05, 0.1, 0.2)), stats=True)
            assert h.tobytes() == scene[f"{name}_hits"].tobytes()
            assert np.array_equal(img, scene[f"{name}_rgba{mode}"])
        # the scheduler never speculates: a ray's node visits and candidate tests are the oracle's
        assert [st["rays"], st["nodes_visited"], st["prims_tested"], st["hits"]] == [int(x) for x in scene[f"{name}_stats"]]
