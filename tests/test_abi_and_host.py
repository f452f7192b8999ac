"""CPU tests of the drop-in boundary and the host logic: the C-ABI library loads and exports every symbol
include/vkhrt_b200.h declares, struct layouts match the header, compute entry points fail loudly without a
GPU (no CPU fallback), and the host helpers (FlyCamera matrices, synthetic groom) are correct.
No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vkhrt_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vkhrt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(V):
    names = declared_functions()
    assert len(names) >= 20
    lib = C.CDLL(V.library_path())
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/vkhrt_b200.h but not exported"
    # and the Python binding lists exactly the same set
    from vkhrt_b200.api import ABI_SYMBOLS
    assert sorted(ABI_SYMBOLS) == names
    # the dynamic symbol table agrees (no accidental C++ mangling of the ABI)
    out = subprocess.run(["nm", "-D", "--defined-only", V.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(vkhrt_[a-z0-9_]+)$", out, flags=re.M))
    assert set(names) <= exported


def test_product_does_not_link_the_oracle(V):
    out = subprocess.run(["ldd", V.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out
    src_dir = os.path.join(ROOT, "vkhrt_b200")
    for dp, _, fs in os.walk(src_dir):
        for f in fs:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp", ".py")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle/" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dp, f)


def test_struct_layouts_match_the_header(V):
    from vkhrt_b200 import api
    assert V.HIT_DTYPE.itemsize == 32 and V.NODE_DTYPE.itemsize == 64
    # compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors
    prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "vkhrt_b200.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(VkhrtSceneDesc), sizeof(VkhrtFrameDesc), sizeof(VkhrtHit), sizeof(VkhrtBvhNode),
        sizeof(VkhrtBvhView), sizeof(VkhrtTiming), sizeof(VkhrtTraceStats));
 printf("%zu %zu %zu %zu %zu %zu\n", offsetof(VkhrtFrameDesc, proj_inverse), offsetof(VkhrtFrameDesc, width), offsetof(VkhrtFrameDesc, spp),
        offsetof(VkhrtFrameDesc, miss_rgb), offsetof(VkhrtFrameDesc, tile_size), offsetof(VkhrtFrameDesc, stream));
 printf("%zu %zu\n", offsetof(VkhrtFrameDesc, ao_samples), offsetof(VkhrtFrameDesc, ao_bias));
 printf("%zu %zu %zu\n", offsetof(VkhrtSceneDesc, line_indices), offsetof(VkhrtSceneDesc, radius), offsetof(VkhrtSceneDesc, device));
 return 0; }
"""
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True).stdout.split("\n")
    sizes = [int(x) for x in out[0].split()]
    assert sizes == [C.sizeof(api.SceneDesc), C.sizeof(api.FrameDesc), 32, 64, C.sizeof(api.BvhView), C.sizeof(api.Timing), C.sizeof(api.TraceStats)]
    F = api.FrameDesc
    assert [int(x) for x in out[1].split()] == [F.proj_inverse.offset, F.width.offset, F.spp.offset, F.miss_rgb.offset, F.tile_size.offset, F.stream.offset]
    assert [int(x) for x in out[2].split()] == [F.ao_samples.offset, F.ao_bias.offset]
    S = api.SceneDesc
    assert [int(x) for x in out[3].split()] == [S.line_indices.offset, S.radius.offset, S.device.offset]
    from oracle import oracle as O
    assert C.sizeof(O.FrameDesc) == C.sizeof(F) and O.FrameDesc.ao_samples.offset == F.ao_samples.offset


def test_header_is_plain_c(V):
    # the boundary must be bindable from C (cgo/JNI/ctypes style): the header compiles as C89-ish C with no C++
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER])


def test_error_strings_and_versions(V):
    L = V.lib()
    assert L.vkhrt_abi_version() == 6
    assert L.vkhrt_error_string(0) == b"ok"
    for code in range(-8, 0):
        assert L.vkhrt_error_string(code) not in (b"ok", b"unknown status")
    assert L.vkhrt_error_string(-99) == b"unknown status"


def test_no_gpu_means_loud_failure_not_a_cpu_fallback(V):
    if V.device_count() > 0:
        pytest.skip("a GPU is present")
    pos, idx = V.generate_groom(4, 2, V.GROOM_STRAIGHT)
    with pytest.raises(V.VkhrtError) as e:
        V.Scene(pos, idx)
    assert e.value.status == -2 and "no CPU path" in str(e.value)       # VKHRT_ERR_NO_DEVICE


def test_argument_validation_without_gpu(V):
    L = V.lib()
    from vkhrt_b200 import api
    h = C.c_void_p()
    assert L.vkhrt_scene_create(None, C.byref(h)) == -1
    d = api.SceneDesc(None, 0, None, 0, None, 0.02, 7, 0)                 # unknown technique
    assert L.vkhrt_scene_create(C.byref(d), C.byref(h)) == -1
    d = api.SceneDesc(None, 3, None, 1, None, 0.02, 0, 0)                 # counts without arrays
    assert L.vkhrt_scene_create(C.byref(d), C.byref(h)) == -1
    assert L.vkhrt_scene_build(None) == -1 and L.vkhrt_render(None, None, None, None) == -1
    assert L.vkhrt_scene_primitive_count(None) == 0
    L.vkhrt_scene_destroy(None)                                            # no-op
    f = V.make_frame(np.eye(4), np.eye(4), 0, 0)
    assert V.frame_local_pixels(f) == 0                                    # bad frame -> 0
    f = V.make_frame(np.eye(4), np.eye(4), 100, 50, tile_size=12)
    assert V.frame_local_pixels(f) == 0                                    # tile size not a multiple of 8


def test_frame_local_pixels_and_tile_layout(V):
    from vkhrt_b200.multi import TileSharding
    for (W, H, T, world) in ((1920, 1080, 64, 1), (1920, 1080, 64, 2), (200, 120, 32, 3), (2720, 1530, 64, 8), (64, 64, 64, 4)):
        lay = TileSharding(W, H, world, T)
        for r in range(world):
            f = V.make_frame(np.eye(4), np.eye(4), W, H, **lay.frame_kwargs(r))
            assert V.frame_local_pixels(f) == lay.shard_pixels
        # every tile belongs to exactly one rank; gather_index is a bijection onto the valid pixels
        owners = [lay.rank_of_tile(t) for t in range(lay.n_tiles)]
        assert sorted(sum((lay.tiles_of_rank(r) for r in range(world)), [])) == list(range(lay.n_tiles))
        assert all(t in lay.tiles_of_rank(o) for t, o in enumerate(owners))
        if world > 1:
            gi = lay.gather_index()
            assert gi.shape[0] == W * H and np.unique(gi).shape[0] == W * H and gi.max() < world * lay.shard_pixels


def test_untile_host_equals_the_numpy_mirror(V):
    """vkhrt_untile_host (what vkhrt_render_multi assembles the frame with) against TileSharding.untile_host, on random shards"""
    from vkhrt_b200.multi import TileSharding
    rng = np.random.default_rng(3)
    for (W, H, T, world) in ((200, 120, 32, 3), (333, 77, 64, 2), (64, 64, 64, 4), (130, 70, 8, 5), (50, 40, 64, 1)):
        lay = TileSharding(W, H, world, T)
        f = V.make_frame(np.eye(4), np.eye(4), W, H, tile_size=T)
        n = world * lay.shard_pixels
        hits = np.frombuffer(rng.integers(0, 256, n * 32, dtype=np.uint8).tobytes(), V.HIT_DTYPE).copy()
        rgba = rng.integers(0, 256, (n, 4), dtype=np.uint8)
        assert V.untile_host(f, world, hits).tobytes() == lay.untile_host(hits).tobytes()
        assert np.array_equal(V.untile_host(f, world, rgba), lay.untile_host(rgba))
    L = V.lib()
    assert L.vkhrt_untile_host(None, 2, None, None, 4) == -1
    buf = np.zeros(64 * 64 * 2, np.uint8)
    f = V.make_frame(np.eye(4), np.eye(4), 64, 64)
    assert L.vkhrt_untile_host(C.byref(f), 2, buf.ctypes.data, buf.ctypes.data, 7) == -1           # element size
    assert L.vkhrt_untile_host(C.byref(f), 1, buf.ctypes.data, buf.ctypes.data, 7) == -1
    # render_multi validates before it touches a device
    assert L.vkhrt_render_multi(None, 2, C.byref(f), None, None) == -1
    arr = (C.c_void_p * 2)(None, None)
    assert L.vkhrt_render_multi(arr, 0, C.byref(f), None, None) == -1 and L.vkhrt_render_multi(arr, 2, C.byref(f), None, None) == -1


# ---------------------------------------------------------------- FlyCamera (source/fly_camera.cpp:25-35)
def _look_at(eye, centre, up):
    f = centre - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, f @ eye
    return m


def _perspective_rh_zo_flipped(fov_deg, aspect, n, f):
    th = np.tan(np.radians(fov_deg) / 2)
    m = np.zeros((4, 4))
    m[0, 0] = 1 / (aspect * th)
    m[1, 1] = -1 / th                       # [1][1] *= -1
    m[2, 2] = f / (n - f)
    m[3, 2] = -1
    m[2, 3] = -(f * n) / (f - n)
    return m


@pytest.mark.parametrize("pos,yaw,pitch,fov,aspect", [((0, 150, 20), -90, 0, 60, 16 / 9), ((3, 140, -7), 35, -20, 45, 1.0), ((0, 0, 0), 180, 60, 90, 2.0)])
def test_camera_matrices_against_glm_definitions(V, pos, yaw, pitch, fov, aspect):
    vi, pi = V.camera_matrices(position=pos, yaw=yaw, pitch=pitch, fov=fov, aspect=aspect, near=0.1, far=1000.0)
    vi, pi = vi.reshape(4, 4).T.astype(np.float64), pi.reshape(4, 4).T.astype(np.float64)   # column-major -> math
    front = np.array([np.cos(np.radians(yaw)) * np.cos(np.radians(pitch)), np.sin(np.radians(pitch)), np.sin(np.radians(yaw)) * np.cos(np.radians(pitch))])
    right = np.cross(front, [0, 1, 0])
    right /= np.linalg.norm(right)
    up = np.cross(right, front)
    eye = np.array(pos, np.float64)
    view = _look_at(eye, eye + front, up)
    proj = _perspective_rh_zo_flipped(fov, aspect, 0.1, 1000.0)
    assert np.allclose(vi, np.linalg.inv(view), atol=2e-4)
    assert np.allclose(pi, np.linalg.inv(proj), rtol=2e-5, atol=1e-5)
    assert np.allclose(vi[:3, 3], eye, atol=1e-4)                          # ray origin = viewInverse * (0,0,0,1)


def test_fly_camera_defaults(V):
    cam = V.FlyCamera()
    assert cam.position == (0.0, 150.0, 20.0) and cam.fov == 60.0 and cam.near == 0.1 and cam.far == 1000.0   # application.cpp:65-73
    vi, _ = cam.matrices()
    assert np.allclose(vi.reshape(4, 4).T[:3, 2], (0, 0, 1), atol=1e-6)   # looks down -z


# ---------------------------------------------------------------- synthetic groom
def test_groom_generator_contract(V):
    pos, idx = V.generate_groom(500, 16, V.GROOM_CURLY)
    pos2, idx2 = V.generate_groom(500, 16, V.GROOM_CURLY)
    assert pos.tobytes() == pos2.tobytes() and idx.tobytes() == idx2.tobytes()      # deterministic
    pos3, _ = V.generate_groom(500, 16, V.GROOM_CURLY, seed=1234)
    assert pos.tobytes() != pos3.tobytes()
    assert pos.shape == (500 * 17, 3) and idx.shape == (500 * 16, 2)
    # Assimp line-mesh shape: consecutive segments of one strand share a vertex, strands do not
    strand = idx.reshape(500, 16, 2)
    assert (strand[:, 1:, 0] == strand[:, :-1, 1]).all()
    assert (strand[1:, 0, 0] == strand[:-1, -1, 1] + 1).all()
    # roots on the head sphere (centre (0,150,0), radius 8), cap y >= -0.2; arc length 6
    st, _ = V.generate_groom(500, 16, V.GROOM_STRAIGHT)
    st = st.reshape(500, 17, 3).astype(np.float64)
    n = (st[:, 0] - [0, 150, 0]) / 8.0
    assert n[:, 1].min() >= -0.2 - 1e-5
    assert np.allclose(np.linalg.norm(st[:, 0] - [0, 150, 0], axis=1), 8.0, atol=1e-4)
    seg = np.linalg.norm(np.diff(st, axis=1), axis=2)
    assert np.allclose(seg, 6.0 / 16, atol=1e-4)
    d0 = st[:, 1] - st[:, 0]
    d1 = st[:, -1] - st[:, -2]
    assert np.allclose(np.cross(d0, d1), 0, atol=1e-5)                    # collinear
    # the straight and curly variants share their roots (same per-strand RNG stream)
    cur = pos.reshape(500, 17, 3).astype(np.float64)
    helix = cur[:, 0] - st[:, 0]
    assert np.allclose(np.linalg.norm(helix, axis=1), 0.25, atol=1e-4)    # curl amplitude at s = 0
    with pytest.raises(V.VkhrtError):
        V.generate_groom(10, 0, V.GROOM_CURLY)
