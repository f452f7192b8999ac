"""The C++ host layer (vkhrt_b200/host/): header-only mirror of ModelLoader / Model / FlyCamera / Renderer over the
C ABI, and the headless executable that stands in for the reference's `main`."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "vkhrt_b200", "_lib", "vkhrt_headless")

LOADER_PROBE = r"""
#include "vkhrt_host.hpp"
#include <cstdio>
using namespace vkhrt_host;
int main(int argc, char** argv) {
    ModelCreation m;
    if (!ModelLoader::LoadModel(argv[1], m)) { std::puts("FAIL"); return 1; }
    std::printf("%zu %zu\n", m.vertexBuffer.size(), m.indexBuffer.size() / 2);
    for (size_t i = 0; i < m.indexBuffer.size(); ++i) std::printf("%u ", m.indexBuffer[i]);
    std::printf("\n%g %g %g\n", m.vertexBuffer.back().x, m.vertexBuffer.back().y, m.vertexBuffer.back().z);
    ModelCreation c = ProcessHairCurves(m), d = ProcessHairDOTS(m), l = ProcessHairLSS(m);
    std::printf("%d %d %d\n", (int)c.technique, (int)d.technique, (int)l.technique);
    (void)argc; return 0;
}
"""


@pytest.fixture(scope="module")
def probe(tmp_path_factory, V):
    d = tmp_path_factory.mktemp("probe")
    src = d / "probe.cpp"
    src.write_text(LOADER_PROBE)
    exe = d / "probe"
    lib_dir = os.path.dirname(V.library_path())
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "vkhrt_b200", "host"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-lvkhrt_b200", f"-Wl,-rpath,{lib_dir}"])
    return str(exe)


def test_obj_polyline_loader_and_process_functions(probe, tmp_path):
    obj = tmp_path / "strands.obj"
    obj.write_text("# two strands\nv 0 0 0\nv 1 0 0\nv 2 1 0\nv 5 5 5\nv 6 5 5\nl 1 2 3\nl -2 -1\n")
    out = subprocess.run([probe, str(obj)], capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "5 3"
    assert out[1].split() == ["0", "1", "1", "2", "3", "4"]         # joints repeat the vertex index (GenerateCurves connectivity)
    assert out[2] == "6 5 5"
    assert out[3] == "0 2 1"                                        # ProcessHairCurves / DOTS / LSS select the technique
    assert subprocess.run([probe, str(tmp_path / "missing.obj")], capture_output=True, text=True).stdout.strip() == "FAIL"
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nl 1 7\n")
    assert subprocess.run([probe, str(bad)], capture_output=True, text=True).stdout.strip() == "FAIL"


def test_gltf_binary_through_the_cpp_loader(probe, V, tmp_path):
    """ModelLoader::LoadModel on a .glb (the reference loads glTF hair files, source/renderer.cpp:33-37): same arrays as the .obj path"""
    pos = np.float32([[0, 0, 0], [1, 0, 0], [2, 1, 0], [5, 5, 5], [6, 5, 5]])
    idx = np.uint32([[0, 1], [1, 2], [3, 4]])
    V.save_lines(str(tmp_path / "strands.glb"), pos, idx)
    out = subprocess.run([probe, str(tmp_path / "strands.glb")], capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "5 3" and out[1].split() == ["0", "1", "1", "2", "3", "4"] and out[2] == "6 5 5"
    (tmp_path / "cut.glb").write_bytes((tmp_path / "strands.glb").read_bytes()[:60])
    assert subprocess.run([probe, str(tmp_path / "cut.glb")], capture_output=True, text=True).stdout.strip() == "FAIL"


def test_synthetic_uri_matches_the_abi_generator(probe, V):
    out = subprocess.run([probe, "synthetic:curly:50:4"], capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "250 200"
    pos, idx = V.generate_groom(50, 4, V.GROOM_CURLY)
    assert [int(x) for x in out[1].split()] == idx.reshape(-1).tolist()
    assert np.allclose([float(x) for x in out[2].split()], pos[-1], rtol=1e-5)


def test_headless_cli_without_gpu(V):
    assert os.path.exists(EXE)
    r = subprocess.run([EXE, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "usage" in r.stdout
    if V.device_count() == 0:
        r = subprocess.run([EXE, "--model", "synthetic:curly:10:4"], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU path" in r.stderr       # loud failure, no fallback
    r = subprocess.run([EXE, "--model", "/nonexistent.obj"], capture_output=True, text=True)
    assert r.returncode == 1 and "[MODEL LOADING]" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("tech,name", [(0, "phantom"), (1, "lss"), (2, "dots")])
def test_headless_matches_python_binding_and_oracle(V, O, tmp_path, tech, name):
    W, H = 192, 108
    hits_path, ppm = tmp_path / "hits.bin", tmp_path / "out.ppm"
    r = subprocess.run([EXE, "--model", "synthetic:curly:3000:16", "--technique", name, "--size", f"{W}x{H}", "--frames", "2",
                        "--hits", str(hits_path), "--ppm", str(ppm)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hits = np.fromfile(hits_path, dtype=V.HIT_DTYPE)
    pos, idx = V.generate_groom(3000, 16, V.GROOM_CURLY)
    vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
    ho, io, _ = O.OracleScene(pos, idx, technique=tech).render(O.make_frame(vi, pi, W, H))
    assert hits.tobytes() == ho.tobytes()
    raw = open(ppm, "rb").read()
    header = f"P6\n{W} {H}\n255\n".encode()
    assert raw.startswith(header)
    assert np.array_equal(np.frombuffer(raw[len(header):], np.uint8).reshape(-1, 3), io[:, :3])


@pytest.mark.gpu
def test_headless_hair_asset_lod_environment_png(V, O, tmp_path):
    """the widened host path end to end: .hair asset -> LOD passes -> build -> environment miss shader -> PNG"""
    import struct
    import zlib
    W, H = 160, 96
    pos, idx = V.generate_groom(1500, 8, V.GROOM_CURLY)
    asset, png, hits_path = tmp_path / "groom.hair", tmp_path / "out.png", tmp_path / "hits.bin"
    V.save_lines(str(asset), pos, idx)
    r = subprocess.run([EXE, "--model", str(asset), "--technique", "phantom", "--size", f"{W}x{H}", "--lod", "1,0,1", "--env", "procedural",
                        "--png", str(png), "--hits", str(hits_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    orc = O.OracleScene(pos, idx, lod=(1, 0, 1))
    env = np.empty((512, 1024, 4), np.float32)
    V.lib().vkhrt_environment_generate(1024, 512, env.ctypes.data)
    orc.set_environment(env)
    vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
    ho, io, _ = orc.render(O.make_frame(vi, pi, W, H, miss_mode=1))
    assert np.fromfile(hits_path, dtype=V.HIT_DTYPE).tobytes() == ho.tobytes()
    data = png.read_bytes()
    n, = struct.unpack(">I", data[33:37])
    assert data[37:41] == b"IDAT"
    raw = np.frombuffer(zlib.decompress(data[41:41 + n]), np.uint8).reshape(H, 1 + 4 * W)[:, 1:].reshape(-1, 4)
    assert np.abs(raw.astype(np.int32) - io.astype(np.int32)).max() <= 1


MERGE_PROBE = r"""
#include "vkhrt_host.hpp"
#include <cstdio>
using namespace vkhrt_host;
int main(int argc, char** argv) {
    std::vector<ModelCreation> parts(argc - 1);
    for (int i = 1; i < argc; ++i) if (!ModelLoader::LoadModel(argv[i], parts[i - 1])) { std::puts("FAIL"); return 1; }
    parts[0].radiusBuffer.assign(parts[0].vertexBuffer.size(), 0.01f);        // one part with per-vertex radii: the others get their constant
    const ModelCreation m = MergeModels(parts);
    std::printf("%zu %zu %zu\n", m.vertexBuffer.size(), m.indexBuffer.size() / 2, m.radiusBuffer.size());
    for (uint32_t f : m.meshFirstSegment) std::printf("%u ", f);
    std::printf("\n");
    for (uint32_t i : m.indexBuffer) std::printf("%u ", i);
    std::printf("\n%g %g\n", m.radiusBuffer.front(), m.radiusBuffer.back());
    return 0;
}
"""


def test_merge_models_concatenates_like_generate_lines(V, tmp_path):
    """MergeModels: several models as one line list (firstVertex / firstIndex rebasing, geometry_processor.cpp:45-67), one mesh each"""
    (tmp_path / "a.obj").write_text("v 0 0 0\nv 1 0 0\nv 2 1 0\nl 1 2 3\n")
    (tmp_path / "b.obj").write_text("v 5 5 5\nv 6 5 5\nl 1 2\n")
    src, exe = tmp_path / "merge.cpp", tmp_path / "merge"
    src.write_text(MERGE_PROBE)
    lib_dir = os.path.dirname(V.library_path())
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "vkhrt_b200", "host"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-lvkhrt_b200", f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.run([str(exe), str(tmp_path / "a.obj"), str(tmp_path / "b.obj")], capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "5 3 5"
    assert out[1].split() == ["0", "2"]
    assert out[2].split() == ["0", "1", "1", "2", "3", "4"]
    assert out[3] == "0.01 0.02"


@pytest.mark.gpu
def test_headless_multi_model_scene(V, O, tmp_path):
    """`--model` twice = the reference's list of scene models (renderer.cpp:33-41) in ONE device scene: records equal the oracle's over the
    concatenated line list, and the per-mesh ray counts equal the mesh lookup of the Python binding."""
    W, H = 200, 120
    p1, i1 = V.generate_groom(2000, 12, V.GROOM_CURLY)
    p2, i2 = V.generate_groom(800, 6, V.GROOM_STRAIGHT, seed=3)
    p2 = p2 + np.float32([5.0, 0.0, 0.0])
    a, b, hits_path = tmp_path / "a.glb", tmp_path / "b.hair", tmp_path / "hits.bin"
    V.save_lines(str(a), p1, i1); V.save_lines(str(b), p2, i2)
    r = subprocess.run([EXE, "--model", str(a), "--model", str(b), "--technique", "phantom", "--size", f"{W}x{H}", "--hits", str(hits_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    l1 = V.load_lines(str(a)); l2 = V.load_lines(str(b))
    pos, idx, _, first = V.merge_meshes([(l1[0], l1[1]), (l2[0], l2[1])])
    vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
    ho, _, _ = O.OracleScene(pos, idx, technique=0).render(O.make_frame(vi, pi, W, H))
    hits = np.fromfile(hits_path, dtype=V.HIT_DTYPE)
    assert hits.tobytes() == ho.tobytes()
    hit = (hits["flags"] & 1) != 0
    mesh = np.searchsorted(first, hits["segment"][hit], side="right") - 1
    lines = r.stdout.split("\n")
    assert f"  mesh 0: {int((mesh == 0).sum())} rays" in lines and f"  mesh 1: {int((mesh == 1).sum())} rays" in lines
    assert (mesh == 1).sum() > 0


@pytest.mark.gpu
def test_headless_frames_in_flight(V, tmp_path):
    """`--in-flight`: the reference's frame loop (renderer.cpp:85-119: MAX_FRAMES_IN_FLIGHT frames submitted, the oldest waited on) through
    Renderer::Submit / Wait -> vkhrt_render_submit / _wait.  The last frame's records equal the blocking path's."""
    W, H = 320, 200
    a, b = tmp_path / "blocking.bin", tmp_path / "inflight.bin"
    common = ["--model", "synthetic:curly:3000:16", "--technique", "phantom", "--size", f"{W}x{H}", "--no-image"]
    r1 = subprocess.run([EXE, *common, "--frames", "2", "--hits", str(a)], capture_output=True, text=True)
    r2 = subprocess.run([EXE, *common, "--frames", "7", "--in-flight", "--hits", str(b)], capture_output=True, text=True)
    assert r1.returncode == 0 and r2.returncode == 0, r1.stderr + r2.stderr
    assert "7 frames, 2 in flight" in r2.stdout
    assert a.read_bytes() == b.read_bytes() and len(a.read_bytes()) == W * H * 32
