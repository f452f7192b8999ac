"""asset ingest (OBJ / HAIR line assets, Radiance .hdr), PNG output, strand LOD
(MergeLines / SplitLines / MergeCurvesFast) and the environment-map miss shader.

CPU part: host code of the product (no GPU needed) and the oracle restatements against hand-derived answers.
GPU part: the device passes against the oracle — bit-identical lines / curves / BVH / hit records; the environment
colour within 1 LSB of the 8-bit image (asin / atan / exp / pow are library functions on both sides)."""
import os
import struct
import zlib

import numpy as np
import pytest

from conftest import default_camera, has_gpu

gpu = pytest.mark.gpu


# ---------------------------------------------------------------------------------- asset ingest (host, CPU)
def test_obj_polylines_negative_indices_comments(V, tmp_path):
    p = tmp_path / "a.obj"
    p.write_text("# two strands\nv 0 0 0\nv 1 0 0\nv 2 1 0 # inline comment\n l 1 2 3\nv 5 5 5\nv 6 5 5\nl -2 -1\nvn 0 1 0\nf 1 2 3\nl 4/1 5/2\n")
    pos, idx, rpv, strands = V.load_lines(str(p))
    assert rpv is None and strands == 3
    assert pos.tolist() == [[0, 0, 0], [1, 0, 0], [2, 1, 0], [5, 5, 5], [6, 5, 5]]
    assert idx.tolist() == [[0, 1], [1, 2], [3, 4], [3, 4]]       # `l 1 2 3` -> two 2-index faces, like Assimp


def test_obj_errors(V, tmp_path):
    p = tmp_path / "bad.obj"
    p.write_text("v 0 0 0\nl 1 2\n")
    with pytest.raises(V.VkhrtError) as e:
        V.load_lines(str(p))
    assert e.value.status == -6                                    # BAD_TOPOLOGY
    with pytest.raises(V.VkhrtError) as e:
        V.load_lines(str(tmp_path / "missing.obj"))
    assert e.value.status == -8                                    # IO
    q = tmp_path / "x.abc"
    q.write_text("")
    with pytest.raises(V.VkhrtError) as e:
        V.load_lines(str(q))
    assert e.value.status == -7                                    # UNSUPPORTED


@pytest.mark.parametrize("ext", ["obj", "hair", "glb"])
def test_line_asset_round_trip_is_exact(V, tmp_path, ext):
    pos, idx = V.generate_groom(37, 5, V.GROOM_CURLY)
    path = str(tmp_path / f"g.{ext}")
    V.save_lines(path, pos, idx)
    p2, i2, rpv, strands = V.load_lines(path)
    assert strands == 37 and rpv is None
    assert p2.tobytes() == pos.tobytes() and np.array_equal(i2, idx)     # %.9g text round-trips fp32 exactly
    if ext == "glb":
        # the container a glTF importer (Assimp in the reference) expects: header, 4-byte aligned JSON + BIN chunks, POSITION min / max
        import json
        raw = open(path, "rb").read()
        magic, version, total = struct.unpack("<4sII", raw[:12])
        jlen, jtype = struct.unpack("<II", raw[12:20])
        assert magic == b"glTF" and version == 2 and total == len(raw) and jtype == 0x4E4F534A and jlen % 4 == 0
        doc = json.loads(raw[20:20 + jlen])
        blen, btype = struct.unpack("<II", raw[20 + jlen:28 + jlen])
        assert btype == 0x004E4942 and blen == pos.nbytes + idx.nbytes == doc["buffers"][0]["byteLength"] and 28 + jlen + blen == total
        acc = doc["accessors"][0]
        assert doc["meshes"][0]["primitives"][0]["mode"] == 1 and acc["count"] == pos.shape[0]
        assert np.array_equal(np.float32(acc["min"]), pos.min(axis=0)) and np.array_equal(np.float32(acc["max"]), pos.max(axis=0))
        V.save_lines(path, np.zeros((0, 3), np.float32), np.zeros((0, 2), np.uint32))       # an empty asset is still a valid file
        p3, i3, _, _ = V.load_lines(path)
        assert p3.shape == (0, 3) and i3.shape == (0, 2)


def test_hair_file_layout_and_thickness(V, tmp_path):
    # Cem Yuksel HAIR: 128-byte header, u16 segments per strand, points, thickness (diameter -> radius/2)
    pts = np.arange(21, dtype=np.float32).reshape(7, 3)
    segs = np.array([2, 3], np.uint16)                             # 3 + 4 points
    thick = np.linspace(0.04, 0.01, 7).astype(np.float32)
    hdr = b"HAIR" + struct.pack("<IIII", 2, 7, 1 | 2 | 4, 0) + struct.pack("<ff3f", 0.04, 0.0, 1, 1, 1) + b"\0" * 88
    assert len(hdr) == 128
    path = tmp_path / "h.hair"
    path.write_bytes(hdr + segs.tobytes() + pts.tobytes() + thick.tobytes())
    pos, idx, rpv, strands = V.load_lines(str(path))
    assert strands == 2 and np.array_equal(pos, pts)
    assert idx.tolist() == [[0, 1], [1, 2], [3, 4], [4, 5], [5, 6]]
    assert np.array_equal(rpv, thick * np.float32(0.5))
    # default segment count, no segments array
    hdr2 = b"HAIR" + struct.pack("<IIII", 2, 6, 2, 2) + struct.pack("<ff3f", 0.04, 0.0, 1, 1, 1) + b"\0" * 88
    path.write_bytes(hdr2 + pts[:6].tobytes())
    pos, idx, rpv, strands = V.load_lines(str(path))
    assert rpv is None and idx.tolist() == [[0, 1], [1, 2], [3, 4], [4, 5]]
    # writer: thickness round trip
    V.save_lines(str(path), pts, [[0, 1], [1, 2], [3, 4], [4, 5], [5, 6]], radius_per_vertex=thick)
    _, i3, r3, s3 = V.load_lines(str(path))
    assert s3 == 2 and len(i3) == 5 and np.array_equal(r3, thick)
    path.write_bytes(hdr[:100])
    with pytest.raises(V.VkhrtError):
        V.load_lines(str(path))


def _rgbe_decode(b):
    """numpy restatement of stb_image's stbi__hdr_convert (req_comp 4)"""
    b = b.astype(np.int32)
    f = np.ldexp(np.float32(1.0), b[..., 3] - 136).astype(np.float32)
    out = np.zeros(b.shape[:-1] + (4,), np.float32)
    out[..., :3] = b[..., :3].astype(np.float32) * f[..., None]
    out[b[..., 3] == 0, :3] = 0
    out[..., 3] = 1
    return out


def test_hdr_flat_and_rle_scanlines(V, tmp_path):
    rng = np.random.default_rng(3)
    W, H = 40, 6
    rgbe = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    rgbe[0, 0] = (5, 6, 7, 0)                                      # zero exponent -> black
    rgbe[:, 0, 0] = 9                                              # keep flat rows from looking like the RLE marker
    want = _rgbe_decode(rgbe)
    head = b"#?RADIANCE\n# comment\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n-Y %d +X %d\n" % (H, W)
    flat = tmp_path / "flat.hdr"
    flat.write_bytes(head + rgbe.tobytes())
    assert np.array_equal(V.load_hdr(str(flat)), want)
    # new-style RLE: per scanline 2 2 hi lo, then each component as runs / literals
    body = b""
    for y in range(H):
        body += bytes([2, 2, W >> 8, W & 255])
        for c in range(4):
            row = rgbe[y, :, c]
            if c == 1:
                row = np.full(W, row[0], np.uint8); rgbe[y, :, 1] = row          # one long run
                body += bytes([128 + W, int(row[0])]) if W <= 127 else b""
            else:
                for x in range(0, W, 16):
                    chunk = row[x:x + 16]
                    body += bytes([len(chunk)]) + chunk.tobytes()
    rle = tmp_path / "rle.hdr"
    rle.write_bytes(head + body)
    assert np.array_equal(V.load_hdr(str(rle)), _rgbe_decode(rgbe))
    # writer -> reader: RGBE quantisation only (<= 1/128 relative on the largest channel)
    env = V.generate_environment(64, 32)
    out = tmp_path / "env.hdr"
    V.save_hdr(str(out), env)
    back = V.load_hdr(str(out))
    assert back.shape == env.shape and (back[..., 3] == 1).all()
    mx = env[..., :3].max(axis=-1, keepdims=True)
    assert (np.abs(back[..., :3] - env[..., :3]) <= mx / 100.0 + 1e-6).all()
    (tmp_path / "bad.hdr").write_bytes(b"P6\n")
    with pytest.raises(V.VkhrtError):
        V.load_hdr(str(tmp_path / "bad.hdr"))


def test_png_writer_decodes_with_zlib(V, tmp_path):
    rng = np.random.default_rng(4)
    for (W, H) in ((7, 5), (300, 90)):                             # second one spans several 64 KiB stored blocks
        img = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        path = tmp_path / "o.png"
        V.save_png(str(path), img, W, H)
        data = path.read_bytes()
        assert data[:8] == b"\x89PNG\r\n\x1a\n"
        pos, chunks = 8, []
        while pos < len(data):
            n, = struct.unpack(">I", data[pos:pos + 4]); typ = data[pos + 4:pos + 8]; body = data[pos + 8:pos + 8 + n]
            crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
            assert zlib.crc32(typ + body) == crc
            chunks.append((typ, body)); pos += 12 + n
        assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
        assert struct.unpack(">IIBBBBB", chunks[0][1]) == (W, H, 8, 6, 0, 0, 0)
        raw = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(H, 1 + 4 * W)     # checks the adler32 too
        assert (raw[:, 0] == 0).all() and np.array_equal(raw[:, 1:].reshape(H, W, 4), img)


# ---------------------------------------------------------------------------------- LOD restatements (oracle, CPU)
def _two_strands():
    # strand A: 5 segments (6 vertices), strand B: 2 segments; B is not connected to A
    a = np.array([[0, 0, 0], [1, 0, 0], [2, 1, 0], [3, 1, 1], [4, 2, 1], [5, 2, 2]], np.float32)
    b = np.array([[10, 0, 0], [10, 1, 0], [10, 2, 1]], np.float32)
    pos = np.concatenate([a, b])
    idx = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [6, 7], [7, 8]], np.uint32)
    return pos, idx


def test_oracle_merge_and_split_lines_by_hand(O):
    pos, idx = _two_strands()
    base = O.OracleScene(pos, idx).lines()
    assert base.shape == (7, 6)
    # MergeLines (geometry_processor.cpp:69-104): pairs (0,1) (2,3) merge, pair (4,5) is not connected -> kept, line 6 dropped (odd)
    m = O.OracleScene(pos, idx, lod=(0, 1, 0)).lines()
    want = np.array([[0, 0, 0, 2, 1, 0], [2, 1, 0, 4, 2, 1], [4, 2, 1, 5, 2, 2], [10, 0, 0, 10, 1, 0]], np.float32)
    assert np.array_equal(m, want)
    # SplitLines (:106-121): midpoint (s + e) * 0.5
    s = O.OracleScene(pos, idx, lod=(1, 0, 0)).lines()
    assert s.shape == (14, 6)
    assert np.array_equal(s[0], [0, 0, 0, 0.5, 0, 0]) and np.array_equal(s[1], [0.5, 0, 0, 1, 0, 0])
    assert np.array_equal(s[0::2, :3], base[:, :3]) and np.array_equal(s[1::2, 3:], base[:, 3:]) and np.array_equal(s[0::2, 3:], s[1::2, :3])
    # split then merge gives the original lines back (exactly: the midpoints are dropped again)
    sm = O.OracleScene(pos, idx, lod=(1, 1, 0)).lines()
    assert np.array_equal(sm, base)


def test_oracle_merge_curves_fast_by_hand(O):
    pos, idx = _two_strands()
    c = O.OracleScene(pos, idx).primitives().reshape(-1, 4, 3)
    m = O.OracleScene(pos, idx, lod=(0, 0, 1)).primitives().reshape(-1, 4, 3)
    assert m.shape[0] == 4                                          # (0,1) (2,3) merged, (4,5) kept as two, curve 6 dropped
    mid = (c[0, 2] + c[1, 1]) * np.float32(0.5)                      # geometry_processor.cpp:190-192
    assert np.array_equal(m[0, 0], c[0, 0]) and np.array_equal(m[0, 3], c[1, 3])
    assert np.array_equal(m[0, 1], (c[0, 1] + mid) * np.float32(0.5)) and np.array_equal(m[0, 2], (mid + c[1, 2]) * np.float32(0.5))
    assert np.array_equal(m[2], c[4]) and np.array_equal(m[3], c[5])
    with pytest.raises(ValueError):
        O.OracleScene(pos, idx, technique=1, lod=(0, 0, 1))          # curve merge needs curves


def test_oracle_environment_lookup_by_hand(O, V):
    pos, idx = _two_strands()
    sc = O.OracleScene(pos, idx)
    env = np.zeros((4, 8, 4), np.float32)
    env[..., 3] = 1
    env[0, :, 0] = 2.0                                              # top row red = what a ray pointing up sees (dir = -d looks down: v -> 0)
    env[3, :, 2] = 1.0                                              # bottom row blue
    sc.set_environment(env)
    tone = lambda x: (1.0 - np.exp(-x)) ** (1.0 / 2.2)              # miss.rmiss:34-35
    # straight up: dir.y = -1 -> v = 0 exactly -> texel row -0.5: repeat addressing blends row 0 with row 3 half and half
    up = sc.environment_miss([0, 1, 0])
    assert abs(up[0] - tone(1.0)) < 1e-6 and up[1] == 0 and abs(up[2] - tone(0.5)) < 1e-6
    # centre of row 0: v = 0.125 -> elevation of dir = -0.375 pi
    c, s_ = np.cos(0.375 * np.pi), np.sin(0.375 * np.pi)
    r0 = sc.environment_miss([0, s_, -c])
    assert abs(r0[0] - tone(2.0)) < 1e-4 and r0[1] == 0 and r0[2] < 0.01
    r3 = sc.environment_miss([0, -s_, -c])
    assert abs(r3[2] - tone(1.0)) < 1e-4 and r3[0] < 0.01
    # horizontal ray: v = 0.5 -> halfway between rows 1 and 2 (both black)
    assert np.array_equal(sc.environment_miss([0, 0, -1]), [0, 0, 0])
    # u: dir = -d; theta = atan(dir.x, -dir.z): d = (0,0,-1) -> dir = (0,0,1) -> theta = atan2(0,-1) = pi -> u = 1 (wraps to texel column 0/7 seam)
    env2 = np.zeros((2, 8, 4), np.float32); env2[:, 0, 1] = 1.0; env2[:, 7, 1] = 1.0
    sc.set_environment(env2)
    assert abs(sc.environment_miss([0, 0, -1])[1] - tone(1.0)) < 1e-6    # repeat addressing across the seam
    assert sc.environment_miss([0, 0, 1])[1] == 0                      # u = 0.5 -> columns 3/4


# ---------------------------------------------------------------------------------- device passes vs oracle (GPU)
@gpu
@pytest.mark.parametrize("tech", [0, 1, 2])
@pytest.mark.parametrize("lod", [(1, 0, 0), (0, 1, 0), (2, 1, 0), (0, 3, 0), (1, 2, 0)])
def test_gpu_line_lod_matches_oracle(V, O, tech, lod):
    pos, idx = V.generate_groom(301, 7, V.GROOM_CURLY)               # 7 segments: odd strands exercise the unconnected-pair branch
    rpv = None
    if tech == 1:
        rpv = np.linspace(0.03, 0.005, pos.shape[0]).astype(np.float32)
    with V.Scene(pos, idx, technique=tech, radius_per_vertex=rpv) as sc:
        sc.apply_lod(*lod)
        orc = O.OracleScene(pos, idx, technique=tech, radius_per_vertex=rpv, lod=lod)
        assert sc.n_segments == orc.lines().shape[0]
        assert sc.lines().tobytes() == orc.lines().tobytes()
        sc.build()
        assert sc.primitives().tobytes() == orc.primitives().tobytes()
        gn, gi, gm, _ = sc.bvh(); on, oi, om, _ = orc.bvh()
        assert gn.tobytes() == on.tobytes() and np.array_equal(gi, oi) and np.array_equal(gm, om)
        W, H = 160, 96
        vi, pi = default_camera(V, W, H)
        hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H))
        ho, io, _ = orc.render(O.make_frame(vi, pi, W, H))
        assert hg.tobytes() == ho.tobytes() and np.array_equal(ig, io)
        assert (hg["flags"] & 1).sum() > 50
        with pytest.raises(V.VkhrtError) as e:
            sc.refit(pos)
        assert e.value.status == -7


@gpu
@pytest.mark.parametrize("lod", [(0, 0, 1), (0, 0, 2), (1, 0, 1), (0, 1, 2)])
def test_gpu_curve_merge_matches_oracle(V, O, lod):
    pos, idx = V.generate_groom(257, 9, V.GROOM_CURLY)
    with V.Scene(pos, idx) as sc:
        sc.apply_lod(*lod)
        orc = O.OracleScene(pos, idx, lod=lod)
        assert sc.n_primitives == orc.n_primitives
        assert sc.lines().tobytes() == orc.lines().tobytes()
        sc.build()
        assert sc.primitives().tobytes() == orc.primitives().tobytes()
        gn, gi, _, _ = sc.bvh(); on, oi, _, _ = orc.bvh()
        assert gn.tobytes() == on.tobytes() and np.array_equal(gi, oi)
        W, H = 160, 96
        vi, pi = default_camera(V, W, H)
        hg, _, _ = sc.render(V.make_frame(vi, pi, W, H), rgba=False)
        ho, _, _ = orc.render(O.make_frame(vi, pi, W, H), rgba=False)
        assert hg.tobytes() == ho.tobytes()


@gpu
def test_gpu_lod_argument_errors(V):
    pos, idx = V.generate_groom(8, 4, V.GROOM_STRAIGHT)
    with V.Scene(pos, idx, technique=V.LSS) as sc:
        with pytest.raises(V.VkhrtError) as e:
            sc.apply_lod(0, 0, 1)
        assert e.value.status == -7
        sc.apply_lod(0, 0, 0)                                        # no-op
        sc.build()
        sc.refit(pos)                                                # still allowed
        with pytest.raises(V.VkhrtError) as e:
            sc.apply_lod(1, 0, 0)
        assert e.value.status == -1                                  # after build
    with V.Scene(pos[:0], idx[:0]) as sc:                           # empty scene through the passes
        sc.apply_lod(1, 1, 1)
        assert sc.n_segments == 0
        sc.build()


@gpu
@pytest.mark.parametrize("spp", [1, 3])
def test_gpu_environment_miss_within_one_lsb(V, O, spp):
    pos, idx = V.generate_groom(2000, 12, V.GROOM_CURLY)
    env = V.generate_environment(256, 128)
    W, H = 320, 200
    vi, pi = V.camera_matrices(position=(0.0, 152.0, 26.0), pitch=8.0, aspect=float(np.float32(W) / np.float32(H)))
    with V.Scene(pos, idx) as sc:
        sc.build()
        with pytest.raises(V.VkhrtError) as e:
            sc.render(V.make_frame(vi, pi, W, H, miss_mode=V.MISS_ENVIRONMENT))
        assert e.value.status == -1                                  # no environment set
        sc.set_environment(env)
        hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H, spp=spp, miss_mode=V.MISS_ENVIRONMENT))
        orc = O.OracleScene(pos, idx)
        orc.set_environment(env)
        ho, io, _ = orc.render(O.make_frame(vi, pi, W, H, spp=spp, miss_mode=1))
        assert hg.tobytes() == ho.tobytes()                          # hit records are untouched by the miss shader
        miss = (hg["flags"] & 1) == 0
        assert 0.2 < miss.mean() < 0.95
        d = np.abs(ig.astype(np.int32) - io.astype(np.int32))
        assert d.max() <= 1 and (d[miss] != 0).mean() < 0.01         # library transcendentals: 1 LSB, rarely
        if spp == 1:
            assert np.array_equal(ig[~miss], io[~miss])
        assert len(np.unique(ig[miss][:, :3], axis=0)) > 50          # a real gradient, not a constant
        # constant mode still works with a map installed
        _, ic, _ = sc.render(V.make_frame(vi, pi, W, H, miss_rgb=(0.25, 0.5, 0.75)))
        assert (ic[miss][:, :3] == np.array([64, 128, 191], np.uint8)).all()
        sc.set_environment(None)
        with pytest.raises(V.VkhrtError):
            sc.render(V.make_frame(vi, pi, W, H, miss_mode=V.MISS_ENVIRONMENT))


# ---------------------------------------------------------------- glTF 2.0 line primitives (the format of the reference's own scene)
def _gltf_doc(bin_len, uri=None, extra=None):
    """two meshes: mesh 0 = LINES with u16 indices over 4 interleaved (strided) vertices, mesh 1 = LINE_STRIP without indices over 3
    vertices, a triangle primitive that must be ignored, and a node hierarchy with a matrix, a TRS child and an untransformed root"""
    doc = {
        "asset": {"version": "2.0"},
        "scene": 0,
        "scenes": [{"nodes": [0, 2]}],
        "nodes": [
            {"name": "root", "matrix": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 10, 20, 30, 1], "children": [1]},
            {"name": "child", "mesh": 1, "translation": [1, 0, 0], "rotation": [0, 0, 0.7071067811865476, 0.7071067811865476], "scale": [2, 2, 2]},
            {"name": "plain", "mesh": 0},
        ],
        "meshes": [
            {"primitives": [{"mode": 1, "attributes": {"POSITION": 0}, "indices": 1},
                            {"mode": 4, "attributes": {"POSITION": 0}, "indices": 1}]},
            {"primitives": [{"mode": 3, "attributes": {"POSITION": 2}}]},
        ],
        "accessors": [
            {"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3"},
            {"bufferView": 1, "componentType": 5123, "count": 6, "type": "SCALAR"},
            {"bufferView": 2, "byteOffset": 4, "componentType": 5126, "count": 3, "type": "VEC3"},
        ],
        "bufferViews": [
            {"buffer": 0, "byteOffset": 0, "byteLength": 80, "byteStride": 20},
            {"buffer": 0, "byteOffset": 80, "byteLength": 12},
            {"buffer": 0, "byteOffset": 92, "byteLength": 40},
        ],
        "buffers": [{"byteLength": bin_len}],
    }
    if uri is not None:
        doc["buffers"][0]["uri"] = uri
    if extra:
        doc.update(extra)
    return doc


def _gltf_bin():
    v0 = np.array([[0, 0, 0], [1, 0, 0], [2, 1, 0], [3, 1, 1]], np.float32)
    inter = b"".join(v0[i].tobytes() + struct.pack("<ff", 9.0, 9.0) for i in range(4))      # 12 bytes position + 8 bytes of other attributes
    ind = np.array([0, 1, 1, 2, 2, 3], np.uint16).tobytes()
    v1 = np.array([[0, 0, 0], [0, 1, 0], [0, 2, 0.5]], np.float32)
    blob = inter + ind + struct.pack("<f", 123.0) + v1.tobytes()
    assert len(blob) == 80 + 12 + 40
    return blob, v0, v1


def _gltf_expected(v0, v1):
    # node "child": world = root(translate 10,20,30) * T(1,0,0) * Rz(90 deg) * S(2): p -> (10 + 1 - 2 y, 20 + 2 x, 30 + 2 z)
    w1 = np.stack([11.0 - 2.0 * v1[:, 1].astype(np.float64), 20.0 + 2.0 * v1[:, 0].astype(np.float64), 30.0 + 2.0 * v1[:, 2].astype(np.float64)], axis=1)
    pos = np.concatenate([w1.astype(np.float32), v0])
    idx = np.array([[0, 1], [1, 2], [3, 4], [4, 5], [5, 6]], np.uint32)
    return pos, idx


@pytest.mark.parametrize("container", ["external", "base64", "glb", "glb-padded"])
def test_gltf_line_primitives(V, tmp_path, container):
    import base64
    import json
    blob, v0, v1 = _gltf_bin()
    if container == "external":
        (tmp_path / "hair data.bin").write_bytes(blob)
        path = tmp_path / "a.gltf"
        path.write_text(json.dumps(_gltf_doc(len(blob), "hair%20data.bin")))
    elif container == "base64":
        path = tmp_path / "a.gltf"
        path.write_text(json.dumps(_gltf_doc(len(blob), "data:application/octet-stream;base64," + base64.b64encode(blob).decode()), indent=2))
    else:
        js = json.dumps(_gltf_doc(len(blob))).encode()
        if container == "glb-padded":
            js += b" "
        js += b" " * (-len(js) % 4)
        body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
        path = tmp_path / "a.glb"
        path.write_bytes(b"glTF" + struct.pack("<II", 2, 12 + len(body)) + body)
    pos, idx, rpv, strands = V.load_lines(str(path))
    epos, eidx = _gltf_expected(v0, v1)
    assert rpv is None and strands == 2
    assert np.array_equal(idx, eidx)
    assert np.array_equal(pos[3:], epos[3:])                       # untransformed node: the file's floats, bit for bit
    assert np.allclose(pos[:3], epos[:3], rtol=0, atol=2e-6)       # transformed node: double arithmetic rounded once
    # consecutive segments of a strand share the vertex, which is what GenerateCurves keys on
    assert idx[0, 1] == idx[1, 0] and idx[2, 1] == idx[3, 0]


def test_gltf_loop_no_scene_and_errors(V, tmp_path):
    import base64
    import json
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    uri = "data:application/octet-stream;base64," + base64.b64encode(tri.tobytes()).decode()
    base = {"asset": {"version": "2.0"}, "meshes": [{"primitives": [{"mode": 2, "attributes": {"POSITION": 0}}]}],
            "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}],
            "bufferViews": [{"buffer": 0, "byteLength": 36}], "buffers": [{"byteLength": 36, "uri": uri}]}
    p = tmp_path / "loop.gltf"
    p.write_text(json.dumps(base))                                  # no nodes at all: meshes as they are
    pos, idx, _, strands = V.load_lines(str(p))
    assert np.array_equal(pos, tri) and np.array_equal(idx, [[0, 1], [1, 2], [2, 0]]) and strands == 1
    # a mesh instanced by two nodes is emitted twice
    two = dict(base, nodes=[{"mesh": 0}, {"mesh": 0, "translation": [0, 0, 5]}])
    p.write_text(json.dumps(two))
    pos, idx, _, _ = V.load_lines(str(p))
    assert pos.shape == (6, 3) and idx.shape == (6, 2) and np.array_equal(pos[3:], tri + np.float32([0, 0, 5]))

    def expect(doc, status):
        p.write_text(json.dumps(doc))
        with pytest.raises(V.VkhrtError) as e:
            V.load_lines(str(p))
        assert e.value.status == status, e.value
    expect(dict(base, extensionsRequired=["KHR_draco_mesh_compression"]), -7)                                          # unsupported
    expect(dict(base, accessors=[{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3"}]), -8)          # runs past its view
    expect(dict(base, accessors=[{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3", "sparse": {"count": 1}}]), -7)
    expect(dict(base, accessors=[{"bufferView": 0, "componentType": 5123, "count": 3, "type": "VEC3"}]), -7)          # positions must be float
    expect(dict(base, bufferViews=[{"buffer": 0, "byteLength": 360}]), -8)                                            # view past the buffer
    expect(dict(base, buffers=[{"byteLength": 36, "uri": "missing.bin"}]), -8)
    expect(dict(base, scenes=[{"nodes": [0]}], nodes=[{"mesh": 0, "children": [0]}]), -8)                                                       # a cycle
    bad_idx = dict(base, meshes=[{"primitives": [{"mode": 1, "attributes": {"POSITION": 0}, "indices": 1}]}])
    bad_idx["accessors"] = base["accessors"] + [{"bufferView": 1, "componentType": 5125, "count": 2, "type": "SCALAR"}]
    bad_idx["bufferViews"] = base["bufferViews"] + [{"buffer": 1, "byteLength": 8}]
    bad_idx["buffers"] = base["buffers"] + [{"byteLength": 8, "uri": "data:application/octet-stream;base64," + base64.b64encode(struct.pack("<II", 0, 7)).decode()}]
    expect(bad_idx, -6)                                                                                                # index out of range
    # a hierarchy that is a DAG (every node lists the next one twice) would expand to 2^24 mesh instances: cut off, not expanded
    bomb = dict(base, scenes=[{"nodes": [0]}], nodes=[{"children": [i + 1, i + 1]} for i in range(24)] + [{"mesh": 0}])
    expect(bomb, -8)
    p.write_text("{ not json")
    with pytest.raises(V.VkhrtError):
        V.load_lines(str(p))


def test_gltf_groom_renders_like_the_same_lines(V, O, tmp_path):
    """a groom written as glTF LINES by hand and read back feeds the oracle exactly like the original arrays (CPU only)"""
    import json
    pos, idx = V.generate_groom(60, 6, V.GROOM_CURLY)
    blob = pos.tobytes() + idx.astype(np.uint32).tobytes()
    (tmp_path / "g.bin").write_bytes(blob)
    doc = {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
           "meshes": [{"primitives": [{"mode": 1, "attributes": {"POSITION": 0}, "indices": 1}]}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": int(pos.shape[0]), "type": "VEC3"},
                         {"bufferView": 1, "componentType": 5125, "count": int(idx.size), "type": "SCALAR"}],
           "bufferViews": [{"buffer": 0, "byteLength": pos.nbytes}, {"buffer": 0, "byteOffset": pos.nbytes, "byteLength": idx.size * 4}],
           "buffers": [{"byteLength": len(blob), "uri": "g.bin"}]}
    (tmp_path / "g.gltf").write_text(json.dumps(doc))
    p2, i2, _, strands = V.load_lines(str(tmp_path / "g.gltf"))
    assert np.array_equal(p2, pos) and np.array_equal(i2, idx) and strands == 60
    W, H = 64, 48
    vi, pi = default_camera(V, W, H)
    f = O.make_frame(vi, pi, W, H)
    h1, _, _ = O.OracleScene(pos, idx).render(f)
    h2, _, _ = O.OracleScene(p2, i2).render(f)
    assert h1.tobytes() == h2.tobytes() and (h1["flags"] & 1).sum() > 10


def test_loaders_survive_truncated_and_corrupt_files(V, tmp_path):
    """every prefix of a valid file and a few hundred random mutations: the readers must return an error or an asset, never crash
    (the reference logs and returns nullptr on a failed load, model_loader.cpp:280-284)"""
    rng = np.random.default_rng(11)
    pos, idx = V.generate_groom(5, 3, V.GROOM_CURLY)
    good = {}
    for ext in ("obj", "hair"):
        p = tmp_path / f"g.{ext}"
        V.save_lines(str(p), pos, idx, radius_per_vertex=np.full(pos.shape[0], 0.02, np.float32) if ext == "hair" else None)
        good[ext] = p.read_bytes()
    hp = tmp_path / "e.hdr"
    V.save_hdr(str(hp), V.generate_environment(16, 8))
    good["hdr"] = hp.read_bytes()
    import base64
    import json
    blob, _, _ = _gltf_bin()
    good["gltf"] = json.dumps(_gltf_doc(len(blob), "data:application/octet-stream;base64," + base64.b64encode(blob).decode())).encode()
    js = json.dumps(_gltf_doc(len(blob))).encode()
    js += b" " * (-len(js) % 4)
    body = struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob
    good["glb"] = b"glTF" + struct.pack("<II", 2, 12 + len(body)) + body
    n_ok = n_err = 0
    for ext, data in good.items():
        cases = [data[:k] for k in range(0, len(data), max(1, len(data) // 60))]
        for _ in range(150):
            b = bytearray(data)
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            cases.append(bytes(b))
        for k, c in enumerate(cases):
            f = tmp_path / f"fuzz{k}.{ext}"
            f.write_bytes(c)
            try:
                if ext == "hdr":
                    img = V.load_hdr(str(f))
                    assert img.ndim == 3 and img.shape[2] == 4
                else:
                    p2, i2, r2, _ = V.load_lines(str(f))
                    assert i2.size == 0 or int(i2.max()) < p2.shape[0]
                n_ok += 1
            except V.VkhrtError as e:
                assert e.status in (-6, -7, -8, -4)
                n_err += 1
            f.unlink()
    assert n_ok > 50 and n_err > 50


def test_oracle_lod_and_miss_shader_against_the_second_transcription(O, V):
    """tests/glsl_transcription.py restates MergeLines / SplitLines / MergeCurvesFast over Python lists and miss.rmiss in float64,
    independently of oracle/vkhrt_oracle.cpp: random ragged polylines (unconnected pairs, odd counts) and random directions"""
    import glsl_transcription as G
    rng = np.random.default_rng(21)
    for trial in range(6):
        # ragged strands of 1..6 segments, some strands start where the previous one ended (connected across strands on purpose)
        pos, idx, base = [], [], 0
        for s in range(int(rng.integers(3, 9))):
            k = int(rng.integers(1, 7))
            p = rng.normal(size=(k + 1, 3)).astype(np.float32)
            if pos and rng.random() < 0.3:
                p[0] = pos[-1]
            pos.extend(p); idx.extend([[base + j, base + j + 1] for j in range(k)]); base += k + 1
        pos = np.array(pos, np.float32); idx = np.array(idx, np.uint32)
        lines = [(pos[a], pos[b]) for a, b in idx]
        flat = lambda ls: np.array([np.concatenate(l) for l in ls], np.float32).reshape(-1, 6)
        for lod, fn in (((0, 1, 0), lambda l: G.merge_lines(l)), ((1, 0, 0), lambda l: G.split_lines(l)),
                        ((1, 2, 0), lambda l: G.merge_lines(G.merge_lines(G.split_lines(l)))), ((0, 2, 0), lambda l: G.merge_lines(G.merge_lines(l)))):
            assert np.array_equal(O.OracleScene(pos, idx, lod=lod).lines(), flat(fn(lines))), (trial, lod)
        curves = [list(c) for c in O.OracleScene(pos, idx).primitives().reshape(-1, 4, 3)]
        for passes in (1, 2):
            want = curves
            for _ in range(passes):
                want = G.merge_curves_fast(want)
            got = O.OracleScene(pos, idx, lod=(0, 0, passes)).primitives().reshape(-1, 4, 3)
            assert np.array_equal(got, np.array(want, np.float32).reshape(-1, 4, 3)), (trial, passes)
    env = V.generate_environment(64, 32)
    sc = O.OracleScene(pos, idx)
    sc.set_environment(env)
    for _ in range(300):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        assert np.abs(sc.environment_miss(d.astype(np.float32)) - G.miss_shader(env, d)).max() < 2e-4


def _read_exr_uncompressed(path):
    """minimal reader for what vkhrt_image_save_exr writes: scanline, NO_COMPRESSION, FLOAT channels"""
    import struct
    b = open(path, "rb").read()
    assert b[:4] == bytes([0x76, 0x2F, 0x31, 0x01]) and struct.unpack_from("<I", b, 4)[0] == 2
    p = 8
    attrs = {}
    while b[p] != 0:
        e = b.index(0, p); name = b[p:e].decode(); p = e + 1
        e = b.index(0, p); typ = b[p:e].decode(); p = e + 1
        size = struct.unpack_from("<i", b, p)[0]; p += 4
        attrs[name] = (typ, b[p:p + size]); p += size
    p += 1
    chans = []
    c = attrs["channels"][1]; q = 0
    while c[q] != 0:
        e = c.index(0, q); chans.append((c[q:e].decode(), struct.unpack_from("<i", c, e + 1)[0])); q = e + 1 + 16
    assert attrs["compression"][1] == b"\x00" and attrs["lineOrder"][1] == b"\x00"
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    W, H = x1 - x0 + 1, y1 - y0 + 1
    offs = struct.unpack_from("<%dQ" % H, b, p)
    img = {}
    for y in range(H):
        yy, size = struct.unpack_from("<ii", b, offs[y])
        assert yy == y and size == W * 4 * len(chans)
        row = np.frombuffer(b, "<f4", W * len(chans), offs[y] + 8).reshape(len(chans), W)
        for k, (nm, ty) in enumerate(chans):
            assert ty == 2
            img.setdefault(nm, np.zeros((H, W), np.float32))[y] = row[k]
    return img, [n for n, _ in chans]


def test_exr_writer_round_trip(V, tmp_path):
    """vkhrt_image_save_exr: header attributes an OpenEXR reader requires, channels in alphabetical order, floats bit-exact"""
    rng = np.random.default_rng(5)
    img = rng.standard_normal((7, 13, 4)).astype(np.float32)
    img[0, 0] = [np.inf, 0.0, -0.0, 1e-38]
    path = str(tmp_path / "frame.exr")
    V.save_exr(path, img)
    planes, names = _read_exr_uncompressed(path)
    assert names == ["A", "B", "G", "R"]
    for k, nm in enumerate("RGBA"):
        assert planes[nm].view(np.uint32).tolist() == img[:, :, k].view(np.uint32).tolist()
    with pytest.raises(V.VkhrtError):
        V.save_exr(str(tmp_path / "no" / "dir.exr"), img)


def test_oracle_material_albedo_term(O):
    """triangle_closest_hit.rchit:77-83: albedo = albedoFactor * texture(albedoMap, texCoord); hair has zero UVs, so the linear /
    repeat sampler returns the mean of the four corner texels.  Shade(n) = abs(n.y) * (0.4, 0.2, 0.1) + 0.3."""
    pos = np.array([[-1, 150, 0], [1, 150, 0]], np.float32)
    idx = np.array([[0, 1]], np.uint32)
    orc = O.OracleScene(pos, idx, technique=1)
    vi, pi = O.camera_matrices(aspect=1.0)
    tex = np.zeros((3, 5, 4), np.float32)
    tex[0, 0] = [1, 0, 0, 1]; tex[0, 4] = [0, 1, 0, 1]; tex[2, 0] = [0, 0, 1, 1]; tex[2, 4] = [1, 1, 1, 1]
    tex[1, 2] = [9, 9, 9, 9]                                   # interior texels are never touched at uv (0, 0)
    orc.set_material((0.5, 1.0, 0.25, 1.0), tex)
    W = H = 33
    _, img, _ = orc.render(O.make_frame(vi, pi, W, H, shade_mode=2))
    h, ref, _ = orc.render(O.make_frame(vi, pi, W, H, shade_mode=0))
    c = (H // 2) * W + W // 2
    assert h["flags"][c] & 1
    ny = abs(float(h["ny"][c]))
    shade = np.array([ny * 0.4 + 0.3, ny * 0.2 + 0.3, ny * 0.1 + 0.3])
    albedo = np.array([0.5 * 0.5, 1.0 * 0.5, 0.25 * 0.5])       # corner mean = (0.5, 0.5, 0.5)
    want = np.floor(np.clip(shade * albedo, 0, 1) * 255 + 0.5)
    assert np.abs(img[c, :3].astype(np.float64) - want).max() <= 1
    assert img[c, 3] == 255 and not np.array_equal(img, ref)
    orc.set_material()                                          # default material: albedo 1 -> plain Shade()
    _, img1, _ = orc.render(O.make_frame(vi, pi, W, H, shade_mode=2))
    assert np.array_equal(img1, ref)


def test_oracle_multi_mesh_scene(V, O):
    """One BLAS per mesh + one TLAS instance each in the reference (renderer.cpp:694-727) = concatenated line lists under one hierarchy here:
    the closest hit over the union is the nearer of the per-mesh closest hits, the hit segment names its mesh, and SHADE_MATERIAL
    multiplies Shade(n) by THAT mesh's albedo (hair_closest_hit.rchit:17-18 -> geometryNodes[...].material)."""
    a_pos = np.array([[-1, 150, 0], [1, 150, 0]], np.float32)                       # a horizontal hair through the image centre
    b_pos = np.array([[-1, 151, 0], [0, 151, 0], [1, 151, 0]], np.float32)          # two segments above it
    c_pos = np.array([[-1, 150, 5], [1, 150, 5]], np.float32)                       # in front of mesh 0: occludes it
    idx1 = np.array([[0, 1]], np.uint32); idx2 = np.array([[0, 1], [1, 2]], np.uint32)
    pos, idx, rad, first = V.merge_meshes([(a_pos, idx1), (b_pos, idx2), (c_pos, idx1)])
    assert list(first) == [0, 1, 3] and idx.tolist() == [[0, 1], [2, 3], [3, 4], [5, 6]] and rad is None
    vi, pi = O.camera_matrices(aspect=1.0)
    W = H = 65
    f0 = O.make_frame(vi, pi, W, H)
    for tech in (0, 1, 2):
        parts = [O.OracleScene(p, i, technique=tech, radius=0.2).render(f0)[0] for p, i in ((a_pos, idx1), (b_pos, idx2), (c_pos, idx1))]
        orc = O.OracleScene(pos, idx, technique=tech, radius=0.2)
        orc.set_meshes(first)
        albedo = [(1.0, 0.5, 0.25, 1.0), (0.25, 1.0, 0.5, 1.0), (0.5, 0.25, 1.0, 1.0)]
        for m, a in enumerate(albedo):
            orc.set_mesh_material(m, a)
        h, img, _ = orc.render(O.make_frame(vi, pi, W, H, shade_mode=2))
        _, plain, _ = orc.render(O.make_frame(vi, pi, W, H, shade_mode=0))
        t = np.stack([p["t"] for p in parts])                    # [3, rays]; +inf for a miss
        nearest = t.argmin(axis=0)
        hit = (h["flags"] & 1) != 0
        assert np.array_equal(hit, np.isfinite(t.min(axis=0)))
        assert np.array_equal(h["t"][hit], t.min(axis=0)[hit])                                    # union closest hit = nearest per-mesh hit
        mesh = np.array([orc.mesh_of_segment(int(s_)) for s_ in h["segment"][hit]])
        assert np.array_equal(mesh, nearest[hit])
        assert 0 not in mesh and {1, 2} <= set(mesh.tolist())                                        # mesh 2 hides mesh 0 everywhere
        local = h["segment"][hit] - first[mesh]                                                      # firstIndex-relative segment
        assert np.array_equal(local, np.array([parts[m]["segment"][k] for m, k in zip(mesh, np.flatnonzero(hit))]))
        for m in (1, 2):
            k = np.flatnonzero(hit)[mesh == m][0]
            ny = abs(float(h["ny"][k]))
            shade = np.array([ny * 0.4 + 0.3, ny * 0.2 + 0.3, ny * 0.1 + 0.3])
            want = np.floor(np.clip(shade * np.array(albedo[m][:3]), 0, 1) * 255 + 0.5)
            assert np.abs(img[k, :3].astype(np.float64) - want).max() <= 1
        orc.set_material()                                                                           # every mesh back to albedo 1
        _, img1, _ = orc.render(O.make_frame(vi, pi, W, H, shade_mode=2))
        assert np.array_equal(img1, plain)


def test_gltf_material_base_color(V, tmp_path):
    """ProcessMaterial (model_loader.cpp:96-99): albedoFactor = the glTF base colour of the line primitive's material"""
    import base64, json, struct
    pos = np.array([[0, 0, 0], [1, 0, 0], [2, 1, 0]], np.float32)
    idx = np.array([0, 1, 1, 2], np.uint32)
    blob = pos.tobytes() + idx.tobytes()
    doc = {"asset": {"version": "2.0"},
           "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 16}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}, {"bufferView": 1, "componentType": 5125, "count": 4, "type": "SCALAR"}],
           "materials": [{"name": "unused"}, {"pbrMetallicRoughness": {"baseColorFactor": [0.25, 0.5, 0.75, 1.0]}}],
           "meshes": [{"primitives": [{"mode": 1, "attributes": {"POSITION": 0}, "indices": 1, "material": 1}]}],
           "nodes": [{"mesh": 0}], "scenes": [{"nodes": [0]}], "scene": 0}
    p = tmp_path / "hair.gltf"
    p.write_text(json.dumps(doc))
    assert V.load_material(str(p)) == (0.25, 0.5, 0.75, 1.0)
    doc["meshes"][0]["primitives"][0].pop("material")
    p.write_text(json.dumps(doc))
    assert V.load_material(str(p)) == (1.0, 1.0, 1.0, 1.0)
    V.save_lines(str(tmp_path / "hair.obj"), pos, idx.reshape(-1, 2))
    assert V.load_material(str(tmp_path / "hair.obj")) == (1.0, 1.0, 1.0, 1.0)
