"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bars (BASELINE.json north_star): BVH build bit-exact; segment ids agree on
>= 99.9 % of rays; relative hit t <= 1e-4; image PSNR >= 45 dB.  Because both sides execute the same
fp32 operation sequence the tests additionally demand (and get) bit-identical hit records."""
import numpy as np
import pytest
from conftest import default_camera, psnr

pytestmark = pytest.mark.gpu

TECHS = [0, 1, 2]
TECH_NAMES = {0: "phantom", 1: "lss", 2: "dots"}


def compare_hits(hg, ho, min_seg_agree=0.999, t_rel=1e-4):
    assert hg.shape == ho.shape
    both = (hg["flags"] & 1).astype(bool) & (ho["flags"] & 1).astype(bool)
    agree = np.mean(hg["segment"] == ho["segment"])
    assert agree >= min_seg_agree, f"segment agreement {agree}"
    same = both & (hg["segment"] == ho["segment"])
    if same.any():
        rel = np.abs(hg["t"][same] - ho["t"][same]) / np.abs(ho["t"][same])
        assert rel.max() <= t_rel, f"max rel t err {rel.max()}"
    return agree


def assert_bit_identical(hg, ho):
    bad = np.nonzero(hg.view(np.uint8).reshape(-1, 32) != ho.view(np.uint8).reshape(-1, 32))[0]
    assert bad.size == 0, f"{np.unique(bad).size} hit records differ, first {hg[bad[0]]} vs {ho[bad[0]]}"


@pytest.fixture(scope="module")
def small_groom(V):
    return V.generate_groom(2000, 16, V.GROOM_CURLY)


@pytest.mark.parametrize("tech", TECHS)
def test_primitives_and_bvh_bit_exact(V, O, small_groom, tech):
    pos, idx = small_groom
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=tech)
        assert sc.n_primitives == orc.n_primitives
        assert np.array_equal(sc.primitives().view(np.uint32), orc.primitives().view(np.uint32))
        nodes, ids, morton, lohi = sc.bvh()
        onodes, oids, omorton, olohi = orc.bvh()
        assert np.array_equal(lohi.view(np.uint32), olohi.view(np.uint32))
        assert np.array_equal(morton, omorton)
        assert np.array_equal(ids, oids)
        assert nodes.tobytes() == onodes.tobytes()


@pytest.mark.parametrize("tech", TECHS)
@pytest.mark.parametrize("mode", [0, 1])
def test_render_parity_small(V, O, small_groom, tech, mode):
    pos, idx = small_groom
    W, H = 320, 200
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=tech)
        fg = V.make_frame(vi, pi, W, H, shade_mode=mode, miss_rgb=(0.1, 0.2, 0.3))
        fo = O.make_frame(vi, pi, W, H, shade_mode=mode, miss_rgb=(0.1, 0.2, 0.3))
        hg, ig, _ = sc.render(fg)
        ho, io, _ = orc.render(fo)
        assert (ho["flags"] & 1).sum() > 1000
        compare_hits(hg, ho)
        assert_bit_identical(hg, ho)
        assert psnr(ig, io) >= 45.0
        assert np.array_equal(ig, io)


def test_config1_straight_phantom_full_frame(V, O):
    """BASELINE config[0]: straight groom 10k x 16, 512x512, Phantom, hit buffer only."""
    pos, idx = V.generate_groom(10000, 16, V.GROOM_STRAIGHT)
    W = H = 512
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=V.PHANTOM) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=0)
        hg, _, _ = sc.render(V.make_frame(vi, pi, W, H), rgba=False)
        ho, _, _ = orc.render(O.make_frame(vi, pi, W, H), rgba=False)
        assert (ho["flags"] & 1).mean() > 0.05
        compare_hits(hg, ho)
        assert_bit_identical(hg, ho)


@pytest.mark.parametrize("tech", TECHS)
def test_spp_and_sample0_hits(V, O, small_groom, tech):
    pos, idx = small_groom
    W, H = 160, 96
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=tech)
        hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H, spp=4))
        ho, io, _ = orc.render(O.make_frame(vi, pi, W, H, spp=4))
        assert_bit_identical(hg, ho)
        assert np.array_equal(ig, io)
        h1, _, _ = sc.render(V.make_frame(vi, pi, W, H, spp=1), rgba=False)
        assert_bit_identical(h1, hg)   # the hit buffer is sample 0 = pixel centre


@pytest.mark.parametrize("world", [2, 3])
def test_tile_shards_and_untile(V, O, small_groom, world):
    import torch
    pos, idx = small_groom
    W, H, T = 200, 120, 32   # partial tiles on both axes
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=V.PHANTOM) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=0)
        full_h, full_i, _ = sc.render(V.make_frame(vi, pi, W, H, tile_size=T))
        shards_h, shards_i = [], []
        for r in range(world):
            fg = V.make_frame(vi, pi, W, H, tile_size=T, tile_first=r, tile_stride=world)
            fo = O.make_frame(vi, pi, W, H, tile_size=T, tile_first=r, tile_stride=world)
            hg, ig, _ = sc.render(fg)
            ho, io, _ = orc.render(fo, n_out=V.frame_local_pixels(fg))
            assert_bit_identical(hg, ho)
            assert np.array_equal(ig, io)
            shards_h.append(hg); shards_i.append(ig)
        gh = torch.from_numpy(np.concatenate(shards_h).view(np.uint8).reshape(-1, 32)).cuda()
        gi = torch.from_numpy(np.concatenate(shards_i)).cuda()
        oh = torch.empty((W * H, 32), dtype=torch.uint8, device="cuda")
        oi = torch.empty((W * H, 4), dtype=torch.uint8, device="cuda")
        f = V.make_frame(vi, pi, W, H, tile_size=T)
        st = torch.cuda.current_stream().cuda_stream
        V.untile(f, world, gh.data_ptr(), oh.data_ptr(), 32, st)
        V.untile(f, world, gi.data_ptr(), oi.data_ptr(), 4, st)
        torch.cuda.synchronize()
        assert np.array_equal(oh.cpu().numpy().reshape(-1).view(V.HIT_DTYPE).view(np.uint8), full_h.view(np.uint8))
        assert np.array_equal(oi.cpu().numpy(), full_i)


@pytest.mark.parametrize("tech,spp,rgba", [(0, 1, False), (1, 2, True), (2, 1, True)])
def test_frames_in_flight(V, small_groom, tech, spp, rgba):
    """vkhrt_render_submit / vkhrt_render_wait (Renderer::Render keeps frames in flight behind fences, renderer.cpp:85-119): a moving
    camera, two page-locked output sets used alternately, at most 2 frames outstanding; every frame must equal the blocking call's."""
    import torch
    pos, idx = small_groom
    W, H = 1280 if tech == 0 else 320, 720 if tech == 0 else 200          # the Phantom frame is large enough for the pool kernel
    cams = [V.camera_matrices(position=(0.3 * k, 150.0 + 0.2 * k, 20.0 - 0.5 * k), yaw=-90.0 + 2 * k, aspect=float(np.float32(W) / np.float32(H))) for k in range(5)]
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        refs = [sc.render(V.make_frame(vi, pi, W, H, spp=spp), rgba=rgba) for vi, pi in cams]
        hh = [torch.zeros((W * H, 32), dtype=torch.uint8).pin_memory() for _ in range(2)]
        ii = [torch.zeros((W * H, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
        got = []
        for k, (vi, pi) in enumerate(cams):
            if k >= 2:                      # frame k-2 used this buffer set: take its result before it is overwritten
                sc.wait()
                got.append((hh[k % 2].numpy().copy(), ii[k % 2].numpy().copy()))
            sc.submit(V.make_frame(vi, pi, W, H, spp=spp, output_memory=V.MEM_HOST), hh[k % 2].data_ptr(), ii[k % 2].data_ptr() if rgba else None)
        for k in (3, 4):
            sc.wait()
            got.append((hh[k % 2].numpy().copy(), ii[k % 2].numpy().copy()))
        with pytest.raises(V.VkhrtError):
            sc.wait()                       # nothing outstanding
        for k, ((gh, gi), (rh, ri, _)) in enumerate(zip(got, refs)):
            assert gh.reshape(-1).tobytes() == rh.tobytes(), k
            if rgba:
                assert np.array_equal(gi, ri), k
        # three submits without a wait: the third waits for the first by itself; frames still complete in order
        for k in range(3):
            sc.submit(V.make_frame(cams[k][0], cams[k][1], W, H, spp=spp, output_memory=V.MEM_HOST), hh[k % 2].data_ptr(), None)
        sc.wait(); sc.wait()
        assert hh[1].numpy().reshape(-1).tobytes() == refs[1][0].tobytes() and hh[0].numpy().reshape(-1).tobytes() == refs[2][0].tobytes()


def test_untile_with_one_shard_is_a_plain_copy(V, small_groom):
    """vkhrt_untile(world = 1): a one-shard render is never compact (tile_stride 1 writes row-major, W*H records), so the device
    untile must be a plain copy like vkhrt_untile_host — on a frame whose size is not a multiple of the tile size (ADVICE r1:
    the tiled index arithmetic scrambled the image and read past the W*H source)."""
    import torch
    pos, idx = small_groom
    W, H, T = 203, 117, 32
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=V.LSS) as sc:
        sc.build()
        h, img, _ = sc.render(V.make_frame(vi, pi, W, H, tile_size=T))
    f = V.make_frame(vi, pi, W, H, tile_size=T)
    st = torch.cuda.current_stream().cuda_stream
    for arr, elem in ((h.view(np.uint8).reshape(-1, 32), 32), (img, 4)):
        src = torch.from_numpy(np.ascontiguousarray(arr)).cuda()              # exactly W*H elements: nothing to read beyond
        dst = torch.zeros_like(src)
        V.untile(f, 1, src.data_ptr(), dst.data_ptr(), elem, st)
        torch.cuda.synchronize()
        assert torch.equal(src, dst)
        assert np.array_equal(V.untile_host(f, 1, arr.reshape(-1).view(V.HIT_DTYPE) if elem == 32 else arr).view(np.uint8).reshape(-1, elem), arr)


def test_candidate_filter_is_neutral_for_thin_long_segments(V, O):
    """The quarter-chord candidate filter (not in the reference) must never drop a hit Prhi reports.  Its bound has to cover the
    convergence tolerance: an accepted hit lies sqrt(r^2 + (5e-5 |B'|)^2) from B(t), which matters when segments are long and thin
    (length / radius beyond ~900).  Radius 0.001 and 0.0004 on 1.5-unit segments (L / r = 1500 and 3750), grazing rays
    included by the dense frame; the oracle has no filter."""
    pos, idx = V.generate_groom(3000, 4, V.GROOM_CURLY)          # 6 units of strand in 4 segments
    W, H = 640, 400
    vi, pi = default_camera(V, W, H)
    for r in (0.001, 0.0004):
        with V.Scene(pos, idx, technique=V.PHANTOM, radius=r) as sc:
            sc.build()
            hg, _, _ = sc.render(V.make_frame(vi, pi, W, H), rgba=False)
        ho, _, _ = O.OracleScene(pos, idx, technique=0, radius=r).render(O.make_frame(vi, pi, W, H), rgba=False)
        assert (ho["flags"] & 1).sum() > 500
        assert_bit_identical(hg, ho)


def test_device_outputs_and_wavefront_api(V, O, small_groom):
    import torch
    pos, idx = small_groom
    W, H = 128, 64
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=V.PHANTOM) as sc:
        sc.build()
        href, iref, _ = sc.render(V.make_frame(vi, pi, W, H))
        st = torch.cuda.current_stream().cuda_stream
        dh = torch.zeros((W * H, 32), dtype=torch.uint8, device="cuda")
        di = torch.zeros((W * H, 4), dtype=torch.uint8, device="cuda")
        f = V.make_frame(vi, pi, W, H, output_memory=V.MEM_DEVICE, stream=st)
        sc.render_into(f, dh.data_ptr(), di.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(dh.cpu().numpy().reshape(-1), href.view(np.uint8))
        assert np.array_equal(di.cpu().numpy(), iref)
        # wavefront: ray buffer from the generator kernel, traced by vkhrt_trace_rays
        rays = torch.zeros((W * H, 8), dtype=torch.float32, device="cuda")
        from vkhrt_b200.api import generate_rays
        generate_rays(f, 0, rays.data_ptr(), 0)
        torch.cuda.synchronize()
        o, d = O.raygen(vi, pi, W, H, 5, 7)
        r = rays[7 * W + 5].cpu().numpy()
        assert np.array_equal(r[:3], o) and np.array_equal(r[4:7], d)
        dh2 = torch.zeros_like(dh)
        sc.trace_rays(rays.data_ptr(), W * H, dh2.data_ptr(), st)
        torch.cuda.synchronize()
        assert np.array_equal(dh2.cpu().numpy().reshape(-1), href.view(np.uint8))
        # oracle on the very same ray buffer
        ho = O.OracleScene(pos, idx, technique=0).trace_rays(rays.cpu().numpy())
        assert np.array_equal(ho.view(np.uint8), href.view(np.uint8))


@pytest.mark.parametrize("tech", TECHS)
def test_refit_equals_rebuild_results(V, O, small_groom, tech):
    pos, idx = small_groom
    W, H = 160, 100
    vi, pi = default_camera(V, W, H)
    rng = np.random.default_rng(7)
    strand = rng.normal(0, 0.05, (2000, 1, 3)).astype(np.float32)
    pos2 = (pos.reshape(2000, 17, 3) + strand * np.linspace(0, 1, 17, dtype=np.float32)[None, :, None]).reshape(-1, 3)
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        sc.refit(pos2)
        hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H))
        orc = O.OracleScene(pos2, idx, technique=tech)
        ho, io, _ = orc.render(O.make_frame(vi, pi, W, H))
        # a refitted tree is a different (valid) tree than a rebuilt one: results must still agree
        compare_hits(hg, ho)
        assert_bit_identical(hg, ho)
        assert np.array_equal(sc.primitives().view(np.uint32), orc.primitives().view(np.uint32))


def test_edge_cases(V, O):
    W, H = 64, 48
    vi, pi = default_camera(V, W, H)
    # empty scene: every ray misses
    with V.Scene(np.zeros((0, 3), np.float32), np.zeros((0, 2), np.uint32)) as sc:
        sc.build()
        h, img, _ = sc.render(V.make_frame(vi, pi, W, H, miss_rgb=(1.0, 0.5, 0.0)))
        assert (h["flags"] == 0).all() and np.isinf(h["t"]).all() and (h["segment"] == 0xFFFFFFFF).all()
        assert (img == np.array([255, 128, 0, 255], np.uint8)).all()
    # one segment in front of the camera, all techniques; ragged strands (1-, 2-, 5-segment) and a
    # reversed-index strand (connectivity by position equality, not index order)
    pos = np.array([[-1, 150, 0], [1, 150, 0],
                    [-2, 151, 0], [-1, 151.2, 0], [0, 151, 0],
                    [-3, 149, 1], [-2, 149.1, 1], [-1, 149, 1], [0, 149.1, 1], [1, 149, 1], [2, 149.1, 1],
                    [3, 152, 0], [2, 152, 0]], np.float32)
    idx = np.array([[0, 1], [2, 3], [3, 4], [5, 6], [6, 7], [7, 8], [8, 9], [9, 10], [11, 12]], np.uint32)
    vi, pi = V.camera_matrices(position=(0.0, 150.5, 4.0), aspect=float(np.float32(W) / np.float32(H)))   # close-up: hairs are > 1 px wide
    for tech in TECHS:
        for p, i in ((pos[:2], idx[:1]), (pos, idx)):
            with V.Scene(p, i, technique=tech, radius=0.1) as sc:   # non-default radius, > 1 px wide
                sc.build()
                orc = O.OracleScene(p, i, technique=tech, radius=0.1)
                hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H))
                ho, io, _ = orc.render(O.make_frame(vi, pi, W, H))
                assert (ho["flags"] & 1).sum() > 0
                assert_bit_identical(hg, ho)
                assert np.array_equal(ig, io)
                n, ids, m, lohi = sc.bvh()
                on, oids, om, olohi = orc.bvh()
                assert n.tobytes() == on.tobytes() and np.array_equal(ids, oids)
    # bad topology is an error, not a crash
    with pytest.raises(V.VkhrtError):
        V.Scene(pos[:2], np.array([[0, 5]], np.uint32))


def test_lss_per_vertex_radius(V, O, small_groom):
    pos, idx = small_groom
    W, H = 200, 120
    vi, pi = default_camera(V, W, H)
    taper = np.tile(np.linspace(0.02, 0.005, 17, dtype=np.float32), 2000)
    with V.Scene(pos, idx, technique=V.LSS, radius_per_vertex=taper) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=1, radius_per_vertex=taper)
        hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H))
        ho, io, _ = orc.render(O.make_frame(vi, pi, W, H))
        assert_bit_identical(hg, ho)
        assert np.array_equal(ig, io)


@pytest.mark.parametrize("tech", TECHS)
def test_per_vertex_radius_all_techniques(V, O, small_groom, tech):
    """north_star input contract: polyline strands with PER-VERTEX radius.  Taper 0.02 -> 0.005 along every strand:
    Phantom = radius(t) + cone slant (cone.glsl:27), DOTS = per-end offsets, LSS = per-end sphere radii.
    Primitives, BVH nodes, hit records, images and counters against the oracle; and a constant per-vertex array must give
    exactly the default-radius scene."""
    pos, idx = small_groom
    W, H = 256, 160
    vi, pi = default_camera(V, W, H)
    taper = np.tile(np.linspace(0.02, 0.005, 17, dtype=np.float32), 2000)
    with V.Scene(pos, idx, technique=tech, radius_per_vertex=taper) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=tech, radius_per_vertex=taper)
        assert np.array_equal(sc.primitives().view(np.uint32), orc.primitives().view(np.uint32))
        assert sc.bvh()[0].tobytes() == orc.bvh()[0].tobytes()
        for spp, ao in ((1, 0), (2, 2)):
            hg, ig, sg = sc.render(V.make_frame(vi, pi, W, H, spp=spp, ao_samples=ao), stats=True)
            ho, io, so = orc.render(O.make_frame(vi, pi, W, H, spp=spp, ao_samples=ao), stats=True)
            assert (ho["flags"] & 1).sum() > 1000
            assert_bit_identical(hg, ho)
            assert np.array_equal(ig, io)
            assert sg["nodes_visited"] == so["nodes_visited"] and sg["prims_tested"] == so["prims_tested"]
        # thinner hair is hit less often than the default 0.02 everywhere
        with V.Scene(pos, idx, technique=tech) as ref:
            ref.build()
            h_ref, i_ref, _ = ref.render(V.make_frame(vi, pi, W, H))
        h_tap, _, _ = sc.render(V.make_frame(vi, pi, W, H))
        assert (h_tap["flags"] & 1).sum() < (h_ref["flags"] & 1).sum()
    const = np.full(pos.shape[0], 0.02, np.float32)
    with V.Scene(pos, idx, technique=tech, radius_per_vertex=const) as sc:
        sc.build()
        hc, ic, _ = sc.render(V.make_frame(vi, pi, W, H))
        assert hc.tobytes() == h_ref.tobytes() and np.array_equal(ic, i_ref)


@pytest.mark.parametrize("tech", TECHS)
def test_material_albedo_shading(V, O, small_groom, tech):
    """shade_mode MATERIAL: Shade(normal) * albedoFactor * texture(albedoMap, (0,0)) (triangle_closest_hit.rchit:77-83)"""
    pos, idx = small_groom
    W, H = 200, 120
    vi, pi = default_camera(V, W, H)
    rng = np.random.default_rng(3)
    tex = rng.random((4, 6, 4)).astype(np.float32)
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=tech)
        for factor, m in (((0.8, 0.5, 0.3, 1.0), None), ((1.0, 0.9, 0.7, 1.0), tex)):
            sc.set_material(factor, m); orc.set_material(factor, m)
            for spp in (1, 2):
                hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H, spp=spp, shade_mode=V.SHADE_MATERIAL, miss_rgb=(0.2, 0.1, 0.0)))
                ho, io, _ = orc.render(O.make_frame(vi, pi, W, H, spp=spp, shade_mode=2, miss_rgb=(0.2, 0.1, 0.0)))
                assert_bit_identical(hg, ho)
                assert np.array_equal(ig, io)
        _, plain, _ = sc.render(V.make_frame(vi, pi, W, H, shade_mode=V.SHADE))
        assert not np.array_equal(plain, ig)
        sc.set_material()
        _, unit, _ = sc.render(V.make_frame(vi, pi, W, H, shade_mode=V.SHADE_MATERIAL))
        assert np.array_equal(unit, plain)                       # the default material leaves Shade() unchanged


@pytest.mark.parametrize("tech", TECHS)
def test_multi_mesh_scene(V, O, tech):
    """Several meshes in one scene (the reference: one BLAS per mesh / hair, one TLAS instance each, renderer.cpp:694-727): the line lists are
    concatenated (firstVertex / firstIndex, geometry_processor.cpp:45-67) under ONE LBVH.  Hit records equal the oracle's and those of the
    same segments given as a single mesh; the record's segment maps back to its mesh (= gl_InstanceCustomIndexEXT); SHADE_MATERIAL uses
    the material of the mesh that was hit."""
    p1, i1 = V.generate_groom(700, 10, V.GROOM_CURLY)
    p2, i2 = V.generate_groom(500, 7, V.GROOM_STRAIGHT, seed=11)
    p3, i3 = V.generate_groom(300, 5, V.GROOM_CURLY, seed=5)
    p2 = p2 + np.float32([4.0, -2.0, 1.0]); p3 = p3 + np.float32([-5.0, 1.0, 2.0])
    pos, idx, rad, first = V.merge_meshes([(p1, i1), (p2, i2), (p3, i3)])
    assert rad is None and list(first) == [0, len(i1), len(i1) + len(i2)]
    W, H = 240, 150
    vi, pi = default_camera(V, W, H)
    mats = [((0.9, 0.4, 0.2, 1.0), None), ((0.2, 0.8, 0.5, 1.0), np.random.default_rng(1).random((3, 5, 4)).astype(np.float32)), ((0.5, 0.5, 1.0, 1.0), None)]
    with V.Scene(pos, idx, technique=tech) as sc, V.Scene(pos, idx, technique=tech) as single:
        sc.set_meshes(first).build()
        single.build()
        assert sc.n_meshes == 3 and single.n_meshes == 1
        orc = O.OracleScene(pos, idx, technique=tech)
        orc.set_meshes(first)
        for m, (factor, tex) in enumerate(mats):
            sc.set_mesh_material(m, factor, tex); orc.set_mesh_material(m, factor, tex)
        for spp in (1, 3):
            hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H, spp=spp, shade_mode=V.SHADE_MATERIAL, miss_rgb=(0.1, 0.1, 0.2)))
            ho, io, _ = orc.render(O.make_frame(vi, pi, W, H, spp=spp, shade_mode=2, miss_rgb=(0.1, 0.1, 0.2)))
            assert_bit_identical(hg, ho)
            assert np.array_equal(ig, io)
        hs, _, _ = single.render(V.make_frame(vi, pi, W, H))
        h1, _, _ = sc.render(V.make_frame(vi, pi, W, H))
        assert_bit_identical(h1, hs)                                   # the mesh table changes no record
        mesh = sc.mesh_of_segments(h1["segment"])
        hit = (h1["flags"] & 1) != 0
        assert np.all(mesh[~hit] == 0xFFFFFFFF)
        want = np.searchsorted(first, h1["segment"][hit], side="right") - 1
        assert np.array_equal(mesh[hit], want.astype(np.uint32))
        assert set(np.unique(mesh[hit])) == {0, 1, 2}                  # the camera sees all three
        assert all(orc.mesh_of_segment(int(s_)) == int(m_) for s_, m_ in zip(h1["segment"][hit][:200], mesh[hit][:200]))
        # one material for the whole scene again
        sc.set_material((0.3, 0.6, 0.9, 1.0)); single.set_material((0.3, 0.6, 0.9, 1.0))
        _, ia, _ = sc.render(V.make_frame(vi, pi, W, H, shade_mode=V.SHADE_MATERIAL))
        _, ib, _ = single.render(V.make_frame(vi, pi, W, H, shade_mode=V.SHADE_MATERIAL))
        assert np.array_equal(ia, ib)
        # a table of ONE mesh: its material is the scene's
        sc.set_meshes([0]).set_mesh_material(0, (0.7, 0.2, 0.9, 1.0)); single.set_material((0.7, 0.2, 0.9, 1.0))
        orc.set_meshes([0]); orc.set_mesh_material(0, (0.7, 0.2, 0.9, 1.0))
        _, ia, _ = sc.render(V.make_frame(vi, pi, W, H, shade_mode=V.SHADE_MATERIAL))
        _, ib, _ = single.render(V.make_frame(vi, pi, W, H, shade_mode=V.SHADE_MATERIAL))
        _, ic, _ = orc.render(O.make_frame(vi, pi, W, H, shade_mode=2))
        assert sc.n_meshes == 1 and np.array_equal(ia, ib) and np.array_equal(ia, ic)
        # the table must start at 0 and ascend; LOD passes renumber segments
        with pytest.raises(V.VkhrtError):
            sc.set_meshes([1, 5])
        with pytest.raises(V.VkhrtError):
            sc.set_meshes([0, 9, 3])
        with pytest.raises(V.VkhrtError):
            sc.set_mesh_material(7)
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.set_meshes(first)
        with pytest.raises(V.VkhrtError):
            sc.apply_lod(1, 0, 0)


def test_multi_mesh_with_per_vertex_radii_and_refit(V, O):
    """meshes with their own per-vertex radii (merge_meshes concatenates them), then a refit of the whole scene: records and the per-mesh
    albedo image stay equal to the oracle's"""
    p1, i1 = V.generate_groom(500, 9, V.GROOM_CURLY)
    p2, i2 = V.generate_groom(400, 6, V.GROOM_CURLY, seed=21)
    p2 = p2 + np.float32([3.0, 1.0, 0.0])
    r1 = np.tile(np.linspace(0.03, 0.008, 10, dtype=np.float32), 500)
    r2 = np.full(p2.shape[0], 0.015, np.float32)
    pos, idx, rad, first = V.merge_meshes([(p1, i1, r1), (p2, i2, r2)])
    assert rad.shape[0] == pos.shape[0]
    with pytest.raises(ValueError):
        V.merge_meshes([(p1, i1, r1), (p2, i2)])
    W, H = 200, 128
    vi, pi = default_camera(V, W, H)
    for tech in TECHS:
        with V.Scene(pos, idx, technique=tech, radius_per_vertex=rad) as sc:
            sc.set_meshes(first).set_mesh_material(0, (1.0, 0.6, 0.3, 1.0)).set_mesh_material(1, (0.3, 0.6, 1.0, 1.0)).build()
            moved = pos + np.float32([0.05, -0.02, 0.03])
            sc.refit(moved)
            orc = O.OracleScene(moved, idx, technique=tech, radius_per_vertex=rad)
            orc.set_meshes(first); orc.set_mesh_material(0, (1.0, 0.6, 0.3, 1.0)); orc.set_mesh_material(1, (0.3, 0.6, 1.0, 1.0))
            hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H, spp=2, shade_mode=V.SHADE_MATERIAL))
            ho, io, _ = orc.render(O.make_frame(vi, pi, W, H, spp=2, shade_mode=2))
            assert_bit_identical(hg, ho)
            assert np.array_equal(ig, io)
            hit = (hg["flags"] & 1) != 0
            assert set(np.unique(sc.mesh_of_segments(hg["segment"][hit]))) == {0, 1}


def test_stats_counters_equal_the_oracles(V, O, small_groom):
    """The warp scheduler only interleaves lanes; each ray's own sequence of node visits and candidate tests is the
    oracle's, so the traversal counters (the N_int / N_prim of the bytes-per-ray roofline) are IDENTICAL."""
    pos, idx = small_groom
    W, H = 256, 160
    vi, pi = default_camera(V, W, H)
    for tech in TECHS:
        with V.Scene(pos, idx, technique=tech) as sc:
            sc.build()
            _, _, sg = sc.render(V.make_frame(vi, pi, W, H), rgba=False, stats=True)
            _, _, so = O.OracleScene(pos, idx, technique=tech).render(O.make_frame(vi, pi, W, H), rgba=False, stats=True)
            assert sg["rays"] == so["rays"] == W * H
            assert sg["hits"] == so["hits"]
            assert sg["nodes_visited"] == so["nodes_visited"]
            assert sg["prims_tested"] == so["prims_tested"]
            if tech == 0:
                # fewer cone iterations than the reference loop: fixed-point exit + conservative candidate filter
                assert 0 < sg["phantom_iterations"] <= so["phantom_iterations"]
            steps, lanes = sg["sched_steps"], sg["sched_lanes"]
            assert steps[0] > 0 and all(l <= 32 * s for s, l in zip(steps, lanes))


def test_full_size_config2_properties(V, O):
    """BASELINE config[1] at full size (3.2 M segments, 1080p): size-independent properties +
    an oracle spot check on a seeded 8192-pixel subset."""
    pos, idx = V.generate_groom(100000, 32, V.GROOM_CURLY)
    W, H = 1920, 1080
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=V.PHANTOM) as sc:
        sc.build()
        nodes, ids, morton, _ = sc.bvh()
        assert sc.n_primitives == 3200000
        n = ids.shape[0]                                                                # BVH leaves = 2 pieces per curve
        assert n == 2 * 3200000 and nodes.shape[0] == n - 1 and sc.n_leaves == n
        assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))          # a permutation
        assert (np.diff(morton.astype(np.int64)) >= 0).all()                            # sortedness
        leaf0 = nodes["child0"] >> 31 == 1; leaf1 = nodes["child1"] >> 31 == 1
        assert leaf0.sum() + leaf1.sum() == n                                           # every leaf referenced once
        leaves = np.concatenate([nodes["child0"][leaf0], nodes["child1"][leaf1]]) & 0x7FFFFFFF
        assert np.array_equal(np.sort(leaves), np.arange(n, dtype=np.uint32))
        h1, _, _ = sc.render(V.make_frame(vi, pi, W, H), rgba=False)
        h2, _, _ = sc.render(V.make_frame(vi, pi, W, H), rgba=False)
        assert h1.tobytes() == h2.tobytes()                                             # idempotent / deterministic
        hit = (h1["flags"] & 1).astype(bool)
        assert 0.05 < hit.mean() < 0.9
        nrm = np.sqrt(h1["nx"][hit] ** 2 + h1["ny"][hit] ** 2 + h1["nz"][hit] ** 2)
        assert np.abs(nrm - 1).max() < 1e-5 and (h1["u"][hit] >= 0).all() and (h1["u"][hit] <= 1).all()
        rng = np.random.default_rng(0x5EED)
        sub = np.sort(rng.choice(W * H, 8192, replace=False)).astype(np.uint64)
        orc = O.OracleScene(pos, idx, technique=0)
        ho, _, _ = orc.render(O.make_frame(vi, pi, W, H), rgba=False, pixel_subset=sub)
        assert_bit_identical(h1[sub.astype(np.int64)], ho)


@pytest.mark.parametrize("tech", TECHS)
def test_pinned_host_buffers_zero_copy_path(V, small_groom, tech):
    """vkhrt_render with a PINNED host hit buffer: the kernel stores the records straight into host memory
    (no device->host copy).  Must equal the pageable-buffer path bit for bit, with and without an image."""
    import torch
    pos, idx = small_groom
    W, H = 200, 120
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        # spp 1 with an image: records mirrored by the kernel (HBM copy for the shading pass); spp > 1 with an image: the records of
        # sample 0 leave on the copy engine while the other samples are traced
        for spp in (1, 2, 5):
            href, iref, _ = sc.render(V.make_frame(vi, pi, W, H, spp=spp))
            for want_rgba in (False, True):
                hh = torch.zeros((W * H, 32), dtype=torch.uint8).pin_memory()
                hi = torch.zeros((W * H, 4), dtype=torch.uint8).pin_memory()
                f = V.make_frame(vi, pi, W, H, spp=spp, output_memory=V.MEM_HOST)
                for _ in range(2):                                                              # twice: the scratch buffers are reused
                    hh.zero_()
                    sc.render_into(f, hh.data_ptr(), hi.data_ptr() if want_rgba else None)      # blocks until complete
                    assert np.array_equal(hh.numpy().reshape(-1), href.view(np.uint8))
                    if want_rgba:
                        assert np.array_equal(hi.numpy(), iref)


@pytest.mark.parametrize("W,H,shard", [(201, 121, None), (333, 77, None), (200, 120, (1, 3)), (1920, 1080, None), (1928, 1083, (0, 2))])
def test_pinned_host_buffer_awkward_sizes_and_line_wise_delivery(V, small_groom, W, H, shard):
    """Phantom hit records into a pinned buffer for record counts that are not a multiple of 4, compact shards with padding
    records, and a frame large enough for the pool kernel: there the records reach the host line-wise (a 128-byte line = 4
    records is copied by four lanes when its last record is written).  The nested runs of this file force that path onto the
    small frames too.  Must equal the pageable-buffer path bit for bit."""
    import torch
    pos, idx = small_groom
    vi, pi = default_camera(V, W, H)
    kw = {} if shard is None else dict(tile_size=32, tile_first=shard[0], tile_stride=shard[1])
    with V.Scene(pos, idx, technique=V.PHANTOM) as sc:
        sc.build()
        href, _, _ = sc.render(V.make_frame(vi, pi, W, H, **kw), rgba=False)
        n = href.shape[0]
        for rep in range(2):                                   # the per-line counters must be reset between frames
            hh = torch.full((n, 32), 0xAB, dtype=torch.uint8).pin_memory()
            sc.render_into(V.make_frame(vi, pi, W, H, output_memory=V.MEM_HOST, **kw), hh.data_ptr(), None)
            got = hh.numpy().reshape(-1)
            assert np.array_equal(got, href.view(np.uint8)), f"{int((got != href.view(np.uint8)).sum())} bytes differ (rep {rep})"


@pytest.mark.parametrize("tech,spp", [(0, 1), (1, 3), (2, 1)])
def test_row_major_shards_share_one_frame_buffer(V, O, small_groom, tech, spp):
    """tile_stride > 1 with row_major_output: every shard writes only its own pixels at their row-major position of a
    full-frame buffer (the layout the multi-GPU 'peer' mode uses over NVLink).  All shards into ONE buffer == full frame."""
    import torch
    pos, idx = small_groom
    W, H, T, world = 200, 120, 32, 3
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        full_h, full_i, _ = sc.render(V.make_frame(vi, pi, W, H, spp=spp, tile_size=T))
        shared = V.SharedBuffer.create(W * H * 32)                       # an IPC-exportable device buffer
        try:
            di = torch.full((W * H, 4), 7, dtype=torch.uint8, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            for r in range(world):
                f = V.make_frame(vi, pi, W, H, spp=spp, tile_size=T, tile_first=r, tile_stride=world, row_major_output=1,
                                 output_memory=V.MEM_DEVICE, stream=st)
                assert V.frame_local_pixels(f) == W * H
                sc.render_into(f, shared.ptr, di.data_ptr())
                torch.cuda.synchronize()
                if r == 0:      # only this shard's tiles have been written so far
                    assert (di.cpu().numpy() == 7).all(axis=1).sum() > W * H // 2
            from vkhrt_b200.multi import _DeviceView
            dh = torch.as_tensor(_DeviceView(shared.ptr, (W * H, 32)), device="cuda")
            assert np.array_equal(dh.cpu().numpy().reshape(-1), full_h.view(np.uint8))
            assert np.array_equal(di.cpu().numpy(), full_i)
        finally:
            shared.close()
        # the oracle implements the same layout
        orc = O.OracleScene(pos, idx, technique=tech)
        acc = np.zeros(W * H, V.HIT_DTYPE)
        for r in range(world):
            fo = O.make_frame(vi, pi, W, H, spp=spp, tile_size=T, tile_first=r, tile_stride=world, row_major_output=1)
            ho, _, _ = orc.render(fo, rgba=False)
            own = (ho["flags"] != 0) | (ho["t"] != 0)
            acc[own] = ho[own]
        assert acc.tobytes() == full_h.tobytes()
        # host output memory cannot be shared between shards
        with pytest.raises(V.VkhrtError):
            sc.render(V.make_frame(vi, pi, W, H, tile_size=T, tile_first=0, tile_stride=2, row_major_output=1))


@pytest.mark.parametrize("cam,radius,fov", [((0.0, 150.0, 20.0), 0.02, 60.0), ((0.0, 150.0, 400.0), 0.02, 4.0),
                                            ((3.0, 158.5, 1.0), 0.005, 90.0), ((900.0, 150.0, 0.0), 0.1, 2.0)])
def test_dots_strip_reject_is_result_neutral(V, O, small_groom, cam, radius, fov):
    """DOTS leaves are 4-triangle strips and the kernel rejects a strip when the ray passes farther than r (inflated) from
    the segment axis before running the 4 triangle tests.  The oracle has no such filter: bit-identical hit records from
    near, far, grazing and inside-the-groom cameras (and from scaled, non-unit wavefront directions) show it never
    removes a hit."""
    import torch
    pos, idx = small_groom
    W, H = 192, 128
    yaw = 180.0 if cam[0] > 100 else -90.0
    vi, pi = V.camera_matrices(position=cam, yaw=yaw, fov=fov, aspect=float(np.float32(W) / np.float32(H)))
    with V.Scene(pos, idx, technique=V.DOTS, radius=radius) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=2, radius=radius)
        hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H))
        ho, io, _ = orc.render(O.make_frame(vi, pi, W, H))
        assert (ho["flags"] & 1).sum() > 200
        assert_bit_identical(hg, ho)
        assert np.array_equal(ig, io)
        # wavefront rays with directions scaled by 0.25 .. 4 (the reject must not assume |d| = 1)
        from vkhrt_b200.api import generate_rays
        st = torch.cuda.current_stream().cuda_stream
        rays = torch.zeros((W * H, 8), dtype=torch.float32, device="cuda")
        generate_rays(V.make_frame(vi, pi, W, H, output_memory=V.MEM_DEVICE, stream=st), 0, rays.data_ptr(), 0)
        scale = torch.tensor([0.25, 1.0, 4.0, 0.5], device="cuda")[torch.arange(W * H, device="cuda") % 4]
        rays[:, 4:7] *= scale[:, None]
        dh = torch.zeros((W * H, 32), dtype=torch.uint8, device="cuda")
        sc.trace_rays(rays.data_ptr(), W * H, dh.data_ptr(), st)
        torch.cuda.synchronize()
        ho2 = orc.trace_rays(rays.cpu().numpy())
        assert np.array_equal(dh.cpu().numpy().reshape(-1), ho2.view(np.uint8).reshape(-1))


# ---------------------------------------------------------------- secondary rays: AO + any-hit
@pytest.mark.parametrize("tech", TECHS)
@pytest.mark.parametrize("spp,ao", [(1, 3), (2, 2)])
def test_ambient_occlusion_parity(V, O, small_groom, tech, spp, ao):
    """ao_samples > 0: AO rays are spawned from the primary hit records inside the traversal kernel's refill step and traced
    terminate-on-first-hit.  Image, hit records and all traversal counters equal the oracle's."""
    pos, idx = small_groom
    W, H = 192, 120
    vi, pi = default_camera(V, W, H)
    kw = dict(spp=spp, ao_samples=ao, ao_distance=1.5, miss_rgb=(0.1, 0.2, 0.3))
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=tech)
        hg, ig, sg = sc.render(V.make_frame(vi, pi, W, H, **kw), stats=True)
        ho, io, so = orc.render(O.make_frame(vi, pi, W, H, **kw), stats=True)
        assert_bit_identical(hg, ho)
        assert np.array_equal(ig, io)
        for k in ("rays", "hits", "nodes_visited", "prims_tested"):
            assert sg[k] == so[k], k
        _, i0, _ = sc.render(V.make_frame(vi, pi, W, H, spp=spp, miss_rgb=(0.1, 0.2, 0.3)))
        hit = (hg["flags"] & 1).astype(bool)
        assert (ig[hit, :3] <= i0[hit, :3]).all() and (ig[hit, :3] < i0[hit, :3]).mean() > 0.2    # AO darkens, and visibly so
        h2, i2, _ = sc.render(V.make_frame(vi, pi, W, H, **kw))                                    # non-stats kernels
        assert h2.tobytes() == hg.tobytes() and np.array_equal(i2, ig)


def test_ambient_occlusion_device_outputs_and_shards(V, O, small_groom):
    """Device output memory + AO: hit records are kept in local HBM for the AO passes and mirrored to the caller's buffer
    (which is rank 0's frame buffer in the multi-GPU peer mode); sharded row-major frames assemble to the full frame."""
    import torch
    pos, idx = small_groom
    W, H, T, world = 200, 120, 32, 2
    vi, pi = default_camera(V, W, H)
    with V.Scene(pos, idx, technique=V.PHANTOM) as sc:
        sc.build()
        full_h, full_i, _ = sc.render(V.make_frame(vi, pi, W, H, tile_size=T, ao_samples=2))
        st = torch.cuda.current_stream().cuda_stream
        dh = torch.zeros((W * H, 32), dtype=torch.uint8, device="cuda")
        di = torch.zeros((W * H, 4), dtype=torch.uint8, device="cuda")
        for r in range(world):
            f = V.make_frame(vi, pi, W, H, tile_size=T, tile_first=r, tile_stride=world, row_major_output=1, ao_samples=2,
                             output_memory=V.MEM_DEVICE, stream=st)
            sc.render_into(f, dh.data_ptr(), di.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(dh.cpu().numpy().reshape(-1), full_h.view(np.uint8))
        assert np.array_equal(di.cpu().numpy(), full_i)
        orc = O.OracleScene(pos, idx, technique=0)
        ho, io, _ = orc.render(O.make_frame(vi, pi, W, H, ao_samples=2))
        assert full_h.tobytes() == ho.tobytes() and np.array_equal(full_i, io)


@pytest.mark.parametrize("tech", TECHS)
def test_any_hit_wavefront_rays(V, O, small_groom, tech):
    """vkhrt_trace_rays_any_hit (terminate on first hit): the accepted hit is the first one in traversal order, which is the
    oracle's order, so even these order-dependent records are bit-identical."""
    import torch
    pos, idx = small_groom
    rng = np.random.default_rng(11)
    n = 20000
    o = (rng.normal(0, 1, (n, 3)) * 6 + (0, 152, 0)).astype(np.float32)
    tgt = pos[rng.integers(0, pos.shape[0], n)] + rng.normal(0, 0.03, (n, 3))
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.concatenate([o, np.full((n, 1), 1e-3, np.float32), d, rng.uniform(2, 30, (n, 1)).astype(np.float32)], axis=1)
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        orc = O.OracleScene(pos, idx, technique=tech)
        dr = torch.from_numpy(rays).cuda()
        dh = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for any_hit in (False, True):
            sc.trace_rays(dr.data_ptr(), n, dh.data_ptr(), st, any_hit=any_hit)
            torch.cuda.synchronize()
            ho = orc.trace_rays(rays, any_hit=any_hit)
            assert (ho["flags"] & 1).sum() > 2000
            assert np.array_equal(dh.cpu().numpy().reshape(-1), ho.view(np.uint8).reshape(-1)), f"any_hit={any_hit}"


@pytest.mark.gpu
@pytest.mark.parametrize("tech", TECHS)
def test_render_multi_from_one_process(V, small_groom, tech):
    """vkhrt_render_multi: several scene handles (one per device; here also several on one device), tiles dealt round-robin,
    one host thread per handle, host re-ordering: the frame must equal vkhrt_render's, bit for bit"""
    pos, idx = small_groom
    W, H = 333, 201                                    # partial tiles on both edges
    vi, pi = default_camera(V, W, H)
    ndev = V.device_count()
    with V.Scene(pos, idx, technique=tech) as ref:
        ref.build()
        h0, i0, _ = ref.render(V.make_frame(vi, pi, W, H, spp=2, miss_rgb=(0.1, 0.2, 0.3)))
    for n in (1, 2, 3):
        scenes = [V.Scene(pos, idx, technique=tech, device=r % ndev).build() for r in range(n)]
        try:
            for tile in (0, 32):
                h, img = V.render_multi(scenes, V.make_frame(vi, pi, W, H, spp=2, miss_rgb=(0.1, 0.2, 0.3), tile_size=tile))
                assert h.tobytes() == h0.tobytes() and np.array_equal(img, i0), (n, tile)
            h, img = V.render_multi(scenes, V.make_frame(vi, pi, W, H), rgba=False)
            assert img is None and h.tobytes() == h0.tobytes()
        finally:
            for sc in scenes:
                sc.close()
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        with pytest.raises(V.VkhrtError):
            V.render_multi([sc, sc], V.make_frame(vi, pi, W, H))                      # one handle twice
        with pytest.raises(V.VkhrtError):
            V.render_multi([sc], V.make_frame(vi, pi, W, H, tile_first=1, tile_stride=2))   # the frame must be whole


@pytest.mark.parametrize("env", [{"VKHRT_POOL_MIN_RATIO": "0"}, {"VKHRT_POOL_MIN_RATIO": "0", "VKHRT_POOL_CFG": "1"},
                                 {"VKHRT_POOL_MIN_RATIO": "0", "VKHRT_POOL_CFG": "2"}, {"VKHRT_POOL": "0"},
                                 {"VKHRT_POOL_MIN_RATIO": "0", "VKHRT_POOL_LSS": "1", "VKHRT_POOL_DOTS": "1"}, {"VKHRT_REFIT_EXIT_CAP": "16"}],
                         ids=["pool-on-every-frame", "pool-56x8", "pool-64x6", "lane-bound-only", "pool-for-lss-and-dots", "refit-exit-list-overflows"])
def test_both_traversal_kernels_pass_the_whole_suite(env):
    """Phantom primary rays have two traversal kernels: the per-warp ray pool (frames >= 3x its resident capacity) and the lane-bound
    kernel (everything else).  The library reads its switches once per process, so the whole parity file is re-run in a subprocess
    with the pool forced onto every frame size (tiny, ragged, sharded, empty ...) and with the pool switched off (full-size C2 on
    the lane-bound kernel).  One variant sends LSS and DOTS primary rays through the pool kernel too (an opt-in switch: measured no faster); the last one shrinks the
    refit's exit list to 16 entries, so that nearly every walker that leaves its tile finishes inside materialise_refit_kernel (the overflow path)."""
    import os
    import subprocess
    import sys
    if os.environ.get("VKHRT_NESTED"):
        pytest.skip("nested run")
    e = dict(os.environ, VKHRT_NESTED="1", **env)
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), os.path.join(here, "test_random_scenes.py"), "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"],
                       env=e, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
