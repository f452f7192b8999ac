"""Every BASELINE.json config AT ITS OWN SIZE against the CPU oracle (VERDICT r1 "what's weak" #2): the full frame is rendered
through the C ABI, and a seeded 65 536-pixel stratified subset of it is compared with the oracle's own build + traversal of the
same groom — hit records bit-identical, and for the multi-sample configs the RGBA of those pixels byte-identical (PSNR >= 45 dB is
the north-star bar).  32-bit index limits, stack depth, the spill area and the sample accumulation only show at these sizes.
The oracle builds its own LBVH (OpenMP): C4 = 64 M leaves, C5 = 128 M leaves, ~20 GB of host memory and about a minute."""
import os
import numpy as np
import pytest
from conftest import default_camera
from oracle.parity import parity_metrics, stratified_pixels

pytestmark = pytest.mark.gpu

N_CHECK = 65536


def host_gb():
    try:
        return os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
    except (ValueError, OSError):
        return 0.0


def check(V, O, pos, idx, tech, W, H, spp, rgba, shade_mode=0, radius_per_vertex=None):
    vi, pi = default_camera(V, W, H)
    sub = stratified_pixels(W, H, N_CHECK)
    with V.Scene(pos, idx, technique=tech, radius_per_vertex=radius_per_vertex) as sc:
        sc.build()
        hg, ig, _ = sc.render(V.make_frame(vi, pi, W, H, spp=spp, shade_mode=shade_mode), rgba=rgba)
    orc = O.OracleScene(pos, idx, technique=tech, radius_per_vertex=radius_per_vertex)
    ho, io, _ = orc.render(O.make_frame(vi, pi, W, H, spp=spp, shade_mode=shade_mode), rgba=rgba, pixel_subset=sub)
    orc.close()
    k = sub.astype(np.int64)
    m = parity_metrics(hg[k], ho, ig[k] if rgba else None, io if rgba else None)
    print(f"parity at size: tech {tech} {W}x{H}x{spp}: {m}")
    assert m["rays_checked"] == N_CHECK and m["hit_fraction"] > 0.3
    assert m["within_tolerance"], m
    assert m["bit_identical"], m
    if rgba:
        assert m["psnr"] >= 45.0 and m["rgba_identical"], m
    return m


@pytest.fixture(scope="module")
def groom_c2(V):
    return V.generate_groom(100000, 32, V.GROOM_CURLY)


def test_config2_phantom_at_size(V, O, groom_c2):
    """BASELINE configs[1]: curly 100k x 32 (3.2 M curves), 1920x1080, Phantom, hit buffer."""
    check(V, O, *groom_c2, V.PHANTOM, 1920, 1080, 1, False)


def test_config2_groom_with_per_vertex_radii_at_size(V, O, groom_c2):
    """The C2 groom with tapered strands (radius 0.02 at the root -> 0.005 at the tip, north_star's "per-vertex radius in"), 1920x1080, Phantom:
    a frame of this size goes through the ray-pool kernel with the taper terms (radius(t), cone slant) in its set-up and march stages."""
    rad = np.tile(np.linspace(0.02, 0.005, 33, dtype=np.float32), 100000)
    check(V, O, *groom_c2, V.PHANTOM, 1920, 1080, 1, False, radius_per_vertex=rad)


def test_config3_lss_at_size(V, O, groom_c2):
    """BASELINE configs[2]: the same groom as linear swept spheres with shading.glsl, 8 spp, RGBA."""
    check(V, O, *groom_c2, V.LSS, 1920, 1080, 8, True)


def test_config4_dots_at_size(V, O):
    """BASELINE configs[3]: DOTS tessellation of 1 M strands x 16 segments (64 M triangles), 1920x1080."""
    if host_gb() < 24:
        pytest.skip("the oracle's 64 M-leaf LBVH needs ~12 GB of host memory")
    pos, idx = V.generate_groom(1000000, 16, V.GROOM_CURLY)
    check(V, O, pos, idx, V.DOTS, 1920, 1080, 1, False)


def test_config5_phantom_64M(V, O):
    """BASELINE configs[4]: 1 M strands x 64 segments (64 M curves), 3840x2160 x 64 spp, Phantom, RGBA."""
    if host_gb() < 40 or os.environ.get("VKHRT_SKIP_C5"):
        pytest.skip("the oracle's 128 M-leaf LBVH needs ~20 GB of host memory")
    pos, idx = V.generate_groom(1000000, 64, V.GROOM_CURLY)
    check(V, O, pos, idx, V.PHANTOM, 3840, 2160, 64, True)
