"""CPU tests that PIN the oracle (oracle/vkhrt_oracle.cpp).  The reference has no tests, golden vectors or
CPU path ("parity unpinned"), so the pins are:
  1. the known-answer values derived independently during the survey;
  2. a second, literal numpy transcription of the GLSL (tests/glsl_transcription.py, generic inverse(mat4));
  3. analytic answers (fp64 cylinder, capsule, triangle);
  4. structural invariants of the LBVH and BVH-vs-brute-force equality of the closest-hit search.
No GPU is needed."""
import numpy as np
import pytest

import glsl_transcription as G

STRAIGHT = np.array([0, 0, 0, 1 / 3, 0, 0, 2 / 3, 0, 0, 1, 0, 0], np.float32)


# ---------------------------------------------------------------- survey KATs (A5)
@pytest.mark.parametrize("origin,t,u,normal,iters", [
    ((0.5, 0.0, 20.0), 19.98, 0.5, (0.0, 0.0, 1.0), 2),
    ((0.5, 0.01, 20.0), 19.982679, 0.5, (0.0, 0.5, 0.866025), 2),
    ((0.25, 0.019, 5.0), 4.9937549, 0.25, (0.0, 0.95, 0.31225), 3),
])
def test_prhi_survey_kats(O, origin, t, u, normal, iters):
    to, uo, no, it = O.prhi(origin, (0, 0, -1), STRAIGHT)
    assert abs(to - t) <= 2e-6 * t
    assert abs(uo - u) < 1e-6
    assert np.allclose(no, normal, atol=2e-5)
    assert it == iters


def test_prhi_rejections(O):
    # outside the bounding cylinder: rejected before any cone iteration (hair_intersection.rint:30)
    assert O.prhi((0.5, 0.0205, 20.0), (0, 0, -1), STRAIGHT) [0] == 0.0
    assert O.prhi((0.5, 0.0205, 20.0), (0, 0, -1), STRAIGHT)[3] == 0
    # ray along the axis: NaN path, nothing reported (Phantom tubes have no end caps)
    assert O.prhi((-5, 0, 0), (1, 0, 0), STRAIGHT)[0] == 0.0
    # a tube behind the origin is never reported (tHit > 0, hair_intersection.rint:146)
    assert O.prhi((0.5, 0, -20.0), (0, 0, -1), STRAIGHT)[0] <= 0.0


def test_ray_cylinder_has_no_t_positive_test(O):
    # cylinder.glsl:8-46 is a boolean over the whole LINE: a cylinder behind the origin passes
    assert O.ray_cylinder((0.5, 0, -20), (0, 0, -1), (0, 0, 0), (1, 0, 0), 0.02)
    assert O.ray_cylinder((0.5, 0, 20), (0, 0, -1), (0, 0, 0), (1, 0, 0), 0.02)
    assert not O.ray_cylinder((0.5, 0.03, 20), (0, 0, -1), (0, 0, 0), (1, 0, 0), 0.02)
    # cap region: just beyond the end, inside the radius
    d = np.array([-1, 0, 0.02]) / np.sqrt(1.0004)
    assert O.ray_cylinder((1.5, 0.0, 0.0), d, (0, 0, 0), (1, 0, 0), 0.05)          # enters through the end cap
    assert not O.ray_cylinder((1.5, 0.0, 0.06), d, (0, 0, 0), (1, 0, 0), 0.05)
    rng = np.random.default_rng(2)
    differ = 0
    for _ in range(300):                                                              # vs the literal transcription
        ro = rng.normal(0, 2, 3).astype(np.float32)
        rd = rng.normal(0, 1, 3)
        rd = (rd / np.linalg.norm(rd)).astype(np.float32)
        p0, p1 = rng.normal(0, 1, 3).astype(np.float32), rng.normal(0, 1, 3).astype(np.float32)
        differ += O.ray_cylinder(ro, rd, p0, p1, 0.3) != G.ray_cylinder_intersect(ro, rd, p0, p1, np.float32(0.3))
    assert differ <= 1                                                                # fused vs unfused: grazing only


# ---------------------------------------------------------------- curve.glsl
def test_curve_point_axis_against_fp64(O):
    rng = np.random.default_rng(3)
    for _ in range(50):
        c = rng.normal(0, 1, 12).astype(np.float32)
        t = float(np.float32(rng.random()))
        p = c.reshape(4, 3).astype(np.float64)
        u = 1 - t
        ref = u ** 3 * p[0] + 3 * u * u * t * p[1] + 3 * u * t * t * p[2] + t ** 3 * p[3]
        dref = -3 * u * u * p[0] + 3 * (3 * t * t - 4 * t + 1) * p[1] + 3 * (2 - 3 * t) * t * p[2] + 3 * t * t * p[3]
        assert np.allclose(O.curve_point(c, t), ref, atol=2e-6)
        assert np.allclose(O.curve_axis(c, t), dref, atol=1e-5)
        # and within rounding of the literal (unfused) numpy float32 transcription
        assert np.allclose(O.curve_point(c, t), G.sample_curve_point(list(c.reshape(4, 3)), t), atol=1e-6)
        assert np.allclose(O.curve_axis(c, t), G.sample_curve_axis(list(c.reshape(4, 3)), t), atol=4e-6)
    assert np.array_equal(O.curve_point(STRAIGHT, 0.0), [0, 0, 0]) and np.array_equal(O.curve_point(STRAIGHT, 1.0), [1, 0, 0])


# ---------------------------------------------------------------- second transcription (generic inverse(mat4))
def test_prhi_against_literal_glsl_transcription(O):
    rng = np.random.default_rng(1)
    n_hit = disagree = 0
    worst = 0.0
    for _ in range(400):
        p0 = rng.normal(0, 1, 3)
        d = rng.normal(0, 1, 3)
        d /= np.linalg.norm(d)
        cv = np.array([p0, p0 + d * 0.33 + rng.normal(0, 0.03, 3), p0 + d * 0.66 + rng.normal(0, 0.03, 3), p0 + d], np.float32)
        tgt = 0.5 * (cv[0] + cv[3]) + rng.normal(0, 0.02, 3)
        ro = (tgt + rng.normal(0, 1, 3) * 10).astype(np.float32)
        rd = tgt - ro
        rd = (rd / np.linalg.norm(rd)).astype(np.float32)
        a = G.prhi(ro, rd, cv)
        b = O.prhi(ro, rd, cv)
        if (a[0] > 0) != (b[0] > 0):
            disagree += 1
        elif a[0] > 0:
            n_hit += 1
            worst = max(worst, abs(a[0] - b[0]) / b[0])
            assert abs(a[1] - b[1]) < 1e-3 and np.allclose(a[2], b[2], atol=2e-3)
    assert n_hit > 100
    assert disagree <= 2            # grazing rays may flip under the generic-vs-rigid inverse rounding
    assert worst <= 1e-4            # north_star tolerance on relative t


# ---------------------------------------------------------------- analytic: straight Bezier == cylinder
def test_phantom_vs_analytic_cylinder(O):
    """Phantom is NOT an analytic tube (median 3e-5, max ~3e-4 relative t, a few %
    grazing misses).  Pin those statistics so a transcription error (which would blow them up) is caught."""
    rng = np.random.default_rng(11)
    rel, missed, false_pos, n_an = [], 0, 0, 0
    for _ in range(1500):
        x = rng.uniform(0.05, 0.95)
        off = rng.uniform(-0.03, 0.03)
        ro = np.array([x + rng.uniform(-2, 2), off * 3 + rng.uniform(-2, 2), 20.0], np.float32)
        tgt = np.array([x, off, 0.0])
        rd = tgt - ro
        rd = (rd / np.linalg.norm(rd)).astype(np.float32)
        an = G.analytic_cylinder(ro, rd, (0, 0, 0), (1, 0, 0), 0.02)
        t, u, n, _ = O.prhi(ro, rd, STRAIGHT)
        if an is not None:
            n_an += 1
            if t > 0:
                rel.append(abs(t - an[0]) / an[0])
                assert abs(u - an[1]) < 2e-2
            else:
                missed += 1
        elif t > 0:
            false_pos += 1
    rel = np.array(rel)
    assert n_an > 400
    assert np.median(rel) < 1e-4 and rel.max() < 1e-3
    assert missed <= 0.06 * n_an and false_pos <= 0.03 * n_an


def test_lss_capsule_analytic(O):
    # constant radius: capsule.  Ray down -z onto the body, the end spheres, and a miss.
    lss = [0, 0, 0, 0.05, 1, 0, 0, 0.05]
    hit, t, u, n = O.lss((0.5, 0.0, 10.0), (0, 0, -1), lss)
    assert hit and abs(t - 9.95) < 1e-5 and abs(u - 0.5) < 1e-6 and np.allclose(n, (0, 0, 1), atol=1e-6)
    hit, t, u, n = O.lss((0.25, 0.03, 10.0), (0, 0, -1), lss)
    assert hit and abs(t - (10 - 0.04)) < 1e-5 and abs(u - 0.25) < 1e-6 and np.allclose(n, (0, 0.6, 0.8), atol=1e-5)
    hit, t, u, n = O.lss((-0.03, 0.0, 10.0), (0, 0, -1), lss)           # start cap
    assert hit and u == 0.0 and abs(t - (10 - 0.04)) < 1e-5 and np.allclose(n, (-0.6, 0, 0.8), atol=1e-5)
    hit, t, u, n = O.lss((1.03, 0.0, 10.0), (0, 0, -1), lss)            # end cap
    assert hit and u == 1.0 and abs(t - (10 - 0.04)) < 1e-5 and np.allclose(n, (0.6, 0, 0.8), atol=1e-5)
    assert not O.lss((0.5, 0.051, 10.0), (0, 0, -1), lss)[0]
    assert not O.lss((1.06, 0.0, 10.0), (0, 0, -1), lss)[0]
    # tapered: envelope of spheres. On the body the hit point is at distance r(u) from the axis point it reports.
    tap = [0, 0, 0, 0.1, 2, 0, 0, 0.02]
    rng = np.random.default_rng(5)
    n_hit = 0
    for _ in range(300):
        ro = np.array([rng.uniform(-0.2, 2.2), rng.uniform(-0.15, 0.15), 6.0], np.float32)
        rd = np.array([rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02), -1.0])
        rd = (rd / np.linalg.norm(rd)).astype(np.float32)
        hit, t, u, n = O.lss(ro, rd, tap)
        if not hit:
            continue
        n_hit += 1
        p = ro.astype(np.float64) + t * rd.astype(np.float64)
        c = np.array([2.0 * u, 0, 0])
        r = 0.1 + u * (0.02 - 0.1)
        assert abs(np.linalg.norm(p - c) - r) < 2e-5          # on the sphere of parameter u
        assert np.allclose(n, (p - c) / np.linalg.norm(p - c), atol=2e-4)
        # no sphere of the family contains the hit point (it is on the envelope, first hit)
        us = np.linspace(0, 1, 201)
        d = np.linalg.norm(p[None, :] - np.stack([2 * us, 0 * us, 0 * us], 1), axis=1) - (0.1 + us * (0.02 - 0.1))
        assert d.min() > -3e-5
    assert n_hit > 60


def test_triangle_analytic(O):
    tri = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    hit, t, u, n = O.tri((0.25, 0.25, 5.0), (0, 0, -1), tri, 0)
    assert hit and t == 5.0 and abs(u - 0.5) < 1e-7 and np.array_equal(n, (0, 0, 1))
    hit, t, u, n = O.tri((0.25, 0.25, -5.0), (0, 0, 1), tri, 1)      # back face: not culled, normal faces the ray
    assert hit and t == 5.0 and abs(u - 0.25) < 1e-7 and np.array_equal(n, (0, 0, -1))
    assert not O.tri((0.75, 0.75, 5.0), (0, 0, -1), tri, 0)[0]
    assert not O.tri((0.25, 0.25, 5.0), (1, 0, 0), tri, 0)[0]         # parallel


# ---------------------------------------------------------------- geometry_processor.cpp
def test_generate_curves_kat(O):
    # polyline (0,0,0)(1,0,0)(2,1,0)
    pos = np.array([[0, 0, 0], [1, 0, 0], [2, 1, 0]], np.float32)
    idx = np.array([[0, 1], [1, 2]], np.uint32)
    c = O.OracleScene(pos, idx).primitives().reshape(2, 4, 3)
    assert np.allclose(c[0], [[0, 0, 0], [1 / 6, 0, 0], [2 / 3, -1 / 6, 0], [1, 0, 0]], atol=1e-7)
    assert np.allclose(c[1], [[1, 0, 0], [4 / 3, 1 / 6, 0], [11 / 6, 5 / 6, 0], [2, 1, 0]], atol=1e-7)
    # two strands that do NOT share a vertex position: no tangent leaks across the boundary
    pos2 = np.array([[0, 0, 0], [1, 0, 0], [5, 5, 5], [6, 5, 5]], np.float32)
    idx2 = np.array([[0, 1], [2, 3]], np.uint32)
    c2 = O.OracleScene(pos2, idx2).primitives().reshape(2, 4, 3)
    assert np.allclose(c2[0], [[0, 0, 0], [1 / 6, 0, 0], [5 / 6, 0, 0], [1, 0, 0]], atol=1e-7)
    assert np.allclose(c2[1], [[5, 5, 5], [5 + 1 / 6, 5, 5], [5 + 5 / 6, 5, 5], [6, 5, 5]], atol=1e-6)


def test_generate_aabbs(O):
    pos = np.array([[0, 0, 0], [1, 0, 0], [2, 1, 0]], np.float32)
    idx = np.array([[0, 1], [1, 2]], np.uint32)
    sc = O.OracleScene(pos, idx)
    c = sc.primitives().reshape(2, 4, 3)
    r = np.float32(0.02)
    b = np.array([O.curve_aabb(c[i], r) for i in range(2)])                      # GenerateAABBs, geometry_processor.cpp:422-436
    assert np.array_equal(b[:, :3], c.min(axis=1) - r) and np.array_equal(b[:, 3:], c.max(axis=1) + r)
    # the BVH leaves are K piece boxes per curve (VKHRT_LEAF_SPLIT_PHANTOM): each inside the reference box (plus a few ulps),
    # and together they hold every point within r of the curve
    K = sc.leaf_split
    lb = sc.aabbs().reshape(2, K, 6)
    assert K in (2, 4) and (lb[:, :, :3] >= b[:, None, :3] - 1e-5).all() and (lb[:, :, 3:] <= b[:, None, 3:] + 1e-5).all()
    for i in range(2):
        for t in np.linspace(0, 1, 101):
            pt = O.curve_point(c[i], np.float32(t))
            k = min(K - 1, int(t * K))
            assert (pt - r >= lb[i, k, :3] - 1e-6).all() and (pt + r <= lb[i, k, 3:] + 1e-6).all()


def test_dots_geometry(O, V):
    pos, idx = V.generate_groom(50, 8, V.GROOM_CURLY)
    tris = O.OracleScene(pos, idx, technique=2).primitives().reshape(-1, 4, 3, 3).astype(np.float64)
    assert tris.shape[0] == idx.shape[0]
    s, e = pos[idx[:, 0]].astype(np.float64), pos[idx[:, 1]].astype(np.float64)
    fwd = (e - s) / np.linalg.norm(e - s, axis=1, keepdims=True)
    for face in range(2):
        a, b = tris[:, 2 * face], tris[:, 2 * face + 1]
        off = a[:, 0] - s                                    # start + v*r
        assert np.allclose(np.linalg.norm(off, axis=1), 0.02, atol=1e-5)
        assert np.abs(np.sum(off * fwd, axis=1)).max() < 1e-5   # orthogonal to the segment
        assert np.allclose(a[:, 1], e - off, atol=1e-5) and np.allclose(a[:, 2], e + off, atol=1e-5)
        assert np.allclose(b[:, 0], s + off, atol=1e-5) and np.allclose(b[:, 1], s - off, atol=1e-5) and np.allclose(b[:, 2], e - off, atol=1e-5)
    o0, o1 = tris[:, 0, 0] - s, tris[:, 2, 0] - s
    assert np.abs(np.sum(o0 * o1, axis=1)).max() < 1e-6      # the two strips are orthogonal


def test_perp_stark_priority(O):
    # geometry_processor.cpp:201-212: axis of the smallest |component|, strict <, priority x then y then z
    for fwd, axis in (((0, 1, 0), (0, 0, 1)), ((1, 0, 0), (0, 0, 1)), ((0, 0, 1), (0, 1, 0)), ((1, 1, 1), (0, 0, 1)), ((0.1, 0.5, 0.8), (1, 0, 0))):
        f = np.array(fwd, np.float64) / np.linalg.norm(fwd)
        pos = np.array([[0, 0, 0], f], np.float32)
        tri = O.OracleScene(pos, np.array([[0, 1]], np.uint32), technique=2, radius=1.0).primitives().reshape(4, 3, 3)
        s_vec = tri[0, 0] - pos[0]
        want = np.cross(f, axis)
        want /= np.linalg.norm(want)
        assert np.allclose(s_vec, want, atol=1e-6), (fwd, s_vec, want)


def test_lss_geometry_and_radius_floor(O):
    pos = np.array([[0, 0, 0], [1, 0, 0], [2, 1, 0]], np.float32)
    idx = np.array([[0, 1], [1, 2]], np.uint32)
    p = O.OracleScene(pos, idx, technique=1).primitives()
    assert np.array_equal(p[0], np.float32([0, 0, 0, 0.02, 1, 0, 0, 0.02])) and np.array_equal(p[1], np.float32([1, 0, 0, 0.02, 2, 1, 0, 0.02]))
    p = O.OracleScene(pos, idx, technique=1, radius_per_vertex=np.array([0.0001, 0.01, 0.03], np.float32)).primitives()
    assert np.array_equal(p[:, 3], np.float32([0.001, 0.01])) and np.array_equal(p[:, 7], np.float32([0.01, 0.03]))   # max(r, 0.001)


# ---------------------------------------------------------------- ray_gen.rgen + fly_camera.cpp
def test_raygen_survey_values(O, V):
    vi, pi = V.camera_matrices(aspect=float(np.float32(1920) / np.float32(1080)))
    for (px, py), want in (((0, 0), (-0.6642564, 0.3734928, -0.6475080)), ((1919, 0), (0.6642564, 0.3734928, -0.6475080)),
                           ((960, 540), (0.0005346, -0.0005346, -0.9999997))):
        o, d = O.raygen(vi, pi, 1920, 1080, px, py)
        assert np.array_equal(o, (0, 150, 20))
        assert np.allclose(d, want, atol=2e-6)
        assert abs(np.linalg.norm(d.astype(np.float64)) - 1) < 1e-6


def test_raygen_closed_form(O, V):
    W, H, fov = 640, 360, 60.0
    vi, pi = V.camera_matrices(aspect=W / H, fov=fov)
    th = np.tan(np.radians(fov) / 2)
    for px, py in ((0, 0), (100, 200), (639, 359), (320, 180)):
        dx, dy = 2 * (px + 0.5) / W - 1, 2 * (py + 0.5) / H - 1
        want = np.array([dx * (W / H) * th, -dy * th, -1.0])
        want /= np.linalg.norm(want)
        assert np.allclose(O.raygen(vi, pi, W, H, px, py)[1], want, atol=3e-6)   # row 0 looks UP
    # sample 0 is the pixel centre; other samples stay inside the pixel
    o0, d0 = O.raygen(vi, pi, W, H, 10, 10, 0)
    for s in range(1, 8):
        _, ds = O.raygen(vi, pi, W, H, 10, 10, s)
        assert not np.array_equal(ds, d0)
        _, dl = O.raygen(vi, pi, W, H, 9, 9, 0)
        _, dr = O.raygen(vi, pi, W, H, 11, 11, 0)
        assert dl[0] / -dl[2] < ds[0] / -ds[2] < dr[0] / -dr[2]      # x slope is linear in the pixel x


# ---------------------------------------------------------------- shading.glsl / debug.glsl
def test_shade_and_debug_palette(O):
    assert np.allclose(O.shade((0, -1, 0)), (0.7, 0.5, 0.4), atol=1e-7)
    assert np.allclose(O.shade((0, 1, 0)), (0.7, 0.5, 0.4), atol=1e-7)          # abs()
    assert np.allclose(O.shade((1, 0, 0)), (0.3, 0.3, 0.3), atol=1e-7)
    pal = [(1, 0, .3), (.8, .2, .3), (.6, .4, .3), (.4, .6, .3), (.2, .8, .3), (0, 1, .3)]
    for i in range(13):
        assert np.allclose(O.shade((0, 0, 0), prim=i, mode=1), pal[i % 6], atol=1e-7)


# ---------------------------------------------------------------- LBVH invariants + closest-hit search
@pytest.mark.parametrize("tech", [0, 1, 2])
def test_lbvh_invariants(O, V, tech):
    pos, idx = V.generate_groom(300, 12, V.GROOM_CURLY)
    sc = O.OracleScene(pos, idx, technique=tech)
    nodes, ids, morton, lohi = sc.bvh()
    n = sc.n_leaves
    K = sc.leaf_split                                                             # leaf pieces per group: 2 or 4 / 2 / 4
    assert K in ((2, 4), (2,), (4,))[tech] and n == idx.shape[0] * K and sc.n_primitives == idx.shape[0] * (4 if tech == 2 else 1)
    boxes = sc.aabbs()
    if tech == 2:   # the K piece boxes of a strip: together they hold its 12 vertices, and none sticks out of the strip's own box
        tr = sc.primitives().reshape(idx.shape[0], 12, 3)
        pb = boxes.reshape(idx.shape[0], K, 6)
        assert (pb[:, :, :3].min(axis=1) <= tr.min(axis=1)).all() and (pb[:, :, 3:].max(axis=1) >= tr.max(axis=1)).all()
        assert (pb[:, :, :3] >= tr.min(axis=1)[:, None] - 2e-4).all() and (pb[:, :, 3:] <= tr.max(axis=1)[:, None] + 2e-4).all()
        vol = np.prod(pb[:, :, 3:] - pb[:, :, :3], axis=2).sum(axis=1) / np.prod(tr.max(axis=1) - tr.min(axis=1), axis=1)
        assert np.median(vol) < 0.5                                               # the point of the split: much less empty space
    assert nodes.shape[0] == n - 1
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))
    assert (np.diff(morton.astype(np.int64)) >= 0).all()
    # stable: equal keys keep ascending primitive ids
    eq = np.nonzero(np.diff(morton.astype(np.int64)) == 0)[0]
    assert (ids[eq] < ids[eq + 1]).all()
    cen = 0.5 * (boxes[:, :3] + boxes[:, 3:])
    assert np.array_equal(lohi[:3], cen.min(axis=0)) and np.array_equal(lohi[3:], cen.max(axis=0))
    # every leaf and every internal node (except the root) is referenced exactly once; boxes are exact unions
    seen_leaf, seen_int = np.zeros(n, int), np.zeros(n - 1, int)

    def box_of(ref):
        if ref >> 31:
            p = ref & 0x7FFFFFFF
            seen_leaf[p] += 1
            return boxes[ids[p], :3], boxes[ids[p], 3:]
        seen_int[ref] += 1
        nd = nodes[ref]
        return np.minimum(nd["lo0"], nd["lo1"]), np.maximum(nd["hi0"], nd["hi1"])
    for nd in nodes:
        for k in (0, 1):
            lo, hi = box_of(int(nd[f"child{k}"]))
            assert np.array_equal(nd[f"lo{k}"], lo) and np.array_equal(nd[f"hi{k}"], hi)
            if int(nd[f"child{k}"]) >> 31:
                assert nd[f"prim{k}"] == ids[int(nd[f"child{k}"]) & 0x7FFFFFFF]
    assert (seen_leaf == 1).all() and seen_int[0] == 0 and (seen_int[1:] == 1).all()


@pytest.mark.parametrize("tech", [0, 1, 2])
def test_bvh_search_equals_brute_force(O, V, tech):
    pos, idx = V.generate_groom(150, 8, V.GROOM_CURLY)
    sc = O.OracleScene(pos, idx, technique=tech)
    W, H = 96, 64
    vi, pi = V.camera_matrices(aspect=W / H)
    f = O.make_frame(vi, pi, W, H)
    h1, i1, _ = sc.render(f)
    h2, i2, _ = sc.render(f, brute=True)
    assert (h1["flags"] & 1).sum() > 50
    assert h1.tobytes() == h2.tobytes() and np.array_equal(i1, i2)


def test_mailbox_is_result_neutral(O, V, monkeypatch):
    """VKHRT_MAILBOX_PHANTOM (include/vkhrt_b200.h): skipping the curve a ray tested last changes no hit record and no node
    visit, only the number of curve tests and cone iterations.  ORC_MAILBOX / ORC_LEAF_SPLIT are the oracle's study overrides."""
    pos, idx = V.generate_groom(400, 12, V.GROOM_CURLY)
    W, H = 128, 96
    vi, pi = V.camera_matrices(aspect=W / H)
    f = O.make_frame(vi, pi, W, H)
    res = {}
    for K in (1, 2, 4):
        for mb in (0, 1):
            monkeypatch.setenv("ORC_LEAF_SPLIT", str(K))
            monkeypatch.setenv("ORC_MAILBOX", str(mb))
            sc = O.OracleScene(pos, idx, technique=0, radius=0.05)
            assert sc.leaf_split == K
            h, img, st = sc.render(f, stats=True)
            res[K, mb] = (h.tobytes(), img.tobytes(), st)
    monkeypatch.delenv("ORC_LEAF_SPLIT"); monkeypatch.delenv("ORC_MAILBOX")
    ref = res[1, 0]
    assert res[1, 1][2]["prims_tested"] == ref[2]["prims_tested"]            # one leaf per curve: a ray never meets a curve twice
    for (K, mb), (hb, ib, st) in res.items():
        assert hb == ref[0] and ib == ref[1], (K, mb)
        assert st["nodes_visited"] == res[K, 0][2]["nodes_visited"]
        assert st["prims_tested"] <= res[K, 0][2]["prims_tested"] and st["phantom_iterations"] <= res[K, 0][2]["phantom_iterations"]
    assert res[4, 1][2]["prims_tested"] < res[4, 0][2]["prims_tested"]        # ... and with pieces it does skip something


@pytest.mark.parametrize("tech", [0, 1, 2])
def test_wide_traversal_study_is_result_neutral(O, V, tech, monkeypatch):
    """ORC_WIDE=1 (oracle study of a 4-wide walk over the same tree, the next step named in DESIGN.md §8): identical hit records
    and images, about half the dependent node fetches for about the same number of slab tests."""
    pos, idx = V.generate_groom(400, 12, V.GROOM_CURLY)
    W, H = 128, 96
    vi, pi = V.camera_matrices(aspect=W / H)
    f = O.make_frame(vi, pi, W, H)
    h2, i2, s2 = O.OracleScene(pos, idx, technique=tech, radius=0.05).render(f, stats=True)
    monkeypatch.setenv("ORC_WIDE", "1")
    h4, i4, s4 = O.OracleScene(pos, idx, technique=tech, radius=0.05).render(f, stats=True)
    monkeypatch.delenv("ORC_WIDE")
    assert (h2["flags"] & 1).sum() > 200
    assert h4.tobytes() == h2.tobytes() and np.array_equal(i4, i2)
    assert s4["nodes_visited"] < 0.62 * s2["nodes_visited"]
    assert s4["sched_steps"][0] <= 1.05 * 2 * s2["nodes_visited"]                   # slab tests: 2 per BVH2 visit vs <= 4 per wide visit
    assert 0 < s4["sched_steps"][1] <= 64                                         # deepest stack


def test_oracle_edge_cases(O, V):
    W, H = 32, 24
    vi, pi = V.camera_matrices(aspect=W / H)
    empty = O.OracleScene(np.zeros((0, 3), np.float32), np.zeros((0, 2), np.uint32))
    h, img, _ = empty.render(O.make_frame(vi, pi, W, H, miss_rgb=(1.0, 0.5, 0.0)))
    assert (h["flags"] == 0).all() and np.isinf(h["t"]).all() and (h["segment"] == 0xFFFFFFFF).all()
    assert (img == np.array([255, 128, 0, 255], np.uint8)).all()
    with pytest.raises(ValueError):
        O.OracleScene(np.zeros((2, 3), np.float32), np.array([[0, 5]], np.uint32))
    # a single segment; t-interval honoured
    pos = np.array([[-1, 150, 0], [1, 150, 0]], np.float32)
    one = O.OracleScene(pos, np.array([[0, 1]], np.uint32), radius=0.5)
    h, _, _ = one.render(O.make_frame(vi, pi, W, H))
    assert (h["flags"] & 1).sum() > 0 and (h["segment"][(h["flags"] & 1) == 1] == 0).all()
    h, _, _ = one.render(O.make_frame(vi, pi, W, H, t_min=0.001, t_max=5.0))      # tube is ~19.5 away
    assert (h["flags"] & 1).sum() == 0


def test_spp_mean_and_unorm8(O, V):
    pos, idx = V.generate_groom(100, 8, V.GROOM_STRAIGHT)
    sc = O.OracleScene(pos, idx, technique=1)
    W, H = 48, 32
    vi, pi = V.camera_matrices(aspect=W / H)
    h1, i1, _ = sc.render(O.make_frame(vi, pi, W, H, spp=1, miss_rgb=(0.2, 0.4, 0.6)))
    h4, i4, _ = sc.render(O.make_frame(vi, pi, W, H, spp=4, miss_rgb=(0.2, 0.4, 0.6)))
    assert h1.tobytes() == h4.tobytes()                   # hit buffer = sample 0 = pixel centre
    miss = (h1["flags"] & 1) == 0
    assert (i1[miss] == np.array([51, 102, 153, 255], np.uint8)).all()    # round(c*255)
    assert (i4[:, 3] == 255).all() and not np.array_equal(i1, i4)


# ---------------------------------------------------------------- secondary rays: AO + any-hit
def test_ao_directions_are_unit_cosine_weighted_and_deterministic(O):
    n = np.array([0.0, 0.6, 0.8], np.float32)
    d = np.array([O.ao_direction(n, pix, 0, k) for pix in range(400) for k in range(8)], np.float64)
    assert np.abs(np.linalg.norm(d, axis=1) - 1).max() < 1e-6
    c = d @ n.astype(np.float64)
    assert (c >= -1e-6).all()                       # hemisphere about n
    assert abs(c.mean() - 2.0 / 3.0) < 0.02         # cosine-weighted: E[cos] = 2/3
    t1 = np.array([1.0, 0.0, 0.0]); t2 = np.cross(n, t1)
    assert abs((d @ t1).mean()) < 0.03 and abs((d @ t2).mean()) < 0.03      # no tangential bias
    assert np.array_equal(O.ao_direction(n, 17, 3, 5), O.ao_direction(n, 17, 3, 5))
    assert not np.array_equal(O.ao_direction(n, 17, 3, 5), O.ao_direction(n, 17, 3, 6))
    assert not np.array_equal(O.ao_direction(n, 17, 3, 5), O.ao_direction(n, 18, 3, 5))


@pytest.mark.parametrize("tech", [0, 1, 2])
def test_any_hit_rays_agree_with_closest_hit_on_occlusion(O, V, tech):
    """Terminate-on-first-hit returns SOME accepted hit: it exists iff the closest-hit search finds one, it lies in the
    ray interval, and it is never closer than the closest hit."""
    pos, idx = V.generate_groom(200, 8, V.GROOM_CURLY)
    sc = O.OracleScene(pos, idx, technique=tech)
    rng = np.random.default_rng(7)
    n = 3000
    o = (rng.normal(0, 1, (n, 3)) * 6 + (0, 152, 0)).astype(np.float32)
    tgt = pos[rng.integers(0, pos.shape[0], n)] + rng.normal(0, 0.03, (n, 3))
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays = np.concatenate([o, np.full((n, 1), 1e-3, np.float32), d, rng.uniform(2, 30, (n, 1)).astype(np.float32)], axis=1)
    hc = sc.trace_rays(rays)
    ha = sc.trace_rays(rays, any_hit=True)
    occ_c, occ_a = (hc["flags"] & 1).astype(bool), (ha["flags"] & 1).astype(bool)
    assert occ_c.sum() > 300 and (~occ_c).sum() > 50
    assert np.array_equal(occ_c, occ_a)
    assert (ha["t"][occ_a] >= hc["t"][occ_a]).all() and (ha["t"][occ_a] <= rays[occ_a, 7]).all()
    assert (ha["t"][occ_a] == hc["t"][occ_a]).mean() > 0.5      # nearest-first order usually finds the closest one first


def test_ao_image_is_the_base_image_times_visibility(O, V):
    pos, idx = V.generate_groom(400, 12, V.GROOM_CURLY)
    sc = O.OracleScene(pos, idx, technique=1)
    W, H = 64, 48
    vi, pi = V.camera_matrices(aspect=W / H)
    h0, i0, s0 = sc.render(O.make_frame(vi, pi, W, H, miss_rgb=(0.1, 0.2, 0.3)), stats=True)
    h1, i1, s1 = sc.render(O.make_frame(vi, pi, W, H, miss_rgb=(0.1, 0.2, 0.3), ao_samples=4), stats=True)
    assert h0.tobytes() == h1.tobytes()                              # hit records stay the primary hits
    hit = (h0["flags"] & 1).astype(bool)
    assert hit.sum() > 200
    assert np.array_equal(i0[~hit], i1[~hit])                        # miss pixels untouched
    assert (i1[hit, :3] <= i0[hit, :3]).all() and (i1[hit, :3] < i0[hit, :3]).any() and (i1[:, 3] == i0[:, 3]).all()
    assert s1["rays"] == s0["rays"] + 4 * hit.sum()
    # every AO pixel value is one of the 5 possible visibilities times the base colour
    base = np.float32(0.3) + np.abs(h0["ny"][hit])[:, None] * np.float32([0.4, 0.2, 0.1])
    levels = np.stack([np.clip(base * np.float32(1 - k / 4), 0, 1) * 255 + 0.5 for k in range(5)]).astype(np.uint8)   # [5, n, 3]
    assert (levels == i1[hit, :3][None]).all(axis=2).any(axis=0).all()
    # ao_distance -> 0: nothing is occluded
    _, i2, _ = sc.render(O.make_frame(vi, pi, W, H, miss_rgb=(0.1, 0.2, 0.3), ao_samples=4, ao_distance=1e-3))
    assert np.array_equal(i2, i0)


def test_oracle_inputs_equal_the_products(V, O):
    """The checker generates its own grooms and cameras (so bench.py --impl reference never loads the product library);
    they must be the product's, bit for bit (one seeded generator shared by oracle and GPU)."""
    for style in (V.GROOM_STRAIGHT, V.GROOM_CURLY):
        for n, segs in ((1, 1), (7, 3), (1000, 16), (5000, 32)):
            a, b = V.generate_groom(n, segs, style), O.generate_groom(n, segs, style)
            assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()
        a, b = V.generate_groom(100, 8, style, seed=12345), O.generate_groom(100, 8, style, seed=12345)
        assert a[0].tobytes() == b[0].tobytes()
    for kw in (dict(), dict(aspect=1.0), dict(aspect=float(np.float32(3840) / np.float32(2160))),
               dict(position=(1.0, 140.0, 30.0), yaw=-80.0, pitch=10.0, fov=45.0, aspect=1.5, near=0.5, far=500.0)):
        a, b = V.camera_matrices(**kw), O.camera_matrices(**kw)
        assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes(), kw


def test_oracle_build_is_thread_count_independent(O):
    """The oracle's LBVH build is OpenMP-parallel for the at-size scenes; nodes must not depend on the thread count."""
    import subprocess, sys, os, hashlib
    code = ("import sys, hashlib; sys.path.insert(0, %r)\n"
            "from oracle import oracle as O\n"
            "pos, idx = O.generate_groom(3000, 16, 1)\n"
            "h = hashlib.sha256()\n"
            "for t in (0, 1, 2):\n"
            "    n, i, m, l = O.OracleScene(pos, idx, technique=t).bvh()\n"
            "    h.update(n.tobytes()); h.update(i.tobytes()); h.update(m.tobytes()); h.update(l.tobytes())\n"
            "print(h.hexdigest())\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = set()
    for nt in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=nt)
        outs.add(subprocess.check_output([sys.executable, "-c", code], env=env).decode().strip())
    assert len(outs) == 1, outs


def test_tapered_prhi_constant_radius_is_the_reference_loop(O):
    """Per-vertex radius: with r0 == r1 the tapered Prhi must return the constant-radius bits (cone.slant = 0, as the reference calls it)."""
    rng = np.random.default_rng(7)
    curve = np.array([[0, 0, 0], [1 / 3, 0.02, 0], [2 / 3, -0.02, 0.01], [1, 0, 0]], np.float32)
    for _ in range(200):
        o = np.array([rng.uniform(0, 1), rng.uniform(-0.03, 0.03), 20.0], np.float32)
        d = np.array([rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01), -1.0], np.float32)
        d /= np.linalg.norm(d).astype(np.float32)
        a = O.prhi(o, d, curve, 0.02)
        b = O.prhi_taper(o, d, curve, 0.02, 0.02)
        assert np.float32(a[0]).tobytes() == np.float32(b[0]).tobytes() and np.float32(a[1]).tobytes() == np.float32(b[1]).tobytes()
        assert a[2].tobytes() == b[2].tobytes() and a[3] == b[3]


def test_tapered_prhi_vs_analytic_cone(O):
    """Straight curve (0,0,0)->(1,0,0) with radius 0.02 -> 0.005: the surface is the cone |yz| = r(x) = 0.02 - 0.015 x.
    fp64 analytic ray/cone roots are the pin (the reference only ever runs slant = 0)."""
    curve = np.array([[0, 0, 0], [1 / 3, 0, 0], [2 / 3, 0, 0], [1, 0, 0]], np.float32)
    r0, r1 = 0.02, 0.005
    rng = np.random.default_rng(11)
    hits = close = 0
    for _ in range(600):
        o = np.array([rng.uniform(0.05, 0.95), rng.uniform(-0.025, 0.025), 5.0], np.float64)
        d = np.array([rng.uniform(-0.02, 0.02), rng.uniform(-0.002, 0.002), -1.0], np.float64)
        d /= np.linalg.norm(d)
        # |(y, z)(s)|^2 = r(x(s))^2, r(x) = r0 + (r1 - r0) x
        k = r1 - r0
        A = d[1] ** 2 + d[2] ** 2 - (k * d[0]) ** 2
        B = 2 * (o[1] * d[1] + o[2] * d[2] - k * d[0] * (r0 + k * o[0]))
        Cq = o[1] ** 2 + o[2] ** 2 - (r0 + k * o[0]) ** 2
        disc = B * B - 4 * A * Cq
        t_ref = None
        if disc > 0:
            s = (-B - np.sqrt(disc)) / (2 * A)
            x = o[0] + s * d[0]
            if 0.0 < x < 1.0 and s > 0:
                t_ref = s
        t, u, n, _ = O.prhi_taper(o.astype(np.float32), d.astype(np.float32), curve, r0, r1)
        if t_ref is not None and disc > 1e-7:          # clear of grazing
            hits += 1
            if t > 0 and abs(t - t_ref) / t_ref < 1e-4:
                close += 1
                x = o[0] + t_ref * d[0]
                assert abs(u - x) < 2e-3            # curve parameter == x on this curve (the cone's axial offset is second order)
    assert hits > 100 and close >= 0.97 * hits, (hits, close)
