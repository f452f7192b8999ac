"""A SECOND, independent restatement of the reference GLSL for the Phantom path, in numpy float32
scalars, used only to cross-check oracle/vkhrt_oracle.cpp (tests/test_oracle_pins.py).

It follows the shader text literally where the oracle takes a documented shortcut:
  * CreateRCCMatrix builds the 4x4 matrix and calls a GENERIC inverse(mat4) (ray.glsl:19-33) —
    here np.linalg.inv — while the oracle applies the exact rigid inverse;
  * TransformCurve multiplies homogeneous points by that matrix (curve.glsl:33-42).
Everything else is the same statement order as the shader (hair_intersection.rint:15-150,
cone.glsl:21-62, cylinder.glsl:8-46, curve.glsl:9-47).  Pure-Python loops: small cases only.
"""
import numpy as np

F = np.float32


def v(*a):
    return np.array(a, dtype=F)


def dot(a, b):
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def cross(a, b):
    return v(F(a[1] * b[2]) - F(a[2] * b[1]), F(a[2] * b[0]) - F(a[0] * b[2]), F(a[0] * b[1]) - F(a[1] * b[0]))


def length(a):
    return F(np.sqrt(dot(a, a)))


def normalize(a):
    return (a * F(F(1.0) / np.sqrt(dot(a, a)))).astype(F)


def sample_curve_point(c, t):                       # curve.glsl:9-21
    t = F(t)
    u = F(1.0) - t
    tt = t * t
    uu = u * u
    uuu = uu * u
    ttt = tt * t
    return (uuu * c[0] + F(3.0) * uu * t * c[1] + F(3.0) * u * tt * c[2] + ttt * c[3]).astype(F)


def sample_curve_axis(c, t):                        # curve.glsl:23-31
    t = F(t)
    u = F(1.0) - t
    return (F(-3.0) * u * u * c[0] + F(3.0) * (F(3.0) * t * t - F(4.0) * t + F(1.0)) * c[1]
            + F(3.0) * (F(2.0) - F(3.0) * t) * t * c[2] + F(3.0) * t * t * c[3]).astype(F)


def curve_distance_to_cylinder(c, p):               # curve.glsl:44-47
    return F(length(cross(p - c[0], p - c[3])) / length(c[3] - c[0]))


def ray_cylinder_intersect(ro, rd, p0, p1, radius):  # cylinder.glsl:8-46
    ba = p1 - p0
    oc = ro - p0
    baba = dot(ba, ba)
    bard = dot(ba, rd)
    baoc = dot(ba, oc)
    k2 = F(baba - F(bard * bard))
    k1 = F(F(baba * dot(oc, rd)) - F(baoc * bard))
    k0 = F(F(F(baba * dot(oc, oc)) - F(baoc * baoc)) - F(F(radius * radius) * baba))
    h = F(F(k1 * k1) - F(k2 * k0))
    if h < 0.0:
        return False
    h = F(np.sqrt(h))
    with np.errstate(all="ignore"):
        t = F(F(-k1 - h) / k2)
        y = F(baoc + F(t * bard))
        if y > 0.0 and y < baba:
            return True
        t = F(F((F(0.0) if y < 0.0 else baba) - baoc) / bard)
        return bool(abs(F(k1 + F(k2 * t))) < h)


def create_rcc_matrix(ro, rd):                      # ray.glsl:13-33, generic inverse
    e3 = normalize(rd)
    w = e3
    e2 = normalize(v(-w[2], 0.0, w[0])) if abs(w[0]) > abs(w[1]) else normalize(v(0.0, w[2], -w[1]))
    e1 = cross(e2, w)
    m = np.zeros((4, 4), np.float64)                # columns e1 e2 e3 origin
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = e1, e2, e3, ro
    m[3, 3] = 1.0
    return np.linalg.inv(m).astype(F)


def transform_curve(c, m):                          # curve.glsl:33-42
    return [(m @ np.array([p[0], p[1], p[2], 1.0], F))[:3].astype(F) for p in c]


def ray_cone_intersect_rcc(center, radius, axis, slant):   # cone.glsl:21-62
    r2 = F(radius * radius)
    drr = F(radius * slant)
    ddd = F(F(axis[0] * axis[0]) + F(axis[1] * axis[1]))
    dp = F(F(center[0] * center[0]) + F(center[1] * center[1]))
    cdd = F(F(center[0] * axis[0]) + F(center[1] * axis[1]))
    cxd = F(F(center[0] * axis[1]) - F(center[1] * axis[0]))
    c = ddd
    b = F(axis[2] * F(drr - cdd))
    cdz2 = F(axis[2] * axis[2])
    ddd = F(ddd + cdz2)
    a = F(F(F(F(F(2.0) * drr) * cdd + F(cxd * cxd)) - F(ddd * r2)) + F(dp * cdz2))
    det = F(F(b * b) - F(a * c))
    real = bool(det > 0.0)
    with np.errstate(all="ignore"):
        s = F(F(b - (F(np.sqrt(det)) if real else F(0.0))) / c)
        dt = F(F(F(s * axis[2]) - cdd) / ddd)
    return real, s, dt


def prhi(ro, rd, curve, radius=0.02):
    """hair_intersection.rint:15-130 -> (t, u, normal, cone iterations); t == 0 means nothing reported."""
    ro, rd = np.asarray(ro, F), np.asarray(rd, F)
    c = [np.asarray(p, F) for p in np.asarray(curve, F).reshape(4, 3)]
    radius = F(radius)
    result, u_out, n_out, iters = F(0.0), F(0.0), v(0, 0, 0), 0
    with np.errstate(all="ignore"):
        rmax = F(curve_distance_to_cylinder(c, sample_curve_point(c, 0.5)) + radius)
        if not ray_cylinder_intersect(ro, rd, c[0], c[3], rmax):
            return float(result), float(u_out), n_out, iters
        rc = transform_curve(c, create_rcc_matrix(ro, rd))
        cd = normalize(rc[3] - rc[0])
        t_start = F(0.0) if dot(cd, v(0, 0, 1)) > 0.0 else F(1.0)
        for _side in range(2):
            t = t_start
            told = dt1 = dt2 = F(0.0)
            for i in range(8):
                iters += 1
                center = sample_curve_point(rc, t)
                axis = sample_curve_axis(rc, t)
                real, s, dt = ray_cone_intersect_rcc(center, radius, axis, F(0.0))
                if real and abs(dt) < F(5e-5):
                    result = F(s + center[2])
                    hit = (ro + result * rd).astype(F)
                    n_out = normalize(hit - sample_curve_point(c, t))
                    u_out = t
                    break
                dt = min(dt, F(0.5))
                dt = max(dt, F(-0.5))
                dt1 = dt2
                dt2 = dt
                if F(dt1 * dt2) < 0.0:
                    tnext = F(F(0.5) * F(told + t)) if (i & 3) == 0 else F(F(F(dt2 * told) - F(dt1 * t)) / F(dt2 - dt1))
                    told = t
                    t = tnext
                else:
                    told = t
                    t = F(t + dt)
                if t < 0.0 or t > 1.0:
                    break
            if result > 0.0:
                break
            t_start = F(F(1.0) - t_start)
    return float(result), float(u_out), n_out, iters


def analytic_cylinder(ro, rd, p0, p1, r):
    """fp64 infinite-precision-ish ray vs finite open cylinder (no caps): smallest t > 0 or None."""
    ro, rd, p0, p1 = (np.asarray(x, np.float64) for x in (ro, rd, p0, p1))
    ba = p1 - p0
    oc = ro - p0
    baba, bard, baoc = ba @ ba, ba @ rd, ba @ oc
    k2 = baba * (rd @ rd) - bard * bard
    k1 = baba * (oc @ rd) - baoc * bard
    k0 = baba * (oc @ oc) - baoc * baoc - r * r * baba
    h = k1 * k1 - k2 * k0
    if h < 0 or k2 == 0:
        return None
    t = (-k1 - np.sqrt(h)) / k2
    y = baoc + t * bard
    if t > 0 and 0 < y < baba:
        return t, y / baba
    return None


# ---- shaders/miss.rmiss:17-38, in float64 (an analytic check of the fp32 oracle, not a bit-level one) --------------
def miss_shader(env, ray_d):
    """env: [h, w, 4] float32 equirectangular map; sampler = linear, repeat (gpu_resources.hpp:46-50)"""
    d = -np.asarray(ray_d, np.float64)
    d = d / np.sqrt((d * d).sum())                                   # :27 normalize(-gl_WorldRayDirectionEXT)
    gamma = np.arcsin(np.clip(d[1], -1.0, 1.0))                      # :19
    theta = np.arctan2(d[0], -d[2])                                  # :20
    u = theta * 0.3183098861837 * 0.5 + 0.5                          # :22
    v = gamma * 0.3183098861837 + 0.5
    h, w = env.shape[:2]
    x, y = u * w - 0.5, v * h - 0.5                                  # Vulkan unnormalised texel coordinates
    x0, y0 = int(np.floor(x)), int(np.floor(y))
    fx, fy = x - x0, y - y0
    e = env.astype(np.float64)
    t = lambda i, j: e[j % h, i % w, :3]
    c = (1 - fy) * ((1 - fx) * t(x0, y0) + fx * t(x0 + 1, y0)) + fy * ((1 - fx) * t(x0, y0 + 1) + fx * t(x0 + 1, y0 + 1))
    c = 1.0 - np.exp(-c * 1.0)                                       # :34 exposure 1
    return c ** (1.0 / 2.2)                                          # :35


# ---- geometry_processor.cpp:69-121, 158-197, transcribed statement by statement over Python lists ----------------
def merge_lines(lines):
    """lines: list of (start, end) float32 triples"""
    new = []
    wanted = len(lines) // 2                                         # :73
    for i in range(wanted):                                          # :76
        old = i * 2
        l1 = lines[old]
        if old + 1 == len(lines):                                    # :82 (never true inside this loop)
            new.append(l1)
            break
        l2 = lines[old + 1]
        if not np.array_equal(l1[1], l2[0]):                         # :91 glm::vec3 !=
            new.append(l1); new.append(l2)
            continue
        new.append((l1[0], l2[1]))                                   # :98-100
    return new


def split_lines(lines):
    new = []
    for s, e in lines:                                               # :113
        mid = ((s + e) * F(0.5)).astype(F)                           # :115
        new.append((s, mid)); new.append((mid, e))
    return new


def merge_curves_fast(curves):
    """curves: list of [start, cp1, cp2, end] float32 triples"""
    new = []
    wanted = len(curves) // 2
    for i in range(wanted):
        c1 = curves[2 * i]
        if 2 * i + 1 == len(curves):
            new.append(c1)
            break
        c2 = curves[2 * i + 1]
        if not np.array_equal(c1[3], c2[0]):                         # :180
            new.append(c1); new.append(c2)
            continue
        mid = ((c1[2] + c2[1]) * F(0.5)).astype(F)                   # :190
        new.append([c1[0], ((c1[1] + mid) * F(0.5)).astype(F), ((mid + c2[2]) * F(0.5)).astype(F), c2[3]])   # :187-192
    return new
