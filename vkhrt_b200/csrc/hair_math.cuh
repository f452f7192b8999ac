// hair_math.cuh — device-side ray/hair-segment math for the sm_100a traversal kernels.
//
// Follows the reference GLSL (paths relative to the reference root):
//   shaders/curve.glsl:9-47, cylinder.glsl:8-46, ray.glsl:13-33, cone.glsl:21-62,
//   hair_intersection.rint:15-150, ray_gen.rgen:16-48, shading.glsl:1-11, debug.glsl:1-7.
//
// Floating-point contract: compiled with -fmad=false and the default -prec-div=true
// -prec-sqrt=true, so every fp32 operation below is ONE IEEE round-to-nearest operation in exactly
// the order written, and the only fused multiply-adds are the explicit fmaf() calls.
//   * shader-side math (everything that follows a .glsl/.rint/.rgen/.rchit file, plus the LSS /
//     triangle / slab tests defined by this project) is written with explicit fmaf() where an
//     FMA-contracting GLSL compiler fuses a*b + c  -> fdot3 / fcross3 / fnormalize3 / fmadd3;
//   * host-side math (geometry_processor.cpp restatements used by build.cu) is unfused
//     -> dot3 / cross3 / normalize3.
// The CPU oracle follows the same contract, so results are a pure function of (ray, primitive),
// bit-identical on both sides, and do not depend on traversal order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vkhrt {

#define VK_DEV __device__ __forceinline__

VK_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
VK_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
VK_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
VK_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
VK_DEV float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
VK_DEV float dot3(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
VK_DEV float3 cross3(float3 a, float3 b)
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
VK_DEV float3 normalize3(float3 a)
{
    float inv = 1.0f / sqrtf(dot3(a, a));
    return a * inv;
}
VK_DEV float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }
// fused (shader-side) helpers
VK_DEV float fdot3(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
VK_DEV float3 fcross3(float3 a, float3 b)
{
    return f3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
VK_DEV float3 fnormalize3(float3 a)
{
    float inv = 1.0f / sqrtf(fdot3(a, a));
    return a * inv;
}
VK_DEV float3 fmadd3(float s, float3 a, float3 b) { return f3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }   // s*a + b

// ---- cubic Bezier in Bernstein form: shaders/curve.glsl:9-31 ---------------------------------
struct Bezier { float3 p0, p1, p2, p3; };

VK_DEV float3 bezier_point(const Bezier& c, float t)
{
    float u = 1.0f - t;
    float tt = t * t;
    float uu = u * u;
    float w0 = uu * u;
    float w3 = tt * t;
    float w1 = (3.0f * uu) * t;
    float w2 = (3.0f * u) * tt;
    return fmadd3(w3, c.p3, fmadd3(w2, c.p2, fmadd3(w1, c.p1, w0 * c.p0)));
}

VK_DEV float3 bezier_axis(const Bezier& c, float t)
{
    float u = 1.0f - t;
    float w0 = (-3.0f * u) * u;
    float w1 = 3.0f * (fmaf(3.0f * t, t, -(4.0f * t)) + 1.0f);
    float w2 = (3.0f * fmaf(-3.0f, t, 2.0f)) * t;
    float w3 = (3.0f * t) * t;
    return fmadd3(w3, c.p3, fmadd3(w2, c.p2, fmadd3(w1, c.p1, w0 * c.p0)));
}

// rmax of Prhi (hair_intersection.rint:21-22): chord distance of B(0.5) + radius.  A function of the
// curve alone, so the geometry kernel evaluates it once per curve instead of once per candidate.
VK_DEV float bezier_bound_radius(const Bezier& c, float radius)
{
    float3 m = bezier_point(c, 0.5f);
    float3 x = fcross3(m - c.p0, m - c.p3);
    float3 ch = c.p3 - c.p0;
    float r = sqrtf(fdot3(x, x)) / sqrtf(fdot3(ch, ch));
    return r + radius;
}

// ---- conservative candidate filter (NOT in the reference; result-neutral by construction) -----
// Prhi reports a hit only from a REAL cone intersection with |dt| < 5e-5 (hair_intersection.rint:67).
// In ray-centric coordinates the hit point h = (0,0,c.z+s) then satisfies |h - B(t)|^2 = r^2 + dt^2 |B'(t)|^2
// (cone.glsl:21-62 with slant 0), so the ray passes within r*(1 + 1e-6) of the curve point B(t), t in [0,1].
// Every B(t) lies within `dev` of one of the sub-chords [B(k/4),B((k+1)/4)] (convex-hull property of the
// de Casteljau pieces), and projecting onto the plane orthogonal to the ray cannot increase
// distances.  Hence: if the ray's projection (the origin of the ray-centric xy-plane) is farther than
// r + dev (+ rounding slack) from ALL four projected quarter-chords, the march cannot report anything and
// is skipped.  bezier_quarter_chord_deviation() (= sagitta/16 of the whole curve) is evaluated once per curve
// by the build kernel.
VK_DEV float point_segment_distance2(float3 p, float3 a, float3 b)      // squared
{
    float3 e = b - a, q = p - a;
    float ee = fdot3(e, e);
    float t = fminf(fmaxf(fdot3(q, e), 0.0f), ee);
    float s = ee > 0.0f ? t / ee : 0.0f;
    float3 r = q - s * e;
    return fdot3(r, r);
}
// de Casteljau split of a cubic at t = 1/2
VK_DEV void bezier_split(const Bezier& c, Bezier& l, Bezier& r)
{
    float3 q01 = 0.5f * (c.p0 + c.p1), q12 = 0.5f * (c.p1 + c.p2), q23 = 0.5f * (c.p2 + c.p3);
    float3 r0 = 0.5f * (q01 + q12), r1 = 0.5f * (q12 + q23);
    float3 m = 0.5f * (r0 + r1);
    l.p0 = c.p0; l.p1 = q01; l.p2 = r0; l.p3 = m;
    r.p0 = m; r.p1 = r1; r.p2 = q23; r.p3 = c.p3;
}
// max distance of the curve from its chord SEGMENT (hull bound), SQUARED.  The oracle takes the maximum of the eight distances; sqrtf is
// monotone (and correctly rounded), so the root of the largest square is bit-for-bit the largest root: one sqrtf per curve instead of eight.
VK_DEV float bezier_chord_deviation2(const Bezier& c)
{
    return fmaxf(point_segment_distance2(c.p1, c.p0, c.p3), point_segment_distance2(c.p2, c.p0, c.p3));
}
// An accepted Prhi hit h lies sqrt(r^2 + (dt |B'(t)|)^2) from B(t) with |dt| < 5e-5 (hair_intersection.rint:67), i.e. up to
// (5e-5 max|B'|)^2 / (2 r) farther than r.  max|B'| <= 3 max|p[i+1] - p[i]| (hull of the derivative's control points).  For hair-like
// segments (length L < ~900 r) the 0.1 % inflation of the filter's bound covers it; this term makes thin or long segments safe too.
VK_DEV float bezier_convergence_slack(const Bezier& c, float radius)
{
    float3 d0 = c.p1 - c.p0, d1 = c.p2 - c.p1, d2 = c.p3 - c.p2;
    float l2 = fmaxf(fmaxf(fdot3(d0, d0), fdot3(d1, d1)), fdot3(d2, d2));         // max |p[i+1] - p[i]|^2
    float x2 = (9.0f * l2) * (5e-5f * 5e-5f);                                      // (5e-5 max|B'|)^2
    return (x2 / (2.0f * radius)) * 1.01f;
}
// max deviation of the four quarter-curves from their chords [B(k/4), B((k+1)/4)], inflated
VK_DEV float bezier_quarter_chord_deviation(const Bezier& c)
{
    Bezier l, r, a, b;
    bezier_split(c, l, r);
    bezier_split(l, a, b);
    float dv2 = fmaxf(bezier_chord_deviation2(a), bezier_chord_deviation2(b));
    bezier_split(r, a, b);
    dv2 = fmaxf(dv2, fmaxf(bezier_chord_deviation2(a), bezier_chord_deviation2(b)));
    return sqrtf(dv2) * 1.001f + 1e-7f;      // the bound must stay conservative under fp32 rounding
}
// is the origin farther than sqrt(b2) from the 2-D segment [a, b]?  Division-free: compares d^2 * |e|^2 with b2 * |e|^2.
// Any NaN makes the comparison false (= "not farther").
VK_DEV bool origin_farther_than(float ax, float ay, float bx, float by, float b2)
{
    float ex = bx - ax, ey = by - ay;
    float ee = fmaf(ex, ex, ey * ey);
    float aa = fmaf(ax, ax, ay * ay), bb = fmaf(bx, bx, by * by);
    float t = -fmaf(ax, ex, ay * ey);                       // projection parameter * ee
    float num = t <= 0.0f ? aa * ee : (t >= ee ? bb * ee : fmaf(aa, ee, -(t * t)));
    return num > b2 * ee * 1.0001f;                         // rounding slack on the products
}
// c = curve in ray-centric coordinates.  true => the march may report a hit (NaNs never reject).
VK_DEV bool quarter_chords_near_ray(const Bezier& c, float radius, float dev)
{
    // B(1/4), B(1/2), B(3/4) projected on the xy-plane
    float x1 = (1.0f / 64.0f) * (fmaf(27.0f, c.p0.x + c.p1.x, c.p3.x) + 9.0f * c.p2.x);
    float y1 = (1.0f / 64.0f) * (fmaf(27.0f, c.p0.y + c.p1.y, c.p3.y) + 9.0f * c.p2.y);
    float x2 = 0.125f * ((c.p0.x + c.p3.x) + 3.0f * (c.p1.x + c.p2.x));
    float y2 = 0.125f * ((c.p0.y + c.p3.y) + 3.0f * (c.p1.y + c.p2.y));
    float x3 = (1.0f / 64.0f) * (fmaf(27.0f, c.p2.x + c.p3.x, c.p0.x) + 9.0f * c.p1.x);
    float y3 = (1.0f / 64.0f) * (fmaf(27.0f, c.p2.y + c.p3.y, c.p0.y) + 9.0f * c.p1.y);
    float bound = (radius + dev) * 1.001f + 4e-6f * (fabsf(c.p0.z) + fabsf(c.p3.z)) + 1e-6f;
    float b2 = bound * bound;
    return !(origin_farther_than(c.p0.x, c.p0.y, x1, y1, b2) && origin_farther_than(x1, y1, x2, y2, b2) &&
             origin_farther_than(x2, y2, x3, y3, b2) && origin_farther_than(x3, y3, c.p3.x, c.p3.y, b2));
}

// ---- shaders/cylinder.glsl:8-46 (boolean, no t>0 test, assumes |d| = 1) ------------------------
VK_DEV bool ray_hits_cylinder(float3 o, float3 d, float3 a, float3 b, float radius)
{
    float3 ba = b - a;
    float3 oc = o - a;
    float baba = fdot3(ba, ba);
    float bard = fdot3(ba, d);
    float baoc = fdot3(ba, oc);
    float k2 = fmaf(-bard, bard, baba);
    float k1 = fmaf(baba, fdot3(oc, d), -(baoc * bard));
    float k0 = fmaf(-(radius * radius), baba, fmaf(baba, fdot3(oc, oc), -(baoc * baoc)));
    float h = fmaf(k1, k1, -(k2 * k0));
    if (h < 0.0f) return false;
    h = sqrtf(h);
    float t = (-k1 - h) / k2;
    float y = fmaf(t, bard, baoc);
    if (y > 0.0f && y < baba) return true;
    t = ((y < 0.0f ? 0.0f : baba) - baoc) / bard;
    return fabsf(fmaf(k2, t, k1)) < h;
}

// ---- ray-centric frame: shaders/ray.glsl:13-33 --------------------------------------------------
// Depends on the ray only, so it is built once per ray (the reference rebuilds it per candidate).
struct RayFrame { float3 e1, e2, e3; };

VK_DEV RayFrame make_ray_frame(float3 d)
{
    RayFrame f;
    f.e3 = fnormalize3(d);
    float3 w = f.e3;
    f.e2 = fabsf(w.x) > fabsf(w.y) ? fnormalize3(f3(-w.z, 0.0f, w.x)) : fnormalize3(f3(0.0f, w.z, -w.y));
    f.e1 = fcross3(f.e2, w);
    return f;
}
// inverse of the rigid matrix [e1 e2 e3 o] applied to a point (curve.glsl:33-42 + ray.glsl:32)
VK_DEV float3 into_frame(const RayFrame& f, float3 o, float3 p)
{
    float3 q = p - o;
    return f3(fdot3(f.e1, q), fdot3(f.e2, q), fdot3(f.e3, q));
}

// ---- shaders/cone.glsl:21-62, specialised to what Prhi reads (s, dt, real/phantom) --------------
struct ConeStep { float s, dt; bool real; };

VK_DEV ConeStep cone_step(float3 c, float radius, float3 ax, float slant)
{
    float r2 = radius * radius;
    float drr = radius * slant;
    float ddd = fmaf(ax.y, ax.y, ax.x * ax.x);
    float dp = fmaf(c.y, c.y, c.x * c.x);
    float cdd = fmaf(c.y, ax.y, c.x * ax.x);
    float cxd = fmaf(c.x, ax.y, -(c.y * ax.x));
    float qc = ddd;
    float qb = ax.z * (drr - cdd);
    float cdz2 = ax.z * ax.z;
    ddd += cdz2;
    float qa = fmaf(dp, cdz2, fmaf(-ddd, r2, fmaf(2.0f * drr, cdd, cxd * cxd)));
    float det = fmaf(qb, qb, -(qa * qc));
    ConeStep r;
    r.real = det > 0.0f;
    r.s = (qb - (r.real ? sqrtf(det) : 0.0f)) / qc;
    r.dt = fmaf(r.s, ax.z, -cdd) / ddd;
    return r;
}

// ---- Phantom Ray-Hair Intersector: hair_intersection.rint:35-130 (after the cylinder early-out) --
// The reference's two nested loops (2 sides x <= 8 cone iterations) are unrolled into a resumable
// state machine so that the traversal kernel can run ONE cone iteration per scheduling step with
// whatever lanes currently hold a candidate (trace.cu).  march_begin = :38-44, march_step = one
// pass through the loop body :56-115 plus the side switch :118-126.
struct MarchState {
    Bezier c;            // curve in ray-centric coordinates (TransformCurve, curve.glsl:33-42)
    float t, told, dt1, dt2;
    float t_start;
    uint32_t it;         // bits 0..3: iteration i of this side; bit 4: side
    float r0, dr;        // per-vertex radius (TAPER kernels only): radius(t) = r0 + t * dr, cone slant = dr
};
enum MarchResult { MARCH_CONTINUE = 0, MARCH_HIT = 1, MARCH_NOTHING = 2 };

VK_DEV void march_begin(MarchState& m, const RayFrame& fr, float3 o, const Bezier& world)
{
    m.c.p0 = into_frame(fr, o, world.p0);
    m.c.p1 = into_frame(fr, o, world.p1);
    m.c.p2 = into_frame(fr, o, world.p2);
    m.c.p3 = into_frame(fr, o, world.p3);
    float3 chord = m.c.p3 - m.c.p0;
    float cz = chord.z * (1.0f / sqrtf(fdot3(chord, chord)));   // z of normalize(chord) == dot(., (0,0,1))
    m.t_start = cz > 0.0f ? 0.0f : 1.0f;
    m.t = m.t_start;
    m.told = m.dt1 = m.dt2 = 0.0f;
    m.it = 0u;
}

// One cone iteration.  On MARCH_HIT *t_hit is `result` (may be <= 0: hair_intersection.rint:146 is applied
// by the caller) and *u_hit the converged curve parameter.
// A side is also left when the march has reached an exact fp32 fixed point (t + dt == t on the plain-step
// branch): every later iteration of that side would recompute the very same state, so the outcome is
// unchanged.
// TAPER: the curve's radius runs linearly from r0 (t = 0) to r0 + dr (t = 1): cone.radius = r(t), cone.slant = dr/dt = dr, the taper
// term of cone.glsl:27 (`drr = radius * slant`) that the reference's caller leaves at 0 (hair_intersection.rint:62).
template <bool TAPER = false>
VK_DEV int march_step(MarchState& m, float radius, float* t_hit, float* u_hit)
{
    const uint32_t i = m.it & 15u;
    float3 centre = bezier_point(m.c, m.t);
    float3 axis = bezier_axis(m.c, m.t);
    ConeStep cs = TAPER ? cone_step(centre, fmaf(m.t, m.dr, m.r0), axis, m.dr) : cone_step(centre, radius, axis, 0.0f);
    if (cs.real && fabsf(cs.dt) < 5e-5f) {
        *t_hit = cs.s + centre.z;
        *u_hit = m.t;
        if (*t_hit > 0.0f || (m.it & 16u)) return MARCH_HIT;     // :119 `if (result > 0.0) break;`
        // converged behind the origin on the first side: the reference keeps looking from the other end
    } else {
        float dt = cs.dt;
        dt = 0.5f < dt ? 0.5f : dt;      // GLSL min(dt, 0.5)
        dt = dt < -0.5f ? -0.5f : dt;    // GLSL max(dt, -0.5)
        m.dt1 = m.dt2;
        m.dt2 = dt;
        float tn;
        const bool plain = !(m.dt1 * m.dt2 < 0.0f);
        if (!plain) tn = (i & 3u) == 0u ? 0.5f * (m.told + m.t) : fmaf(m.dt2, m.told, -(m.dt1 * m.t)) / (m.dt2 - m.dt1);
        else tn = m.t + dt;
        m.told = m.t;
        const bool fixed_point = plain && tn == m.t;
        m.t = tn;
        if (!(tn < 0.0f || tn > 1.0f) && !fixed_point && i < 7u) { m.it++; return MARCH_CONTINUE; }
    }
    // this side is over without a positive result
    if (m.it & 16u) return MARCH_NOTHING;
    m.t_start = 1.0f - m.t_start;
    m.t = m.t_start;
    m.told = m.dt1 = m.dt2 = 0.0f;
    m.it = 16u;
    return MARCH_CONTINUE;
}

// ---- LSS: ray vs linear swept sphere (defined by this project; RT hardware in the reference) -----
// lss = {p0, r0, p1, r1}.  Returns hit flag; t, u on hit.  See DESIGN.md §4.4 for the derivation.
// kNormal: also evaluate the shading normal of UnpackLSSGeometry (triangle_closest_hit.rchit:43-58),
// normalize(hit - mix(p0,p1,u)), from the same re-originated operands (used once per ray, for the winner).
template <bool kNormal>
VK_DEV bool lss_intersect(float3 o, float3 d, float3 p0, float r0, float3 p1, float r1, float* t_out, float* u_out,
                          float3* n_out)
{
    float3 ba = p1 - p0;
    float3 oa0 = o - p0;
    float dd = fdot3(d, d);
    float t0 = (0.0f - fdot3(d, oa0)) / dd;
    float3 oa = fmadd3(t0, d, oa0);
    float m0 = fdot3(ba, ba), m1 = fdot3(ba, oa), m2 = fdot3(ba, d), m3 = fdot3(d, oa), m5 = fdot3(oa, oa);
    float rr = r0 - r1;
    float d2 = fmaf(-rr, rr, m0);
    float q0 = fmaf(-r0, r0, m5);
    float bt = 0.0f, bu = 0.0f;
    bool found = false;
    if (d2 > 0.0f) {
        float a1 = fmaf(-r0, rr, m1);
        float k2 = fmaf(d2, dd, -(m2 * m2));
        float k1 = fmaf(d2, m3, -(m2 * a1));
        float k0 = fmaf(d2, q0, -(a1 * a1));
        float h = fmaf(k1, k1, -(k2 * k0));
        if (h >= 0.0f) {
            float t = (-k1 - sqrtf(h)) / k2;
            float y = fmaf(t, m2, a1);
            if (y > 0.0f && y < d2) { bt = t; bu = y / d2; found = true; }
        }
    }
    if (!found) {
        float h1 = fmaf(m3, m3, -(dd * q0));
        if (h1 > 0.0f) { bt = (-m3 - sqrtf(h1)) / dd; bu = 0.0f; found = true; }
        float3 ob = oa - ba;
        float m6 = fdot3(d, ob), m7 = fdot3(ob, ob);
        float h2 = fmaf(m6, m6, -(dd * fmaf(-r1, r1, m7)));
        if (h2 > 0.0f) {
            float t = (-m6 - sqrtf(h2)) / dd;
            if (!found || t < bt) { bt = t; bu = 1.0f; found = true; }
        }
    }
    *t_out = bt + t0;
    *u_out = bu;
    if (kNormal) *n_out = fnormalize3(fmadd3(-bu, ba, fmadd3(bt, d, oa)));
    return found;
}

// ---- DOTS: ray vs triangle, Moeller-Trumbore, no culling (defined by this project) ---------------
VK_DEV bool tri_intersect(float3 o, float3 d, float3 v0, float3 v1, float3 v2, uint32_t parity, float* t_out, float* u_out)
{
    float3 e1 = v1 - v0, e2 = v2 - v0;
    float3 p = fcross3(d, e2);
    float det = fdot3(e1, p);
    if (det == 0.0f || det != det) return false;
    float inv = 1.0f / det;
    float3 tv = o - v0;
    float b1 = fdot3(tv, p) * inv;
    if (!(b1 >= 0.0f && b1 <= 1.0f)) return false;
    float3 q = fcross3(tv, e1);
    float b2 = fdot3(d, q) * inv;
    if (!(b2 >= 0.0f && b1 + b2 <= 1.0f)) return false;
    *t_out = fdot3(e2, q) * inv;
    *u_out = parity ? b2 : (b1 + b2);
    return true;
}
VK_DEV float3 tri_normal(float3 d, float3 v0, float3 v1, float3 v2)
{
    float3 n = fnormalize3(fcross3(v1 - v0, v2 - v0));
    if (fdot3(n, d) > 0.0f) n = f3(-n.x, -n.y, -n.z);
    return n;
}

// ---- DOTS strip: the 4 triangles of one segment, rebuilt from the 64-byte strip record --------------
// GenerateDisjointOrthogonalTriangleStrips (geometry_processor.cpp:221-271): face f in {0,1}, offset off_f = v_f * r,
//   triangle 2f   = (start+off, end-off, end+off)      triangle 2f+1 = (start+off, start-off, end-off)
// each vertex is ONE fp32 add/sub of the stored operands, exactly what the generator computes.
VK_DEV void strip_triangle(float3 s, float3 e, float3 off, uint32_t k, float3* v0, float3* v1, float3* v2)
{
    *v0 = s + off;
    if (k == 0u) { *v1 = e - off; *v2 = e + off; }
    else         { *v1 = s - off; *v2 = e - off; }
}
// per-vertex radius: the offsets at the start / end vertices are v * r0 / v * r1
VK_DEV void strip_triangle_taper(float3 s, float3 e, float3 offs, float3 offe, uint32_t k, float3* v0, float3* v1, float3* v2)
{
    *v0 = s + offs;
    if (k == 0u) { *v1 = e - offe; *v2 = e + offe; }
    else         { *v1 = s - offs; *v2 = e - offe; }
}
// Conservative strip reject (NOT in the reference; result-neutral by construction).  Every point of the four
// triangles is (a point of the segment) + c * off_f with |c| <= 1 and |off_f| = r, so a ray that hits any of them
// passes within r of the segment's axis LINE: |(start - o) . n| <= r |n| with n = d x (end - start).
// The bound is inflated by 1 % plus a rounding allowance that grows with the distance of the strip from the ray origin
// (fp32 barycentrics of a sliver seen from afar accept points slightly outside the triangle).  Parallel ray / NaN: never rejects.
VK_DEV bool ray_near_strip_axis(float3 o, float3 d, float3 s, float3 e, float radius)
{
    float3 a = e - s, so = s - o;
    float3 n = fcross3(d, a);
    float lhs = fabsf(fdot3(so, n));
    float nn = sqrtf(fdot3(n, n));
    float slack = 1e-5f * ((fabsf(so.x) + fabsf(so.y) + fabsf(so.z)) * (fabsf(a.x) + fabsf(a.y) + fabsf(a.z)));
    float rhs = fmaf(radius * 1.01f, nn, slack);
    return !(lhs > rhs);
}

// ---- primary ray: shaders/ray_gen.rgen:16-24 ------------------------------------------------------
struct Camera { float vi[16]; float pi[16]; };   // CameraUniformData, column-major

VK_DEV void primary_ray(const Camera& cam, uint32_t W, uint32_t H, uint32_t px, uint32_t py, float sx, float sy,
                        float3* o, float3* d)
{
    float pcx = (float)px + sx, pcy = (float)py + sy;
    float u = pcx / (float)W, v = pcy / (float)H;
    float dx = fmaf(u, 2.0f, -1.0f), dy = fmaf(v, 2.0f, -1.0f);
    *o = f3(cam.vi[12], cam.vi[13], cam.vi[14]);
    float3 tg;
    tg.x = (fmaf(cam.pi[4], dy, cam.pi[0] * dx) + cam.pi[8]) + cam.pi[12];
    tg.y = (fmaf(cam.pi[5], dy, cam.pi[1] * dx) + cam.pi[9]) + cam.pi[13];
    tg.z = (fmaf(cam.pi[6], dy, cam.pi[2] * dx) + cam.pi[10]) + cam.pi[14];
    float3 nd = fnormalize3(tg);
    d->x = fmaf(cam.vi[8], nd.z, fmaf(cam.vi[4], nd.y, cam.vi[0] * nd.x));
    d->y = fmaf(cam.vi[9], nd.z, fmaf(cam.vi[5], nd.y, cam.vi[1] * nd.x));
    d->z = fmaf(cam.vi[10], nd.z, fmaf(cam.vi[6], nd.y, cam.vi[2] * nd.x));
}

// ---- ambient-occlusion rays (NOT in the reference: the SURVEY.md §8(f) "secondary rays" row) ------
// One AO ray per (pixel, spp sample, ao index): origin = hit point pushed `bias` along the shading normal,
// direction = normalize(n + s) with s uniform on the unit sphere (cosine-weighted about n).  s comes from
// Marsaglia's 1972 rejection construction driven by an integer hash, so the whole thing needs only + * sqrt and is
// bit-identical on the CPU oracle and here (no sin/cos).
VK_DEV uint32_t hash32(uint32_t h)
{
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}
VK_DEV float hash_unit(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }   // [0,1), exact
VK_DEV float3 ao_direction(float3 n, uint32_t pixel, uint32_t sample, uint32_t index)
{
    uint32_t seed = hash32(pixel * 0x9E3779B1u + sample * 0x85EBCA77u + index * 0xC2B2AE3Du + 0x27D4EB2Fu);
    float x1 = 0.0f, x2 = 0.0f, S = 2.0f;
    for (uint32_t j = 0; j < 8u && !(S < 1.0f); ++j) {
        uint32_t a = hash32(seed + j * 0x9E3779B9u), b = hash32(a ^ 0x68E31DA4u);
        x1 = fmaf(2.0f, hash_unit(a), -1.0f);
        x2 = fmaf(2.0f, hash_unit(b), -1.0f);
        S = fmaf(x1, x1, x2 * x2);
    }
    if (!(S < 1.0f)) { x1 = 0.0f; x2 = 0.0f; S = 0.0f; }     // 8 rejections in a row (p ~ 4e-6): the pole
    float q = 2.0f * sqrtf(1.0f - S);
    float3 v = f3(fmaf(x1, q, n.x), fmaf(x2, q, n.y), fmaf(-2.0f, S, 1.0f) + n.z);
    float l2 = fdot3(v, v);
    if (!(l2 > 1e-8f)) return n;                             // s = -n
    return v * (1.0f / sqrtf(l2));
}

// ---- ray/box slab test in the fused form t = fma(plane, 1/d, -(o/d)) ------------------------------
// A zero direction component would make o*(1/d) infinite and the fused slab NaN: |1/d| is clamped to 1e20,
// for which the slab degenerates to the exact "is o inside [lo,hi]" test.
VK_DEV float safe_rcp(float d)
{
    float r = 1.0f / d;
    if (!(fabsf(r) <= 1e20f)) r = (d < 0.0f || (d == 0.0f && signbit(d))) ? -1e20f : 1e20f;
    return r;
}
VK_DEV bool slab_test(float3 lo, float3 hi, float3 id, float3 noid, float tmin, float tcur, float* tnear)
{
    float tx0 = fmaf(lo.x, id.x, noid.x), tx1 = fmaf(hi.x, id.x, noid.x);
    float ty0 = fmaf(lo.y, id.y, noid.y), ty1 = fmaf(hi.y, id.y, noid.y);
    float tz0 = fmaf(lo.z, id.z, noid.z), tz1 = fmaf(hi.z, id.z, noid.z);
    float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fmaxf(fminf(tz0, tz1), tmin));
    float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fminf(fmaxf(tz0, tz1), tcur));
    *tnear = tn;
    return tn <= tf;
}

// ---- closest-hit colour: shaders/shading.glsl:1-11, debug.glsl:1-7 --------------------------------
VK_DEV float3 shade_normal(float3 n)
{
    float k = fabsf(fdot3(n, f3(0.0f, -1.0f, 0.0f)));
    return f3(fmaf(k, 0.4f, 0.3f), fmaf(k, 0.2f, 0.3f), fmaf(k, 0.1f, 0.3f));
}
VK_DEV float3 debug_palette(uint32_t prim)
{
    uint32_t i = prim % 6u;
    // (1,0,.3) (.8,.2,.3) (.6,.4,.3) (.4,.6,.3) (.2,.8,.3) (0,1,.3)
    const float r[6] = {1.0f, 0.8f, 0.6f, 0.4f, 0.2f, 0.0f};
    const float g[6] = {0.0f, 0.2f, 0.4f, 0.6f, 0.8f, 1.0f};
    return f3(r[i], g[i], 0.3f);
}

// ---- miss colour: shaders/miss.rmiss:17-38 ----------------------------------------------------------
// Equirectangular lookup of normalize(-rayDirection) with a linear, repeat-addressed sampler (gpu_resources.hpp:46-50:
// texel coordinate u*W - 0.5, fp32 weights), then 1 - exp(-c * exposure) and gamma 1/2.2.  asinf/atan2f/expf/powf are the
// CUDA math library's (<= 2-4 ulp), so this is the one colour that is compared with the oracle to 1 LSB of the 8-bit
// image rather than bit for bit.
VK_DEV int wrap_texel(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
VK_DEV float3 environment_miss(const float4* __restrict__ env, uint32_t env_w, uint32_t env_h, float3 ray_d)
{
    const float3 dir = fnormalize3(f3(-ray_d.x, -ray_d.y, -ray_d.z));
    const float gy = fminf(fmaxf(dir.y, -1.0f), 1.0f);
    const float gamma = asinf(gy);
    const float theta = atan2f(dir.x, -dir.z);
    const float u = fmaf(theta * 0.3183098861837f, 0.5f, 0.5f);
    const float v = gamma * 0.3183098861837f + 0.5f;
    const int W = (int)env_w, H = (int)env_h;
    const float x = fmaf(u, (float)W, -0.5f), y = fmaf(v, (float)H, -0.5f);
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int i0 = wrap_texel((int)x0, W), i1 = wrap_texel((int)x0 + 1, W), j0 = wrap_texel((int)y0, H), j1 = wrap_texel((int)y0 + 1, H);
    const float4 t00 = __ldg(env + (size_t)j0 * W + i0), t10 = __ldg(env + (size_t)j0 * W + i1);
    const float4 t01 = __ldg(env + (size_t)j1 * W + i0), t11 = __ldg(env + (size_t)j1 * W + i1);
    const float a[3] = {t00.x, t00.y, t00.z}, b[3] = {t10.x, t10.y, t10.z}, c[3] = {t01.x, t01.y, t01.z}, e[3] = {t11.x, t11.y, t11.z};
    float out[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float top = fmaf(fx, b[k] - a[k], a[k]);
        const float bot = fmaf(fx, e[k] - c[k], c[k]);
        float r = fmaf(fy, bot - top, top);
        r = 1.0f - expf(-(r * 1.0f));
        out[k] = powf(r, 1.0f / 2.2f);
    }
    return f3(out[0], out[1], out[2]);
}

VK_DEV uint32_t to_unorm8(float c)
{
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint32_t)(c * 255.0f + 0.5f);
}

}  // namespace vkhrt
