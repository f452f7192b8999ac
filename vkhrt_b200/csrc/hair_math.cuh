// hair_math.cuh — device-side ray/hair-segment math for the sm_100a traversal kernels.
//
// Follows the reference GLSL (paths relative to the reference root):
//   shaders/curve.glsl:9-47, cylinder.glsl:8-46, ray.glsl:13-33, cone.glsl:21-62,
//   hair_intersection.rint:15-150, ray_gen.rgen:16-48, shading.glsl:1-11, debug.glsl:1-7.
//
// Floating-point contract: this translation unit is compiled with -fmad=false and
// the default -prec-div=true -prec-sqrt=true, so every fp32 operation below is one IEEE
// round-to-nearest operation in exactly the order written.  Results are therefore a pure
// function of (ray, primitive) and do not depend on traversal order.
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
    return ((w0 * c.p0 + w1 * c.p1) + w2 * c.p2) + w3 * c.p3;
}

VK_DEV float3 bezier_axis(const Bezier& c, float t)
{
    float u = 1.0f - t;
    float w0 = (-3.0f * u) * u;
    float w1 = 3.0f * ((((3.0f * t) * t) - (4.0f * t)) + 1.0f);
    float w2 = (3.0f * (2.0f - (3.0f * t))) * t;
    float w3 = (3.0f * t) * t;
    return ((w0 * c.p0 + w1 * c.p1) + w2 * c.p2) + w3 * c.p3;
}

// rmax of Prhi (hair_intersection.rint:21-22): chord distance of B(0.5) + radius.  A function of the
// curve alone, so the geometry kernel evaluates it once per curve instead of once per candidate.
VK_DEV float bezier_bound_radius(const Bezier& c, float radius)
{
    float3 m = bezier_point(c, 0.5f);
    float3 x = cross3(m - c.p0, m - c.p3);
    float3 ch = c.p3 - c.p0;
    float r = sqrtf(dot3(x, x)) / sqrtf(dot3(ch, ch));
    return r + radius;
}

// ---- shaders/cylinder.glsl:8-46 (boolean, no t>0 test, assumes |d| = 1) ------------------------
VK_DEV bool ray_hits_cylinder(float3 o, float3 d, float3 a, float3 b, float radius)
{
    float3 ba = b - a;
    float3 oc = o - a;
    float baba = dot3(ba, ba);
    float bard = dot3(ba, d);
    float baoc = dot3(ba, oc);
    float k2 = baba - bard * bard;
    float k1 = baba * dot3(oc, d) - baoc * bard;
    float k0 = (baba * dot3(oc, oc) - baoc * baoc) - (radius * radius) * baba;
    float h = k1 * k1 - k2 * k0;
    if (h < 0.0f) return false;
    h = sqrtf(h);
    float t = (-k1 - h) / k2;
    float y = baoc + t * bard;
    if (y > 0.0f && y < baba) return true;
    t = ((y < 0.0f ? 0.0f : baba) - baoc) / bard;
    return fabsf(k1 + k2 * t) < h;
}

// ---- ray-centric frame: shaders/ray.glsl:13-33 --------------------------------------------------
// Depends on the ray only, so it is built once per ray (the reference rebuilds it per candidate).
struct RayFrame { float3 e1, e2, e3; };

VK_DEV RayFrame make_ray_frame(float3 d)
{
    RayFrame f;
    f.e3 = normalize3(d);
    float3 w = f.e3;
    f.e2 = fabsf(w.x) > fabsf(w.y) ? normalize3(f3(-w.z, 0.0f, w.x)) : normalize3(f3(0.0f, w.z, -w.y));
    f.e1 = cross3(f.e2, w);
    return f;
}
// inverse of the rigid matrix [e1 e2 e3 o] applied to a point (curve.glsl:33-42 + ray.glsl:32)
VK_DEV float3 into_frame(const RayFrame& f, float3 o, float3 p)
{
    float3 q = p - o;
    return f3(dot3(f.e1, q), dot3(f.e2, q), dot3(f.e3, q));
}

// ---- shaders/cone.glsl:21-62, specialised to what Prhi reads (s, dt, real/phantom) --------------
struct ConeStep { float s, dt; bool real; };

VK_DEV ConeStep cone_step(float3 c, float radius, float3 ax, float slant)
{
    float r2 = radius * radius;
    float drr = radius * slant;
    float ddd = ax.x * ax.x + ax.y * ax.y;
    float dp = c.x * c.x + c.y * c.y;
    float cdd = c.x * ax.x + c.y * ax.y;
    float cxd = c.x * ax.y - c.y * ax.x;
    float qc = ddd;
    float qb = ax.z * (drr - cdd);
    float cdz2 = ax.z * ax.z;
    ddd += cdz2;
    float qa = (((2.0f * drr) * cdd + cxd * cxd) - ddd * r2) + dp * cdz2;
    float det = qb * qb - qa * qc;
    ConeStep r;
    r.real = det > 0.0f;
    r.s = (qb - (r.real ? sqrtf(det) : 0.0f)) / qc;
    r.dt = (r.s * ax.z - cdd) / ddd;
    return r;
}

// ---- Phantom Ray-Hair Intersector: hair_intersection.rint:35-130 (after the cylinder early-out) --
// Returns the reported distance (0 => nothing reported) and the converged curve parameter.
// `iters` counts cone evaluations (debug statistics only).
// The loop leaves a side early only when the march has reached an exact fp32 fixed point
// (t + dt == t on the plain-step branch): every later iteration would then recompute the very
// same state, so the result is unchanged (proof in DESIGN.md §4.3).
template <bool kCountIters>
VK_DEV float phantom_march(const RayFrame& fr, float3 o, const Bezier& world, float radius, float* u_out, uint32_t* iters)
{
    Bezier c;
    c.p0 = into_frame(fr, o, world.p0);
    c.p1 = into_frame(fr, o, world.p1);
    c.p2 = into_frame(fr, o, world.p2);
    c.p3 = into_frame(fr, o, world.p3);

    float3 chord = c.p3 - c.p0;
    float cz = chord.z * (1.0f / sqrtf(dot3(chord, chord)));   // z of normalize(chord); dot with (0,0,1)
    float t_start = cz > 0.0f ? 0.0f : 1.0f;
    float result = 0.0f;

#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        float t = t_start;
        float told = 0.0f, dt1 = 0.0f, dt2 = 0.0f;
#pragma unroll 1
        for (uint32_t i = 0; i < 8u; ++i) {
            if (kCountIters) (*iters)++;
            float3 centre = bezier_point(c, t);
            float3 axis = bezier_axis(c, t);
            ConeStep cs = cone_step(centre, radius, axis, 0.0f);
            if (cs.real && fabsf(cs.dt) < 5e-5f) {
                result = cs.s + centre.z;
                *u_out = t;
                break;
            }
            float dt = cs.dt;
            dt = 0.5f < dt ? 0.5f : dt;      // GLSL min(dt, 0.5)
            dt = dt < -0.5f ? -0.5f : dt;    // GLSL max(dt, -0.5)
            dt1 = dt2;
            dt2 = dt;
            float tn;
            bool plain = !(dt1 * dt2 < 0.0f);
            if (!plain) {
                tn = ((i & 3u) == 0u) ? 0.5f * (told + t) : (dt2 * told - dt1 * t) / (dt2 - dt1);
            } else {
                tn = t + dt;
            }
            told = t;
            bool fixed_point = plain && (tn == t);
            t = tn;
            if (t < 0.0f || t > 1.0f) break;
            if (fixed_point) break;
        }
        if (result > 0.0f) break;
        t_start = 1.0f - t_start;
    }
    return result;
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
    float dd = dot3(d, d);
    float t0 = (0.0f - dot3(d, oa0)) / dd;
    float3 oa = oa0 + t0 * d;
    float m0 = dot3(ba, ba), m1 = dot3(ba, oa), m2 = dot3(ba, d), m3 = dot3(d, oa), m5 = dot3(oa, oa);
    float rr = r0 - r1;
    float d2 = m0 - rr * rr;
    float bt = 0.0f, bu = 0.0f;
    bool found = false;
    if (d2 > 0.0f) {
        float a1 = m1 - r0 * rr;
        float k2 = d2 * dd - m2 * m2;
        float k1 = d2 * m3 - m2 * a1;
        float k0 = d2 * (m5 - r0 * r0) - a1 * a1;
        float h = k1 * k1 - k2 * k0;
        if (h >= 0.0f) {
            float t = (-k1 - sqrtf(h)) / k2;
            float y = a1 + t * m2;
            if (y > 0.0f && y < d2) { bt = t; bu = y / d2; found = true; }
        }
    }
    if (!found) {
        float h1 = m3 * m3 - dd * (m5 - r0 * r0);
        if (h1 > 0.0f) { bt = (-m3 - sqrtf(h1)) / dd; bu = 0.0f; found = true; }
        float3 ob = oa - ba;
        float m6 = dot3(d, ob), m7 = dot3(ob, ob);
        float h2 = m6 * m6 - dd * (m7 - r1 * r1);
        if (h2 > 0.0f) {
            float t = (-m6 - sqrtf(h2)) / dd;
            if (!found || t < bt) { bt = t; bu = 1.0f; found = true; }
        }
    }
    *t_out = bt + t0;
    *u_out = bu;
    if (kNormal) *n_out = normalize3((oa + bt * d) - bu * ba);
    return found;
}

// ---- DOTS: ray vs triangle, Moeller-Trumbore, no culling (defined by this project) ---------------
VK_DEV bool tri_intersect(float3 o, float3 d, float3 v0, float3 v1, float3 v2, uint32_t parity, float* t_out, float* u_out)
{
    float3 e1 = v1 - v0, e2 = v2 - v0;
    float3 p = cross3(d, e2);
    float det = dot3(e1, p);
    if (det == 0.0f || det != det) return false;
    float inv = 1.0f / det;
    float3 tv = o - v0;
    float b1 = dot3(tv, p) * inv;
    if (!(b1 >= 0.0f && b1 <= 1.0f)) return false;
    float3 q = cross3(tv, e1);
    float b2 = dot3(d, q) * inv;
    if (!(b2 >= 0.0f && b1 + b2 <= 1.0f)) return false;
    *t_out = dot3(e2, q) * inv;
    *u_out = parity ? b2 : (b1 + b2);
    return true;
}
VK_DEV float3 tri_normal(float3 d, float3 v0, float3 v1, float3 v2)
{
    float3 n = normalize3(cross3(v1 - v0, v2 - v0));
    if (dot3(n, d) > 0.0f) n = f3(-n.x, -n.y, -n.z);
    return n;
}

// ---- primary ray: shaders/ray_gen.rgen:16-24 ------------------------------------------------------
struct Camera { float vi[16]; float pi[16]; };   // CameraUniformData, column-major

VK_DEV void primary_ray(const Camera& cam, uint32_t W, uint32_t H, uint32_t px, uint32_t py, float sx, float sy,
                        float3* o, float3* d)
{
    float pcx = (float)px + sx, pcy = (float)py + sy;
    float u = pcx / (float)W, v = pcy / (float)H;
    float dx = u * 2.0f - 1.0f, dy = v * 2.0f - 1.0f;
    *o = f3(cam.vi[12], cam.vi[13], cam.vi[14]);
    float3 tg;
    tg.x = ((cam.pi[0] * dx + cam.pi[4] * dy) + cam.pi[8]) + cam.pi[12];
    tg.y = ((cam.pi[1] * dx + cam.pi[5] * dy) + cam.pi[9]) + cam.pi[13];
    tg.z = ((cam.pi[2] * dx + cam.pi[6] * dy) + cam.pi[10]) + cam.pi[14];
    float3 nd = normalize3(tg);
    d->x = (cam.vi[0] * nd.x + cam.vi[4] * nd.y) + cam.vi[8] * nd.z;
    d->y = (cam.vi[1] * nd.x + cam.vi[5] * nd.y) + cam.vi[9] * nd.z;
    d->z = (cam.vi[2] * nd.x + cam.vi[6] * nd.y) + cam.vi[10] * nd.z;
}

// ---- closest-hit colour: shaders/shading.glsl:1-11, debug.glsl:1-7 --------------------------------
VK_DEV float3 shade_normal(float3 n)
{
    float k = fabsf((n.x * 0.0f + n.y * -1.0f) + n.z * 0.0f);
    return f3(k * 0.4f + 0.3f, k * 0.2f + 0.3f, k * 0.1f + 0.3f);
}
VK_DEV float3 debug_palette(uint32_t prim)
{
    uint32_t i = prim % 6u;
    // (1,0,.3) (.8,.2,.3) (.6,.4,.3) (.4,.6,.3) (.2,.8,.3) (0,1,.3)
    const float r[6] = {1.0f, 0.8f, 0.6f, 0.4f, 0.2f, 0.0f};
    const float g[6] = {0.0f, 0.2f, 0.4f, 0.6f, 0.8f, 1.0f};
    return f3(r[i], g[i], 0.3f);
}
VK_DEV uint32_t to_unorm8(float c)
{
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    return (uint32_t)(c * 255.0f + 0.5f);
}

}  // namespace vkhrt
