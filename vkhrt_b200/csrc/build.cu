// build.cu — primitive generation + deterministic LBVH build, all on the device.
//
// Replaces (reference paths):
//   source/resources/model/geometry_processor.cpp:45-67   GenerateLines
//   ...:123-156 GenerateCurves, :422-436 GenerateAABBs, :201-271 DOTS, :273-298 LSS
//   source/bottom_level_acceleration_structure.cpp:34-78 (driver BLAS build) -> LBVH:
//     centroid bounds -> 63-bit Morton keys -> hand-written LSD radix sort (no CUB) ->
//     Karras hierarchy -> fused "materialise primitives in sorted order + bottom-up refit".
//
// Compiled with -fmad=false: generated control points / boxes are bit-identical to the CPU oracle.
#include "scene.h"
#include "hair_math.cuh"
#include <vector>

namespace vkhrt {

// ------------------------------------------------------------------------------------------------
// primitive generation (one thread = one primitive), shared by the centroid pass and the
// materialise pass so that nothing but the final sorted arrays is ever stored.
// ------------------------------------------------------------------------------------------------
struct MeshIn {
    const float* pos;
    const uint32_t* idx;
    const float* rpv;
    uint32_t n_segments;
    float radius;
    const float* curves;    // nullable: explicit curve list (12 floats each) left by MergeCurvesFast; replaces GenerateCurves
};

VK_DEV float3 load_pos(const MeshIn& m, uint32_t v) { return f3(m.pos[3 * v], m.pos[3 * v + 1], m.pos[3 * v + 2]); }
VK_DEV bool same_point(float3 a, float3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }   // glm::vec3 ==

struct Aabb { float3 lo, hi; };
// radii at the two ends of segment i: per-vertex when given (north_star: "polyline strands with per-vertex radius in"), else the scene's
VK_DEV void segment_radii(const MeshIn& m, uint32_t i, float* r0, float* r1)
{
    *r0 = m.rpv ? m.rpv[m.idx[2 * i]] : m.radius;
    *r1 = m.rpv ? m.rpv[m.idx[2 * i + 1]] : m.radius;
}

// GenerateCurves (geometry_processor.cpp:123-156), tension 1, for segment i
VK_DEV Bezier gen_curve(const MeshIn& m, uint32_t i)
{
    if (m.curves) {
        const float* c = m.curves + 12 * (size_t)i;
        Bezier b;
        b.p0 = f3(c[0], c[1], c[2]); b.p1 = f3(c[3], c[4], c[5]); b.p2 = f3(c[6], c[7], c[8]); b.p3 = f3(c[9], c[10], c[11]);
        return b;
    }
    // All index pairs first, then all six vertices, then the selection: two dependent round trips to memory instead of four (the
    // neighbours' far vertices are fetched whether or not the strand continues; they are the next / previous thread's own vertices).
    const uint2* pairs = reinterpret_cast<const uint2*>(m.idx);
    const bool has_prev = i > 0, has_next = i + 1 < m.n_segments;
    const uint2 ic = __ldg(pairs + i), ip = __ldg(pairs + (has_prev ? i - 1 : i)), in = __ldg(pairs + (has_next ? i + 1 : i));
    const float3 s = load_pos(m, ic.x), e = load_pos(m, ic.y);
    const float3 pe = load_pos(m, ip.y), pp = load_pos(m, ip.x), ns = load_pos(m, in.x), nn = load_pos(m, in.y);
    const float3 p0 = (has_prev && same_point(pe, s)) ? pp : s;
    const float3 p3 = (has_next && same_point(e, ns)) ? nn : e;
    const float k = 1.0f / 6.0f;
    Bezier c;
    c.p0 = s;
    c.p1 = s + (e - p0) * k;
    c.p2 = e - (p3 - s) * k;
    c.p3 = e;
    return c;
}
// GenerateAABBs (geometry_processor.cpp:422-436)
VK_DEV Aabb curve_box(const Bezier& c, float r)
{
    Aabb b;
    b.lo.x = fminf(fminf(c.p0.x, c.p3.x), fminf(c.p1.x, c.p2.x)) - r;
    b.lo.y = fminf(fminf(c.p0.y, c.p3.y), fminf(c.p1.y, c.p2.y)) - r;
    b.lo.z = fminf(fminf(c.p0.z, c.p3.z), fminf(c.p1.z, c.p2.z)) - r;
    b.hi.x = fmaxf(fmaxf(c.p0.x, c.p3.x), fmaxf(c.p1.x, c.p2.x)) + r;
    b.hi.y = fmaxf(fmaxf(c.p0.y, c.p3.y), fmaxf(c.p1.y, c.p2.y)) + r;
    b.hi.z = fmaxf(fmaxf(c.p0.z, c.p3.z), fmaxf(c.p1.z, c.p2.z)) + r;
    return b;
}

// GenerateLinearSweptSpheres (geometry_processor.cpp:273-298) + per-vertex radius extension
struct LssPrim { float3 p0, p1; float r0, r1; };
VK_DEV LssPrim gen_lss(const MeshIn& m, uint32_t i)
{
    uint32_t a = m.idx[2 * i], b = m.idx[2 * i + 1];
    LssPrim s;
    s.p0 = load_pos(m, a);
    s.p1 = load_pos(m, b);
    s.r0 = fmaxf(m.rpv ? m.rpv[a] : m.radius, 0.001f);
    s.r1 = fmaxf(m.rpv ? m.rpv[b] : m.radius, 0.001f);
    return s;
}
VK_DEV Aabb lss_box(const LssPrim& s)
{
    Aabb b;
    b.lo = f3(fminf(s.p0.x - s.r0, s.p1.x - s.r1), fminf(s.p0.y - s.r0, s.p1.y - s.r1), fminf(s.p0.z - s.r0, s.p1.z - s.r1));
    b.hi = f3(fmaxf(s.p0.x + s.r0, s.p1.x + s.r1), fmaxf(s.p0.y + s.r0, s.p1.y + s.r1), fmaxf(s.p0.z + s.r0, s.p1.z + s.r1));
    return b;
}

// GenerateDisjointOrthogonalTriangleStrips (geometry_processor.cpp:201-271): triangle `prim` = 4*segment + 2*face + k
struct TriPrim { float3 v0, v1, v2; };
VK_DEV float3 perp_stark(float3 u)
{
    float ax = fabsf(u.x), ay = fabsf(u.y), az = fabsf(u.z);
    uint32_t uyx = (ax - ay) < 0.0f ? 1u : 0u;
    uint32_t uzx = (ax - az) < 0.0f ? 1u : 0u;
    uint32_t uzy = (ay - az) < 0.0f ? 1u : 0u;
    uint32_t xm = uyx & uzx;
    uint32_t ym = (1u ^ xm) & uzy;
    uint32_t zm = 1u ^ (xm | ym);
    return normalize3(cross3(u, f3((float)xm, (float)ym, (float)zm)));
}
VK_DEV TriPrim gen_tri(const MeshIn& m, uint32_t prim)
{
    uint32_t seg = prim >> 2, face = (prim >> 1) & 1u, k = prim & 1u;
    float3 s = load_pos(m, m.idx[2 * seg]), e = load_pos(m, m.idx[2 * seg + 1]);
    float3 fwd = normalize3(e - s);
    float3 sv = perp_stark(fwd);
    float3 v = face ? cross3(fwd, sv) : sv;
    float r0, r1;
    segment_radii(m, seg, &r0, &r1);
    float3 offs = v * r0, offe = v * r1;
    TriPrim t;
    if (k == 0) { t.v0 = s + offs; t.v1 = e - offe; t.v2 = e + offe; }
    else        { t.v0 = s + offs; t.v1 = s - offs; t.v2 = e - offe; }
    return t;
}
VK_DEV Aabb tri_box(const TriPrim& t)
{
    Aabb b;
    b.lo = f3(fminf(fminf(t.v0.x, t.v1.x), t.v2.x), fminf(fminf(t.v0.y, t.v1.y), t.v2.y), fminf(fminf(t.v0.z, t.v1.z), t.v2.z));
    b.hi = f3(fmaxf(fmaxf(t.v0.x, t.v1.x), t.v2.x), fmaxf(fmaxf(t.v0.y, t.v1.y), t.v2.y), fmaxf(fmaxf(t.v0.z, t.v1.z), t.v2.z));
    return b;
}

// ---- leaf pieces (VKHRT_LEAF_SPLIT_*, include/vkhrt_b200.h): leaf = group * K + piece ----------------------------------
// Rounding guard of a piece box: the piece's corner points are recomputed in fp32 (midpoints, s + (e-s)*a), so the box is
// padded by a few ulps of its largest coordinate.  The same statement order is in the oracle (piece_guard).
VK_DEV float piece_guard(const Aabb& b)
{
    float m = fmaxf(fmaxf(fmaxf(fabsf(b.lo.x), fabsf(b.lo.y)), fabsf(b.lo.z)), fmaxf(fmaxf(fabsf(b.hi.x), fabsf(b.hi.y)), fabsf(b.hi.z)));
    return m * 8e-7f + 1e-7f;
}
VK_DEV void grow_box(Aabb& b, float3 p)
{
    b.lo = f3(fminf(b.lo.x, p.x), fminf(b.lo.y, p.y), fminf(b.lo.z, p.z));
    b.hi = f3(fmaxf(b.hi.x, p.x), fmaxf(b.hi.y, p.y), fmaxf(b.hi.z, p.z));
}
VK_DEV Aabb pad_box(Aabb b, float pad)
{
    b.lo = f3(b.lo.x - pad, b.lo.y - pad, b.lo.z - pad);
    b.hi = f3(b.hi.x + pad, b.hi.y + pad, b.hi.z + pad);
    return b;
}
// piece `piece` of K = 2^levels equal parameter ranges of a cubic Bezier: repeated de Casteljau halving, control hull + radius
template <int K>
VK_DEV Aabb curve_piece_box(Bezier c, uint32_t piece, float r)
{
    if (K == 1) return curve_box(c, r);
#pragma unroll
    for (int half = K >> 1; half >= 1; half >>= 1) {
        float3 q01 = (c.p0 + c.p1) * 0.5f, q12 = (c.p1 + c.p2) * 0.5f, q23 = (c.p2 + c.p3) * 0.5f;
        float3 r0 = (q01 + q12) * 0.5f, r1 = (q12 + q23) * 0.5f;
        float3 mid = (r0 + r1) * 0.5f;
        if (piece & (uint32_t)half) { c.p0 = mid; c.p1 = r1; c.p2 = q23; }
        else { c.p1 = q01; c.p2 = r0; c.p3 = mid; }
    }
    Aabb b; b.lo = c.p0; b.hi = c.p0;
    grow_box(b, c.p1); grow_box(b, c.p2); grow_box(b, c.p3);
    return pad_box(b, r + piece_guard(b));
}
// piece `piece` of K equal parts of a DOTS strip: the two crossed quads between s + (e-s)*a and s + (e-s)*b
template <int K>
VK_DEV Aabb strip_piece_box(const MeshIn& m, uint32_t seg, uint32_t piece)
{
    float3 s = load_pos(m, m.idx[2 * seg]), e = load_pos(m, m.idx[2 * seg + 1]);
    float3 fwd = normalize3(e - s);
    float3 sv = perp_stark(fwd), tv = cross3(fwd, sv);
    float a = (float)piece / (float)K, bb = (float)(piece + 1u) / (float)K;
    float r0, r1;
    segment_radii(m, seg, &r0, &r1);
    float dr = r1 - r0;
    float ra = r0 + dr * a, rb = r0 + dr * bb;       // the strip's edges are straight: its half-width is linear along the segment
    float3 se = e - s;
    float3 A = s + se * a, B = s + se * bb;
    float3 a0 = sv * ra, a1 = tv * ra, b0 = sv * rb, b1 = tv * rb;
    Aabb b; b.lo = A + a0; b.hi = b.lo;
    grow_box(b, A - a0); grow_box(b, A + a1); grow_box(b, A - a1);
    grow_box(b, B + b0); grow_box(b, B - b0); grow_box(b, B + b1); grow_box(b, B - b1);
    return pad_box(b, piece_guard(b));
}
// piece `piece` of K equal parts of a linear swept sphere: the two end spheres of the part (radius is linear along the segment)
template <int K>
VK_DEV Aabb lss_piece_box(const LssPrim& s, uint32_t piece)
{
    if (K == 1) return lss_box(s);
    float a = (float)piece / (float)K, bb = (float)(piece + 1u) / (float)K;
    float3 se = s.p1 - s.p0;
    float dr = s.r1 - s.r0;
    LssPrim q;
    q.p0 = s.p0 + se * a; q.p1 = s.p0 + se * bb;
    q.r0 = s.r0 + dr * a; q.r1 = s.r0 + dr * bb;
    Aabb b = lss_box(q);
    return pad_box(b, piece_guard(b));
}
template <int TECH> struct LeafSplit { static constexpr uint32_t K = TECH == VKHRT_TECHNIQUE_PHANTOM ? VKHRT_LEAF_SPLIT_PHANTOM : (TECH == VKHRT_TECHNIQUE_LSS ? VKHRT_LEAF_SPLIT_LSS : VKHRT_LEAF_SPLIT_DOTS); };

// box of BVH leaf `leaf` = piece (leaf % K) of group (leaf / K)
template <int TECH>
VK_DEV Aabb leaf_box(const MeshIn& m, uint32_t leaf)
{
    constexpr uint32_t K = LeafSplit<TECH>::K;
    const uint32_t group = leaf / K, piece = leaf % K;
    if (TECH == VKHRT_TECHNIQUE_PHANTOM) {
        float r0, r1;
        segment_radii(m, group, &r0, &r1);
        return curve_piece_box<(int)K>(gen_curve(m, group), piece, fmaxf(r0, r1));       // the curve's larger end radius
    }
    if (TECH == VKHRT_TECHNIQUE_LSS) return lss_piece_box<(int)K>(gen_lss(m, group), piece);
    if (K > 1) return strip_piece_box<(int)K>(m, group, piece);
    Aabb b = tri_box(gen_tri(m, 4u * group));
#pragma unroll
    for (uint32_t k = 1; k < 4; ++k) {
        Aabb c = tri_box(gen_tri(m, 4u * group + k));
        b.lo = f3(fminf(b.lo.x, c.lo.x), fminf(b.lo.y, c.lo.y), fminf(b.lo.z, c.lo.z));
        b.hi = f3(fmaxf(b.hi.x, c.hi.x), fmaxf(b.hi.y, c.hi.y), fmaxf(b.hi.z, c.hi.z));
    }
    return b;
}

// order-preserving float <-> uint for atomicMin/Max
VK_DEV uint32_t f2ord(float f) { uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
static inline float ord2f_host(uint32_t o) { uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o; float f; memcpy(&f, &b, 4); return f; }

// pass 1a: centroids + their bounds
template <int TECH>
__global__ void __launch_bounds__(256) centroid_kernel(MeshIn m, uint32_t n_prims, float4* __restrict__ centroids, uint32_t* __restrict__ bounds /*6*/)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const float inf = __int_as_float(0x7f800000);
    float3 lo = f3(inf, inf, inf), hi = f3(-inf, -inf, -inf);
    if (i < n_prims) {
        Aabb b = leaf_box<TECH>(m, i);
        float3 c = (b.lo + b.hi) * 0.5f;
        centroids[i] = make_float4(c.x, c.y, c.z, 0.0f);
        lo = c; hi = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o)); hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    __shared__ float red[8][6];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[w][0] = lo.x; red[w][1] = lo.y; red[w][2] = lo.z; red[w][3] = hi.x; red[w][4] = hi.y; red[w][5] = hi.z; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[0][threadIdx.x];
        for (int k = 1; k < 8; ++k) v = threadIdx.x < 3 ? fminf(v, red[k][threadIdx.x]) : fmaxf(v, red[k][threadIdx.x]);
        if (threadIdx.x < 3) atomicMin(&bounds[threadIdx.x], f2ord(v)); else atomicMax(&bounds[threadIdx.x], f2ord(v));
    }
}

// pass 1b: 63-bit Morton keys (21 bits per axis)
VK_DEV uint64_t spread21(uint32_t v)
{
    uint64_t x = v & 0x1FFFFFull;
    x = (x | (x << 32)) & 0x1F00000000FFFFull;
    x = (x | (x << 16)) & 0x1F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
VK_DEV uint32_t quantise21(float c, float lo, float scale)
{
    float q = (c - lo) * scale;
    q = fminf(fmaxf(q, 0.0f), 2097151.0f);
    return (uint32_t)q;
}
VK_DEV float ord2f(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }
// the centroid bounds stay on the device (ordered-uint min/max of centroid_kernel): no host round trip in the middle of the build
__global__ void __launch_bounds__(256) morton_kernel(const float4* __restrict__ centroids, uint32_t n, const uint32_t* __restrict__ bounds,
                                                     uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 lo = f3(ord2f(bounds[0]), ord2f(bounds[1]), ord2f(bounds[2]));
    const float3 ext = f3(ord2f(bounds[3]) - lo.x, ord2f(bounds[4]) - lo.y, ord2f(bounds[5]) - lo.z);
    const float3 scale = f3(ext.x > 0.0f ? 2097152.0f / ext.x : 0.0f, ext.y > 0.0f ? 2097152.0f / ext.y : 0.0f, ext.z > 0.0f ? 2097152.0f / ext.z : 0.0f);
    float4 c = centroids[i];
    uint64_t mx = spread21(quantise21(c.x, lo.x, scale.x));
    uint64_t my = spread21(quantise21(c.y, lo.y, scale.y));
    uint64_t mz = spread21(quantise21(c.z, lo.z, scale.z));
    keys[i] = (mx << 2) | (my << 1) | mz;
    vals[i] = i;
}

// ------------------------------------------------------------------------------------------------
// Radix sort of (64-bit Morton key, 32-bit leaf id): LSD, 8-bit digits, stable, ONE read and ONE write of the data per
// digit pass ("onesweep": Adinets & Merrill 2022; hand-written, no CUB / Thrust).
//   os_histogram_kernel  one pass over the keys: the 256-bin histograms of all 8 digits at once
//   os_scan_kernel       exclusive scan of each of the 8 histograms -> where every digit's run starts in the output
//   os_pass_kernel       per digit: a CTA takes the next tile of 3072 keys (ticket counter: tiles start in order), ranks them
//                        stably (warp-match ranking inside 256-key warp chunks, chunk prefixes per digit), learns where its
//                        keys of each digit go by DECOUPLED LOOK-BACK over the previous tiles' per-digit counts (one 64-bit
//                        status word per tile and digit = flag + count, so no fence is needed), regroups the tile by digit in
//                        shared memory and writes each digit's run with consecutive threads.
// The status words carry the pass number in their flag (2 pass + 1 = tile count, 2 pass + 2 = inclusive prefix), so one
// buffer, zeroed once per sort, serves all passes.  8 x (12 B read + 12 B write) + 8 B per key instead of the 8 x (20 + 12) + 8
// histogram/scan/scatter passes of round 1 (C2: 6.4 M keys in 1.21 ms = 5.3 Gkeys/s before).
// ------------------------------------------------------------------------------------------------
constexpr int OS_BLOCK = 384;
constexpr int OS_IPT = 8;                       // keys per thread
constexpr int OS_TILE = OS_BLOCK * OS_IPT;      // 3072 keys per tile (44 KB of static shared memory)
constexpr int OS_WARPS = OS_BLOCK / 32;
constexpr int OS_PASSES = 8;

__global__ void __launch_bounds__(512) os_histogram_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ hist /* [8][256] */)
{
    __shared__ uint32_t sh[OS_PASSES][256];
    for (int k = threadIdx.x; k < OS_PASSES * 256; k += blockDim.x) (&sh[0][0])[k] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
#pragma unroll
        for (int p = 0; p < OS_PASSES; ++p) atomicAdd(&sh[p][(uint32_t)(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < OS_PASSES * 256; k += blockDim.x) { const uint32_t v = (&sh[0][0])[k]; if (v) atomicAdd(hist + k, v); }
}
// in place: hist[p][d] -> number of keys whose digit p is smaller than d
__global__ void __launch_bounds__(256) os_scan_kernel(uint32_t* __restrict__ hist)
{
    __shared__ uint32_t wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int p = 0; p < OS_PASSES; ++p) {
        const uint32_t v = hist[p * 256 + threadIdx.x];
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        uint32_t base = 0;
        for (int k = 0; k < w; ++k) base += wsum[k];
        hist[p * 256 + threadIdx.x] = base + inc - v;
        __syncthreads();
    }
}
VK_DEV unsigned long long os_ld(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
VK_DEV void os_st(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

#ifndef VKHRT_OS_MIN_BLOCKS
#define VKHRT_OS_MIN_BLOCKS 4       // CTAs per SM the register allocation is held to: 2 (80 registers) / 3 (56) / 4 (40, 36 B spilled) sort C2 in 0.84 / 0.78 / 0.73 ms
#endif
__global__ void __launch_bounds__(OS_BLOCK, VKHRT_OS_MIN_BLOCKS) os_pass_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                           uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int pass,
                                                           const uint32_t* __restrict__ digit_start /* [256] of this pass */,
                                                           unsigned long long* __restrict__ status /* [tiles][256] */, uint32_t* __restrict__ ticket /* [8] */)
{
    __shared__ uint16_t warp_cnt[OS_WARPS][256];   // keys of digit d in the chunks of the warps before w (after the prefix step)
    __shared__ uint32_t dig_first[256];            // first position of digit d in the regrouped tile
    __shared__ uint32_t dig_dst[256];              // global position of this tile's first key of digit d, minus dig_first[d]
    __shared__ uint32_t scan_w[8];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_key[OS_TILE];
    __shared__ uint32_t s_val[OS_TILE];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int shift = 8 * pass;
    if (tid == 0) s_tile = atomicAdd(ticket + pass, 1u);
    for (int k = tid; k < OS_WARPS * 256; k += OS_BLOCK) (&warp_cnt[0][0])[k] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    // warp w owns the contiguous chunk [tile * OS_TILE + w * 256, + 256); round r = 32 consecutive keys: ranks are stable
    const uint32_t wbase = tile * OS_TILE + w * (32 * OS_IPT);
    uint64_t key[OS_IPT];
    uint32_t val[OS_IPT];
    uint16_t rank[OS_IPT];
#pragma unroll
    for (int r = 0; r < OS_IPT; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys_in[i] : ~0ull;
        val[r] = ok ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < OS_IPT; ++r) {
        const bool ok = wbase + r * 32 + lane < n;
        const uint32_t d = ok ? ((uint32_t)(key[r] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t pre = ok ? warp_cnt[w][d] : 0u;
        __syncwarp();
        if (ok && (peers & ((1u << lane) - 1u)) == 0u) warp_cnt[w][d] = (uint16_t)(pre + __popc(peers));
        __syncwarp();
        rank[r] = (uint16_t)(pre + __popc(peers & ((1u << lane) - 1u)));
    }
    __syncthreads();
    // digit d = tid (threads 0..255): prefix over the warps' chunks, tile count, look-back
    uint32_t count = 0;
    if (tid < 256) {
#pragma unroll
        for (int k = 0; k < OS_WARPS; ++k) { const uint32_t c = warp_cnt[k][tid]; warp_cnt[k][tid] = (uint16_t)count; count += c; }
        const unsigned long long AGG = (unsigned long long)(2 * pass + 1) << 56, INC = (unsigned long long)(2 * pass + 2) << 56;
        unsigned long long* mine = status + (size_t)tile * 256 + tid;
        if (tile > 0) os_st(mine, AGG | count);
        uint32_t excl = 0;
        for (int t = (int)tile - 1; t >= 0; --t) {
            unsigned long long sw;
            do { sw = os_ld(status + (size_t)t * 256 + tid); } while ((sw >> 56) != (AGG >> 56) && (sw >> 56) != (INC >> 56));
            excl += (uint32_t)sw;
            if ((sw >> 56) == (INC >> 56)) break;
        }
        os_st(mine, INC | (unsigned long long)(excl + count));
        // exclusive scan of the tile's digit counts -> where each digit starts in the regrouped tile
        uint32_t inc = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) scan_w[w] = inc;
        dig_first[tid] = inc - count;                  // within the warp; the warp bases are added below
        dig_dst[tid] = digit_start[tid] + excl;
    }
    __syncthreads();
    if (tid < 256) {
        uint32_t base = 0;
        for (int k = 0; k < w; ++k) base += scan_w[k];
        const uint32_t first = dig_first[tid] + base;
        dig_first[tid] = first;
        dig_dst[tid] -= first;
    }
    __syncthreads();
    // regroup the tile by digit in shared memory ...
#pragma unroll
    for (int r = 0; r < OS_IPT; ++r) {
        if (wbase + r * 32 + lane < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
            const uint32_t at = dig_first[d] + warp_cnt[w][d] + rank[r];
            s_key[at] = key[r]; s_val[at] = val[r];
        }
    }
    __syncthreads();
    // ... and write every digit's run with consecutive threads
    const uint32_t in_tile = min((uint32_t)OS_TILE, n - tile * OS_TILE);
#pragma unroll
    for (int r = 0; r < OS_IPT; ++r) {
        const uint32_t at = r * OS_BLOCK + tid;
        if (at < in_tile) {
            const uint64_t k = s_key[at];
            const uint32_t o = dig_dst[(uint32_t)(k >> shift) & 255u] + at;
            keys_out[o] = k; vals_out[o] = s_val[at];
        }
    }
}

// exclusive scan of uint32, in place: 2048 elements per block, block totals scanned recursively
constexpr int SC_BLOCK = 256, SC_IPT = 8, SC_TILE = SC_BLOCK * SC_IPT;
__global__ void __launch_bounds__(SC_BLOCK) scan_tile_kernel(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ totals)
{
    __shared__ uint32_t wsum[SC_BLOCK / 32];
    uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_IPT;
    uint32_t v[SC_IPT], s = 0;
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) { v[k] = (base + k < n) ? data[base + k] : 0u; s += v[k]; }
    uint32_t inc = s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < SC_BLOCK / 32 ? wsum[lane] : 0u, xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
        if (lane < SC_BLOCK / 32) wsum[lane] = xi - x;
        if (lane == SC_BLOCK / 32 - 1 && totals) totals[blockIdx.x] = xi;
    }
    __syncthreads();
    uint32_t run = wsum[w] + (inc - s);
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) { if (base + k < n) data[base + k] = run; run += v[k]; }
}
__global__ void __launch_bounds__(SC_BLOCK) scan_add_kernel(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ offsets)
{
    uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_IPT;
    uint32_t off = offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SC_IPT; ++k) if (base + k < n) data[base + k] += off;
}

static int exclusive_scan(uint32_t* d, uint32_t n, uint32_t* tmp, cudaStream_t st)
{
    uint32_t nb = (n + SC_TILE - 1) / SC_TILE;
    if (nb <= 1) {
        scan_tile_kernel<<<1, SC_BLOCK, 0, st>>>(d, n, nullptr);
        count_launch();
        return VKHRT_OK;
    }
    scan_tile_kernel<<<nb, SC_BLOCK, 0, st>>>(d, n, tmp);
    count_launch();
    int rc = exclusive_scan(tmp, nb, tmp + ((nb + 63) & ~63u), st);
    if (rc) return rc;
    scan_add_kernel<<<nb, SC_BLOCK, 0, st>>>(d, n, tmp);
    count_launch();
    return VKHRT_OK;
}
static size_t scan_tmp_elems(uint32_t n)
{
    size_t tot = 0;
    while (n > SC_TILE) { n = (n + SC_TILE - 1) / SC_TILE; tot += (n + 63) & ~63u; }
    return tot + 64;
}

// ------------------------------------------------------------------------------------------------
// Karras 2012 hierarchy over the sorted keys. Internal node i in [0, n-2], root = 0.
// delta(i,j) = common-prefix length of the 64-bit keys, ties broken by the sorted position.
// ------------------------------------------------------------------------------------------------
// `a` = m[i], loaded once by the caller (every search of node i compares against the same key)
VK_DEV int prefix_len(const uint64_t* __restrict__ m, int n, int i, uint64_t a, int j)
{
    if ((unsigned)j >= (unsigned)n) return -1;
    const uint64_t x = a ^ __ldg(m + j);
    if (x == 0ull) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)x);
}

// The bottom-up box pass works on tiles of RefitTile<TECH>::N consecutive (Morton-sorted) leaves, one CTA each: a node whose leaf range lies
// inside one tile is only ever reached by threads of that CTA (node_local = 1) and is handled in shared memory.
// Tile size per technique (measured on the C2 groom, refit ms at 256 / 512 leaves per tile: Phantom 0.545 / 0.566 — its record arithmetic is long,
// more and smaller CTAs overlap their phases better; LSS 0.358 / 0.354, DOTS 0.768 / 0.743 — short records, the deeper local walk wins).
template <int TECH> struct RefitTile { static constexpr int SHIFT = TECH == VKHRT_TECHNIQUE_PHANTOM ? 8 : 9, N = 1 << SHIFT; };
static int refit_tile_shift(int tech) { return tech == VKHRT_TECHNIQUE_PHANTOM ? RefitTile<VKHRT_TECHNIQUE_PHANTOM>::SHIFT : RefitTile<VKHRT_TECHNIQUE_LSS>::SHIFT; }
__global__ void __launch_bounds__(256) karras_kernel(const uint64_t* __restrict__ morton, const uint32_t* __restrict__ sorted_ids, int n,
                                                     uint32_t* __restrict__ nodes_u32 /* 16 words per node */,
                                                     uint32_t* __restrict__ parent_internal, uint32_t* __restrict__ parent_leaf, uint8_t* __restrict__ node_local,
                                                     uint32_t* __restrict__ child0, int tile_shift)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const uint64_t key_i = __ldg(morton + i);
    int d = (prefix_len(morton, n, i, key_i, i + 1) - prefix_len(morton, n, i, key_i, i - 1)) >= 0 ? 1 : -1;
    int dmin = prefix_len(morton, n, i, key_i, i - d);
    int lmax = 2;
    while (prefix_len(morton, n, i, key_i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (prefix_len(morton, n, i, key_i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = prefix_len(morton, n, i, key_i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (prefix_len(morton, n, i, key_i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    const bool local = (lo >> tile_shift) == (hi >> tile_shift);
    uint32_t c0, c1, p0 = 0, p1 = 0;
    if (lo == gamma) { c0 = VKHRT_BVH_LEAF | (uint32_t)gamma; p0 = sorted_ids[gamma]; parent_leaf[gamma] = ((uint32_t)i << 1); }
    else { c0 = (uint32_t)gamma; parent_internal[gamma] = ((uint32_t)i << 1); }
    if (hi == gamma + 1) { c1 = VKHRT_BVH_LEAF | (uint32_t)(gamma + 1); p1 = sorted_ids[gamma + 1]; parent_leaf[gamma + 1] = ((uint32_t)i << 1) | 1u; }
    else { c1 = (uint32_t)(gamma + 1); parent_internal[gamma + 1] = ((uint32_t)i << 1) | 1u; }
    // The four non-box words of a node follow from child0 and one bit (child1 = child0's position + 1; prim words = sorted_ids there).
    // Nodes inside one refit tile get their WHOLE 64-byte record from materialise_refit_kernel (full 16-byte stores: scattering these four
    // words here cost a read-modify-write of every 32-byte sector of the node array); only the few nodes shared between tiles are written here.
    child0[i] = c0;
    node_local[i] = (uint8_t)((local ? 1u : 0u) | ((c1 & VKHRT_BVH_LEAF) ? 2u : 0u));
    if (!local) {
        uint32_t* nd = nodes_u32 + (size_t)i * 16;
        nd[3] = c0; nd[7] = c1; nd[11] = p0; nd[15] = p1;
    }
    if (i == 0) parent_internal[0] = 0xFFFFFFFFu;
}

// ------------------------------------------------------------------------------------------------
// Materialise primitives at their sorted position and refit boxes bottom-up (one thread per leaf;
// the second thread to arrive at a node carries the union upward).  min/max are exact, so the
// boxes do not depend on arrival order.
// ------------------------------------------------------------------------------------------------
template <int TECH>
__global__ void __launch_bounds__(RefitTile<TECH>::N, (TECH == VKHRT_TECHNIQUE_PHANTOM ? 1024 : 2048) / RefitTile<TECH>::N) materialise_refit_kernel(MeshIn m, uint32_t n_prims, const uint32_t* __restrict__ sorted_ids,
                                                                float4* __restrict__ primA, float4* __restrict__ primB, float2* __restrict__ primR,
                                                                float* nodes_f32, const uint32_t* __restrict__ parent_internal,
                                                                const uint32_t* __restrict__ parent_leaf, uint32_t* flags,
                                                                const uint8_t* __restrict__ node_local, const uint32_t* __restrict__ child0,
                                                                float4* __restrict__ exits /* 2 per entry */, uint32_t* __restrict__ exit_count, uint32_t exit_cap)
{
    constexpr int REFIT_TILE = RefitTile<TECH>::N;
    // Everything the walk needs for the nodes INSIDE this CTA's tile of leaves lives in shared memory (index = node - tile base):
    // the parent link and the "local" bit (loaded coalesced), an arrival counter, and both child boxes, which are written to the
    // node records by the whole CTA at the end (12 of every 16 words, consecutive threads -> consecutive words).
    __shared__ float s_box[REFIT_TILE][12];        // lo0 hi0 lo1 hi1 (the box words of VkhrtBvhNode, without the child / prim words)
    __shared__ uint32_t s_arrived[REFIT_TILE];
    __shared__ uint32_t s_parent[REFIT_TILE];
    __shared__ uint32_t s_child0[REFIT_TILE];      // child 0 of the node; child 1 sits one position further (bit 1 of s_local: it is a leaf)
    __shared__ uint32_t s_leaf[REFIT_TILE];        // sorted_ids of the tile's leaves (the prim words of the node records)
    __shared__ uint8_t s_local[REFIT_TILE];        // bit 0: the node's leaf range lies inside this tile
    const uint32_t tile_base = blockIdx.x * REFIT_TILE;
    uint32_t p = 0;
    {
        const uint32_t node = tile_base + threadIdx.x;
        const bool in = n_prims > 1 && node < n_prims - 1;
        s_arrived[threadIdx.x] = 0u;
        s_parent[threadIdx.x] = in ? parent_internal[node] : 0u;
        s_child0[threadIdx.x] = in ? child0[node] : 0u;
        s_local[threadIdx.x] = in ? node_local[node] : (uint8_t)0;
        s_leaf[threadIdx.x] = node < n_prims ? sorted_ids[node] : 0u;
        if (node < n_prims && n_prims > 1) p = parent_leaf[node];          // needed after the record's arithmetic: fetched with the rest
    }
    __syncthreads();
    uint32_t pos = tile_base + threadIdx.x;
    bool walking = pos < n_prims;
    Aabb box;
    box.lo = box.hi = f3(0, 0, 0);
    if (walking) {
    // the leaf's record is its GROUP's (a curve / LSS / strip reached through any of its pieces is tested whole); its box is the piece's
    const uint32_t leaf = s_leaf[threadIdx.x];
    const uint32_t prim = leaf / LeafSplit<TECH>::K;
    if (TECH == VKHRT_TECHNIQUE_PHANTOM) {
        Bezier c = gen_curve(m, prim);
        float r0, r1;
        segment_radii(m, prim, &r0, &r1);
        const float rbig = fmaxf(r0, r1);
        box = curve_piece_box<(int)LeafSplit<TECH>::K>(c, leaf % LeafSplit<TECH>::K, rbig);
        float rmax = bezier_bound_radius(c, rbig);
        primA[2 * (size_t)pos] = make_float4(c.p0.x, c.p0.y, c.p0.z, rmax);
        primA[2 * (size_t)pos + 1] = make_float4(c.p3.x, c.p3.y, c.p3.z, __uint_as_float(prim));
        primB[2 * (size_t)pos] = make_float4(c.p1.x, c.p1.y, c.p1.z, bezier_quarter_chord_deviation(c) + bezier_convergence_slack(c, fminf(r0, r1)));
        primB[2 * (size_t)pos + 1] = make_float4(c.p2.x, c.p2.y, c.p2.z, 0.0f);
        if (primR) primR[pos] = make_float2(r0, r1);
    } else if (TECH == VKHRT_TECHNIQUE_LSS) {
        LssPrim s = gen_lss(m, prim);
        box = lss_piece_box<(int)LeafSplit<TECH>::K>(s, leaf % LeafSplit<TECH>::K);
        primA[2 * (size_t)pos] = make_float4(s.p0.x, s.p0.y, s.p0.z, s.r0);
        primA[2 * (size_t)pos + 1] = make_float4(s.p1.x, s.p1.y, s.p1.z, s.r1);
    } else {
        // strip record: the traversal kernel rebuilds the 12 vertices as start/end -+ offset with the very same
        // single fp32 add/sub gen_tri() performs, so they are bit-identical to the generator's
        box = leaf_box<TECH>(m, leaf);
        float3 s = load_pos(m, m.idx[2 * prim]), e = load_pos(m, m.idx[2 * prim + 1]);
        float3 fwd = normalize3(e - s);
        float3 sv = perp_stark(fwd);
        float3 tv = cross3(fwd, sv);
        primA[4 * (size_t)pos] = make_float4(s.x, s.y, s.z, __uint_as_float(prim));
        primA[4 * (size_t)pos + 1] = make_float4(e.x, e.y, e.z, 0.0f);
        if (m.rpv) {
            // per-vertex radius: the record keeps the unit frame vectors and the two end radii; the kernel forms the offsets
            // v * r0 / v * r1 with the one multiply gen_tri() performs
            float r0, r1;
            segment_radii(m, prim, &r0, &r1);
            primA[4 * (size_t)pos + 2] = make_float4(sv.x, sv.y, sv.z, r0);
            primA[4 * (size_t)pos + 3] = make_float4(tv.x, tv.y, tv.z, r1);
        } else {
            float3 off0 = sv * m.radius, off1 = tv * m.radius;
            primA[4 * (size_t)pos + 2] = make_float4(off0.x, off0.y, off0.z, 0.0f);
            primA[4 * (size_t)pos + 3] = make_float4(off1.x, off1.y, off1.z, 0.0f);
        }
    }
    if (n_prims == 1) {   // single primitive: node 0 holds the same leaf in both slots
        float* nd = nodes_f32;
        for (int k = 0; k < 2; ++k) {
            nd[8 * k + 0] = box.lo.x; nd[8 * k + 1] = box.lo.y; nd[8 * k + 2] = box.lo.z;
            nd[8 * k + 4] = box.hi.x; nd[8 * k + 5] = box.hi.y; nd[8 * k + 6] = box.hi.z;
        }
        return;
    }
    // phase 1: up through the nodes inside this tile (shared memory only); a walker that reaches a node shared with other
    // CTAs parks there (p, box) until the tile's boxes have been written out
    while (walking) {
        const uint32_t node = p >> 1, slot = p & 1u;
        const uint32_t li = node - tile_base;                 // < REFIT_TILE exactly for the nodes indexed inside this tile
        if (!(li < (uint32_t)REFIT_TILE && (s_local[li] & 1))) break;
        // both children of this node come from threads of this CTA: ~50 cycles per level instead of a round trip to L2
        float* mine = s_box[li] + 6 * slot;
        mine[0] = box.lo.x; mine[1] = box.lo.y; mine[2] = box.lo.z; mine[3] = box.hi.x; mine[4] = box.hi.y; mine[5] = box.hi.z;
        __threadfence_block();
        if (atomicAdd(&s_arrived[li], 1u) == 0u) { walking = false; break; }         // first arrival: the sibling will carry on
        __threadfence_block();
        const volatile float* sib = s_box[li] + 6 * (1u - slot);
        box.lo.x = fminf(box.lo.x, sib[0]); box.lo.y = fminf(box.lo.y, sib[1]); box.lo.z = fminf(box.lo.z, sib[2]);
        box.hi.x = fmaxf(box.hi.x, sib[3]); box.hi.y = fmaxf(box.hi.y, sib[4]); box.hi.z = fmaxf(box.hi.z, sib[5]);
        if (node == 0) { walking = false; break; }
        p = s_parent[li];
    }
    }   // if (walking)
    __syncthreads();
    // the tile's local nodes: whole records, one 16-byte store per thread and quarter record (consecutive threads -> consecutive addresses):
    // {lo0, child0} {hi0, child1} {lo1, prim0} {hi1, prim1}
    for (uint32_t k = threadIdx.x; k < (uint32_t)REFIT_TILE * 4u; k += REFIT_TILE) {
        const uint32_t li = k >> 2, q = k & 3u;
        const uint32_t fl = s_local[li];
        if (fl & 1u) {
            const uint32_t c0 = s_child0[li], g = (c0 & ~VKHRT_BVH_LEAF) - tile_base;        // child positions g, g + 1: inside the tile
            const bool leaf1 = (fl & 2u) != 0u;
            const uint32_t w = q == 0u ? c0 : (q == 1u ? ((c0 & ~VKHRT_BVH_LEAF) + 1u) | (leaf1 ? VKHRT_BVH_LEAF : 0u)
                                                       : (q == 2u ? ((c0 & VKHRT_BVH_LEAF) ? s_leaf[g] : 0u) : (leaf1 ? s_leaf[g + 1u] : 0u)));
            const float* b = s_box[li] + 3u * q;
            reinterpret_cast<float4*>(nodes_f32)[(size_t)(tile_base + li) * 4 + q] = make_float4(b[0], b[1], b[2], __uint_as_float(w));
        }
    }
    // phase 2: the few walkers that left the tile (a handful per CTA) go on through the nodes shared between CTAs in a kernel of
    // their own (upper_refit_kernel), so that this CTA's 512 threads do not stay resident for one thread's chain of L2 round trips
    if (walking) {
        const uint32_t at = atomicAdd(exit_count, 1u);
        if (at < exit_cap) {
            exits[2 * (size_t)at] = make_float4(box.lo.x, box.lo.y, box.lo.z, __uint_as_float(p));
            exits[2 * (size_t)at + 1] = make_float4(box.hi.x, box.hi.y, box.hi.z, 0.0f);
            walking = false;
        }
    }
    while (walking) {     // (list full: finish here)
        const uint32_t node = p >> 1, slot = p & 1u;
        // publish my box in the parent's slot (L2 is the coherence point: .cg stores / loads), then ONE acq_rel atomic both
        // releases it and, for the second arrival, acquires the sibling's
        float* nd = nodes_f32 + (size_t)node * 16 + 8 * slot;
        __stcg(reinterpret_cast<float2*>(nd), make_float2(box.lo.x, box.lo.y)); __stcg(nd + 2, box.lo.z);
        __stcg(reinterpret_cast<float2*>(nd + 4), make_float2(box.hi.x, box.hi.y)); __stcg(nd + 6, box.hi.z);
        uint32_t arrived;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(arrived) : "l"(flags + node) : "memory");
        if (arrived == 0u) break;                            // first arrival: the sibling will carry on
        flags[node] = 0u;                                    // clean for the next pass
        const float* sb = nodes_f32 + (size_t)node * 16 + 8 * (1u - slot);
        const float2 l01 = __ldcg(reinterpret_cast<const float2*>(sb)), h01 = __ldcg(reinterpret_cast<const float2*>(sb + 4));
        const float lz = __ldcg(sb + 2), hz = __ldcg(sb + 6);
        box.lo.x = fminf(box.lo.x, l01.x); box.lo.y = fminf(box.lo.y, l01.y); box.lo.z = fminf(box.lo.z, lz);
        box.hi.x = fmaxf(box.hi.x, h01.x); box.hi.y = fmaxf(box.hi.y, h01.y); box.hi.z = fmaxf(box.hi.z, hz);
        if (node == 0) break;
        p = parent_internal[node];
    }
}

// the walkers that left their tile: one thread each, up through the nodes shared between CTAs
__global__ void __launch_bounds__(256) upper_refit_kernel(const float4* __restrict__ exits, const uint32_t* __restrict__ exit_count, uint32_t exit_cap,
                                                          float* nodes_f32, const uint32_t* __restrict__ parent_internal, uint32_t* flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(*exit_count, exit_cap)) return;
    const float4 e0 = exits[2 * (size_t)i], e1 = exits[2 * (size_t)i + 1];
    Aabb box;
    box.lo = xyz(e0); box.hi = xyz(e1);
    uint32_t p = __float_as_uint(e0.w);
    for (;;) {
        const uint32_t node = p >> 1, slot = p & 1u;
        const uint32_t up = __ldg(parent_internal + node);     // fetched while the stores and the atomic are in flight (used only by the second arrival)
        float* nd = nodes_f32 + (size_t)node * 16 + 8 * slot;
        __stcg(reinterpret_cast<float2*>(nd), make_float2(box.lo.x, box.lo.y)); __stcg(nd + 2, box.lo.z);
        __stcg(reinterpret_cast<float2*>(nd + 4), make_float2(box.hi.x, box.hi.y)); __stcg(nd + 6, box.hi.z);
        uint32_t arrived;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(arrived) : "l"(flags + node) : "memory");
        if (arrived == 0u) return;
        flags[node] = 0u;                                      // both children are in: the counter is clean for the next pass (no memset per refit)
        const float* sb = nodes_f32 + (size_t)node * 16 + 8 * (1u - slot);
        const float2 l01 = __ldcg(reinterpret_cast<const float2*>(sb)), h01 = __ldcg(reinterpret_cast<const float2*>(sb + 4));
        const float lz = __ldcg(sb + 2), hz = __ldcg(sb + 6);
        box.lo.x = fminf(box.lo.x, l01.x); box.lo.y = fminf(box.lo.y, l01.y); box.lo.z = fminf(box.lo.z, lz);
        box.hi.x = fmaxf(box.hi.x, h01.x); box.hi.y = fmaxf(box.hi.y, h01.y); box.hi.z = fmaxf(box.hi.z, hz);
        if (node == 0) return;
        p = up;
    }
}

// single-primitive scene: node 0 = the same leaf twice
__global__ void single_node_kernel(uint32_t* nodes_u32, const uint32_t* sorted_ids)
{
    nodes_u32[3] = VKHRT_BVH_LEAF | 0u; nodes_u32[7] = VKHRT_BVH_LEAF | 0u;
    nodes_u32[11] = sorted_ids[0]; nodes_u32[15] = sorted_ids[0];
}

// primitives in ORIGINAL order for vkhrt_scene_get_primitives (ModelCreation buffers)
template <int TECH>
__global__ void __launch_bounds__(256) export_kernel(MeshIn m, uint32_t n_prims, float* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prims) return;
    if (TECH == VKHRT_TECHNIQUE_PHANTOM) {
        Bezier c = gen_curve(m, i);
        float* o = out + (size_t)i * 12;
        o[0] = c.p0.x; o[1] = c.p0.y; o[2] = c.p0.z; o[3] = c.p1.x; o[4] = c.p1.y; o[5] = c.p1.z;
        o[6] = c.p2.x; o[7] = c.p2.y; o[8] = c.p2.z; o[9] = c.p3.x; o[10] = c.p3.y; o[11] = c.p3.z;
    } else if (TECH == VKHRT_TECHNIQUE_LSS) {
        LssPrim s = gen_lss(m, i);
        float* o = out + (size_t)i * 8;
        o[0] = s.p0.x; o[1] = s.p0.y; o[2] = s.p0.z; o[3] = s.r0; o[4] = s.p1.x; o[5] = s.p1.y; o[6] = s.p1.z; o[7] = s.r1;
    } else {
        TriPrim t = gen_tri(m, i);
        float* o = out + (size_t)i * 9;
        o[0] = t.v0.x; o[1] = t.v0.y; o[2] = t.v0.z; o[3] = t.v1.x; o[4] = t.v1.y; o[5] = t.v1.z; o[6] = t.v2.x; o[7] = t.v2.y; o[8] = t.v2.z;
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
static inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

static float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

int build_scene(DeviceScene& sc, bool refit_only)
{
    VK_CUDA(cudaSetDevice(sc.device));
    cudaStream_t st = sc.stream;
    const uint32_t n = sc.n_leaves;
    MeshIn m{sc.d_positions, sc.d_indices, sc.d_radius_pv, sc.n_segments, sc.radius, sc.d_curves};
    sc.timing = VkhrtTiming{};
    sc.timing.lod_ms = sc.lod_ms;
    if (n == 0) { sc.n_nodes = 0; sc.built = true; return VKHRT_OK; }
    const int tech = sc.technique;
    const size_t primA_per = tech == VKHRT_TECHNIQUE_DOTS ? 4 : 2;

    cudaEvent_t* ev = sc.ev;
    if (!refit_only) {
        sc.n_nodes = n > 1 ? n - 1 : 1;
        // one allocation for everything the scene keeps (a cudaMalloc per array used to cost more than the build's kernels) ...
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t o_nodes = 0, o_ids = o_nodes + up((size_t)sc.n_nodes * 64), o_morton = o_ids + up((size_t)n * 4), o_pint = o_morton + up((size_t)n * 8),
                     o_pleaf = o_pint + up((size_t)sc.n_nodes * 4), o_flags = o_pleaf + up((size_t)n * 4), o_primA = o_flags + up((size_t)sc.n_nodes * 4),
                     o_primB = o_primA + up((size_t)n * primA_per * 16), o_primR = o_primB + (tech == VKHRT_TECHNIQUE_PHANTOM ? up((size_t)n * 32) : 0),
                     o_local = o_primR + ((tech == VKHRT_TECHNIQUE_PHANTOM && sc.d_radius_pv) ? up((size_t)n * 8) : 0),
                     o_child0 = o_local + up((size_t)sc.n_nodes), o_exits = o_child0 + up((size_t)sc.n_nodes * 4), o_exit_count = o_exits + up(((size_t)n / 8 + 4096) * 32),
                     total = o_exit_count + 256;
        if (sc.arena_bytes < total) {
            if (sc.d_arena) cudaFree(sc.d_arena);
            sc.d_arena = nullptr; sc.arena_bytes = 0;
            VK_CUDA(cudaMalloc((void**)&sc.d_arena, total));
            sc.arena_bytes = total;
        }
        sc.d_nodes = (float4*)(sc.d_arena + o_nodes); sc.d_sorted_ids = (uint32_t*)(sc.d_arena + o_ids); sc.d_sorted_morton = (uint64_t*)(sc.d_arena + o_morton);
        sc.d_parent_internal = (uint32_t*)(sc.d_arena + o_pint); sc.d_parent_leaf = (uint32_t*)(sc.d_arena + o_pleaf);
        sc.d_refit_flags = (uint32_t*)(sc.d_arena + o_flags); sc.d_primA = (float4*)(sc.d_arena + o_primA);
        sc.d_primB = tech == VKHRT_TECHNIQUE_PHANTOM ? (float4*)(sc.d_arena + o_primB) : nullptr;
        sc.d_primR = (tech == VKHRT_TECHNIQUE_PHANTOM && sc.d_radius_pv) ? (float2*)(sc.d_arena + o_primR) : nullptr;
        sc.d_node_local = (uint8_t*)(sc.d_arena + o_local); sc.d_child0 = (uint32_t*)(sc.d_arena + o_child0);
        sc.d_refit_exits = (float4*)(sc.d_arena + o_exits); sc.d_refit_exit_count = (uint32_t*)(sc.d_arena + o_exit_count);
        sc.refit_exit_cap = (uint32_t)std::min<size_t>((size_t)n / 8 + 4096, 0x7FFFFFFFu);
        // test switch (read once): a tiny exit list makes the walkers that do not fit finish inside materialise_refit_kernel, the path a
        // scene with more tile exits than the list holds would take
        static const int cap_override = [] { const char* v = getenv("VKHRT_REFIT_EXIT_CAP"); return v ? atoi(v) : 0; }();
        if (cap_override > 0) sc.refit_exit_cap = std::min<uint32_t>(sc.refit_exit_cap, (uint32_t)cap_override);

        // ... and one for the build's scratch, kept with the scene (a per-frame rebuild of a dynamic groom allocates nothing)
        const uint32_t os_tiles = cdiv(n, OS_TILE);
        const size_t s_cent = 0, s_bounds = s_cent + up((size_t)n * 16), s_keys = s_bounds + 256, s_vals = s_keys + up((size_t)n * 8),
                     s_hist = s_vals + up((size_t)n * 4), s_ticket = s_hist + up(OS_PASSES * 256 * 4), s_status = s_ticket + 256,
                     s_total = s_status + up((size_t)os_tiles * 256 * 8);
        if (sc.build_scratch_bytes < s_total) {
            if (sc.d_build_scratch) cudaFree(sc.d_build_scratch);
            sc.d_build_scratch = nullptr; sc.build_scratch_bytes = 0;
            VK_CUDA(cudaMalloc((void**)&sc.d_build_scratch, s_total));
            sc.build_scratch_bytes = s_total;
        }
        unsigned char* d_scratch = sc.d_build_scratch;
#define VK_CUDA_S(call) VK_CUDA(call)
        float4* d_cent = (float4*)(d_scratch + s_cent); uint32_t* d_bounds = (uint32_t*)(d_scratch + s_bounds);
        uint64_t* d_keys_alt = (uint64_t*)(d_scratch + s_keys); uint32_t* d_vals_alt = (uint32_t*)(d_scratch + s_vals);
        uint32_t* d_hist = (uint32_t*)(d_scratch + s_hist); uint32_t* d_ticket = (uint32_t*)(d_scratch + s_ticket);
        unsigned long long* d_status = (unsigned long long*)(d_scratch + s_status);
        VK_CUDA(cudaEventRecord(ev[0], st));            // the build proper starts here: allocations are not part of it

        // 1a centroids + bounds
        uint32_t init_bounds[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
        VK_CUDA_S(cudaMemcpyAsync(d_bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, st));
        const uint32_t g = cdiv(n, 256);
        if (tech == VKHRT_TECHNIQUE_PHANTOM) centroid_kernel<VKHRT_TECHNIQUE_PHANTOM><<<g, 256, 0, st>>>(m, n, d_cent, d_bounds);
        else if (tech == VKHRT_TECHNIQUE_LSS) centroid_kernel<VKHRT_TECHNIQUE_LSS><<<g, 256, 0, st>>>(m, n, d_cent, d_bounds);
        else centroid_kernel<VKHRT_TECHNIQUE_DOTS><<<g, 256, 0, st>>>(m, n, d_cent, d_bounds);
        count_launch();
        VK_CUDA_S(cudaEventRecord(ev[1], st));
        // 1b morton keys
        uint64_t* keys[2] = {sc.d_sorted_morton, d_keys_alt};
        uint32_t* vals[2] = {sc.d_sorted_ids, d_vals_alt};
        morton_kernel<<<g, 256, 0, st>>>(d_cent, n, d_bounds, keys[0], vals[0]);
        count_launch();
        VK_CUDA_S(cudaEventRecord(ev[2], st));
        // 2 onesweep radix sort: 8 digit passes over the 63-bit key (even count => the result lands back in keys[0]/vals[0])
        VK_CUDA_S(cudaMemsetAsync(d_hist, 0, (size_t)(s_total - s_hist), st));          // histograms, tickets, status words
        os_histogram_kernel<<<std::min<uint32_t>(cdiv(n, 512 * 8), (uint32_t)sc.sm_count * 4u), 512, 0, st>>>(keys[0], n, d_hist);
        os_scan_kernel<<<1, 256, 0, st>>>(d_hist);
        count_launch(2);
        int cur = 0;
        for (int pass = 0; pass < OS_PASSES; ++pass) {
            os_pass_kernel<<<os_tiles, OS_BLOCK, 0, st>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, pass, d_hist + pass * 256, d_status, d_ticket);
            count_launch();
            cur ^= 1;
        }
        VK_CUDA_S(cudaEventRecord(ev[3], st));
        // 3 hierarchy
        VK_CUDA_S(cudaMemsetAsync(sc.d_nodes, 0, (size_t)sc.n_nodes * 64, st));
        if (n > 1) {
            karras_kernel<<<cdiv(n - 1, 256), 256, 0, st>>>(sc.d_sorted_morton, sc.d_sorted_ids, (int)n, (uint32_t*)sc.d_nodes,
                                                            sc.d_parent_internal, sc.d_parent_leaf, sc.d_node_local, sc.d_child0, refit_tile_shift(tech));
        } else {
            single_node_kernel<<<1, 1, 0, st>>>((uint32_t*)sc.d_nodes, sc.d_sorted_ids);
        }
        count_launch();
        VK_CUDA_S(cudaEventRecord(ev[4], st));
        VK_CUDA_S(cudaMemcpyAsync(sc.h_bounds, d_bounds, sizeof(sc.h_bounds), cudaMemcpyDeviceToHost, st));     // read after the final synchronise
#undef VK_CUDA_S
    } else {
        VK_CUDA(cudaEventRecord(ev[0], st));
        VK_CUDA(cudaEventRecord(ev[1], st)); VK_CUDA(cudaEventRecord(ev[2], st));
        VK_CUDA(cudaEventRecord(ev[3], st)); VK_CUDA(cudaEventRecord(ev[4], st));
    }
    // 4 materialise + refit (the arrival counters of the nodes shared between tiles clean themselves: zeroed once per build)
    if (!refit_only) VK_CUDA(cudaMemsetAsync(sc.d_refit_flags, 0, (size_t)sc.n_nodes * 4, st));
    VK_CUDA(cudaMemsetAsync(sc.d_refit_exit_count, 0, 4, st));
    const uint32_t g = cdiv(n, 1u << refit_tile_shift(tech));
    if (tech == VKHRT_TECHNIQUE_PHANTOM)
        materialise_refit_kernel<VKHRT_TECHNIQUE_PHANTOM><<<g, RefitTile<VKHRT_TECHNIQUE_PHANTOM>::N, 0, st>>>(m, n, sc.d_sorted_ids, sc.d_primA, sc.d_primB, sc.d_primR, (float*)sc.d_nodes, sc.d_parent_internal, sc.d_parent_leaf, sc.d_refit_flags, sc.d_node_local, sc.d_child0, sc.d_refit_exits, sc.d_refit_exit_count, sc.refit_exit_cap);
    else if (tech == VKHRT_TECHNIQUE_LSS)
        materialise_refit_kernel<VKHRT_TECHNIQUE_LSS><<<g, RefitTile<VKHRT_TECHNIQUE_LSS>::N, 0, st>>>(m, n, sc.d_sorted_ids, sc.d_primA, sc.d_primB, sc.d_primR, (float*)sc.d_nodes, sc.d_parent_internal, sc.d_parent_leaf, sc.d_refit_flags, sc.d_node_local, sc.d_child0, sc.d_refit_exits, sc.d_refit_exit_count, sc.refit_exit_cap);
    else
        materialise_refit_kernel<VKHRT_TECHNIQUE_DOTS><<<g, RefitTile<VKHRT_TECHNIQUE_DOTS>::N, 0, st>>>(m, n, sc.d_sorted_ids, sc.d_primA, sc.d_primB, sc.d_primR, (float*)sc.d_nodes, sc.d_parent_internal, sc.d_parent_leaf, sc.d_refit_flags, sc.d_node_local, sc.d_child0, sc.d_refit_exits, sc.d_refit_exit_count, sc.refit_exit_cap);
    // the walkers that left their tiles (their number is only known on the device: the grid covers the list's capacity, idle threads return at once)
    if (n > 1) upper_refit_kernel<<<cdiv(sc.refit_exit_cap, 256), 256, 0, st>>>(sc.d_refit_exits, sc.d_refit_exit_count, sc.refit_exit_cap, (float*)sc.d_nodes, sc.d_parent_internal, sc.d_refit_flags);
    count_launch(2);
    VK_CUDA(cudaEventRecord(ev[5], st));
    VK_CUDA(cudaStreamSynchronize(st));
    VK_CUDA(cudaGetLastError());
    if (!refit_only) for (int k = 0; k < 3; ++k) { sc.scene_lo[k] = ord2f_host(sc.h_bounds[k]); sc.scene_hi[k] = ord2f_host(sc.h_bounds[3 + k]); }
    sc.timing.geometry_ms = ev_ms(ev[0], ev[1]);
    sc.timing.morton_ms = ev_ms(ev[1], ev[2]);
    sc.timing.sort_ms = ev_ms(ev[2], ev[3]);
    sc.timing.hierarchy_ms = ev_ms(ev[3], ev[4]);
    sc.timing.refit_ms = ev_ms(ev[4], ev[5]);
    sc.timing.build_total_ms = ev_ms(ev[0], ev[5]);
    sc.built = true;
    return VKHRT_OK;
}

int export_primitives(DeviceScene& sc, float* host_out, size_t out_floats)
{
    VK_CUDA(cudaSetDevice(sc.device));
    const uint32_t n = sc.n_prims;
    const size_t per = sc.technique == VKHRT_TECHNIQUE_PHANTOM ? 12 : (sc.technique == VKHRT_TECHNIQUE_LSS ? 8 : 9);
    if (out_floats < (size_t)n * per) { set_last_error("export_primitives: output too small"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (n == 0) return VKHRT_OK;
    MeshIn m{sc.d_positions, sc.d_indices, sc.d_radius_pv, sc.n_segments, sc.radius, sc.d_curves};
    float* d_out = nullptr;
    VK_CUDA(cudaMalloc(&d_out, (size_t)n * per * 4));
    const uint32_t g = cdiv(n, 256);
    if (sc.technique == VKHRT_TECHNIQUE_PHANTOM) export_kernel<VKHRT_TECHNIQUE_PHANTOM><<<g, 256, 0, sc.stream>>>(m, n, d_out);
    else if (sc.technique == VKHRT_TECHNIQUE_LSS) export_kernel<VKHRT_TECHNIQUE_LSS><<<g, 256, 0, sc.stream>>>(m, n, d_out);
    else export_kernel<VKHRT_TECHNIQUE_DOTS><<<g, 256, 0, sc.stream>>>(m, n, d_out);
    count_launch();
    cudaError_t e = cudaMemcpyAsync(host_out, d_out, (size_t)n * per * 4, cudaMemcpyDeviceToHost, sc.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sc.stream);
    cudaFree(d_out);
    if (e != cudaSuccess) { set_last_error(std::string("export_primitives: ") + cudaGetErrorString(e)); return VKHRT_ERR_CUDA; }
    return VKHRT_OK;
}

// ------------------------------------------------------------------------------------------------
// Strand level of detail: MergeLines / SplitLines / MergeCurvesFast of
// source/resources/model/geometry_processor.cpp:69-104, 106-121, 158-197 as device passes over the line list.
// A line is two float4 {start.xyz, r0} {end.xyz, r1} (the per-vertex radius rides along: outer radii survive a merge,
// a split midpoint gets the mean).  The merges are stream compactions: pair i = elements (2i, 2i+1) emits one element
// if the two are connected (exact float compare of end/start, glm::vec3 ==) and both otherwise; like the reference,
// the last element of an odd-sized list is dropped.  Output offsets come from the block scan of the radix sort.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lines_from_mesh_kernel(MeshIn m, uint32_t n, float4* __restrict__ lines)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = m.idx[2 * i], b = m.idx[2 * i + 1];
    float3 s = load_pos(m, a), e = load_pos(m, b);
    lines[2 * (size_t)i] = make_float4(s.x, s.y, s.z, m.rpv ? m.rpv[a] : m.radius);
    lines[2 * (size_t)i + 1] = make_float4(e.x, e.y, e.z, m.rpv ? m.rpv[b] : m.radius);
}
__global__ void __launch_bounds__(256) lines_to_mesh_kernel(const float4* __restrict__ lines, uint32_t n, float* __restrict__ pos,
                                                            uint32_t* __restrict__ idx, float* __restrict__ rpv)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 s = lines[2 * (size_t)i], e = lines[2 * (size_t)i + 1];
    float* p = pos + 6 * (size_t)i;
    p[0] = s.x; p[1] = s.y; p[2] = s.z; p[3] = e.x; p[4] = e.y; p[5] = e.z;
    idx[2 * (size_t)i] = 2 * i; idx[2 * (size_t)i + 1] = 2 * i + 1;
    if (rpv) { rpv[2 * (size_t)i] = s.w; rpv[2 * (size_t)i + 1] = e.w; }
}
// SplitLines: middlePoint = (start + end) * 0.5f
__global__ void __launch_bounds__(256) split_lines_kernel(const float4* __restrict__ in, uint32_t n, float4* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 s = in[2 * (size_t)i], e = in[2 * (size_t)i + 1];
    float4 mid = make_float4((s.x + e.x) * 0.5f, (s.y + e.y) * 0.5f, (s.z + e.z) * 0.5f, (s.w + e.w) * 0.5f);
    float4* o = out + 4 * (size_t)i;
    o[0] = s; o[1] = mid; o[2] = mid; o[3] = e;
}
VK_DEV bool same_xyz(float4 a, float4 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// MergeLines, pass 1: how many lines pair i emits (slot n_pairs = 0 so that the exclusive scan ends with the total)
__global__ void __launch_bounds__(256) merge_lines_count_kernel(const float4* __restrict__ in, uint32_t n_pairs, uint32_t* __restrict__ counts)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_pairs) return;
    counts[i] = i == n_pairs ? 0u : (same_xyz(in[4 * (size_t)i + 1], in[4 * (size_t)i + 2]) ? 1u : 2u);
}
__global__ void __launch_bounds__(256) merge_lines_scatter_kernel(const float4* __restrict__ in, uint32_t n_pairs, const uint32_t* __restrict__ offsets,
                                                                  float4* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const float4* a = in + 4 * (size_t)i;
    float4* o = out + 2 * (size_t)offsets[i];
    if (same_xyz(a[1], a[2])) { o[0] = a[0]; o[1] = a[3]; }
    else { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = a[3]; }
}
// GenerateCurves into an explicit array (input of MergeCurvesFast)
__global__ void __launch_bounds__(256) curves_materialise_kernel(MeshIn m, uint32_t n, float* __restrict__ curves)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Bezier c = gen_curve(m, i);
    float* o = curves + 12 * (size_t)i;
    o[0] = c.p0.x; o[1] = c.p0.y; o[2] = c.p0.z; o[3] = c.p1.x; o[4] = c.p1.y; o[5] = c.p1.z;
    o[6] = c.p2.x; o[7] = c.p2.y; o[8] = c.p2.z; o[9] = c.p3.x; o[10] = c.p3.y; o[11] = c.p3.z;
}
VK_DEV bool curves_connected(const float* a, const float* b) { return a[9] == b[0] && a[10] == b[1] && a[11] == b[2]; }
__global__ void __launch_bounds__(256) merge_curves_count_kernel(const float* __restrict__ in, uint32_t n_pairs, uint32_t* __restrict__ counts)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_pairs) return;
    counts[i] = i == n_pairs ? 0u : (curves_connected(in + 24 * (size_t)i, in + 24 * (size_t)i + 12) ? 1u : 2u);
}
// MergeCurvesFast: middlePoint = (a.cp2 + b.cp1) * 0.5; cp1 = (a.cp1 + middlePoint) * 0.5; cp2 = (middlePoint + b.cp2) * 0.5
__global__ void __launch_bounds__(256) merge_curves_scatter_kernel(const float* __restrict__ in, uint32_t n_pairs, const uint32_t* __restrict__ offsets,
                                                                   float* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const float* a = in + 24 * (size_t)i;
    const float* b = a + 12;
    float* o = out + 12 * (size_t)offsets[i];
    if (curves_connected(a, b)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float mid = (a[6 + k] + b[3 + k]) * 0.5f;
            o[k] = a[k];
            o[3 + k] = (a[3 + k] + mid) * 0.5f;
            o[6 + k] = (mid + b[6 + k]) * 0.5f;
            o[9 + k] = b[9 + k];
        }
    } else {
        for (int k = 0; k < 24; ++k) o[k] = a[k];
    }
}
__global__ void __launch_bounds__(256) lines_from_curves_kernel(const float* __restrict__ curves, uint32_t n, float radius, float4* __restrict__ lines)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* c = curves + 12 * (size_t)i;
    lines[2 * (size_t)i] = make_float4(c[0], c[1], c[2], radius);
    lines[2 * (size_t)i + 1] = make_float4(c[9], c[10], c[11], radius);
}
__global__ void __launch_bounds__(256) export_lines_kernel(MeshIn m, uint32_t n, float* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 s = load_pos(m, m.idx[2 * i]), e = load_pos(m, m.idx[2 * i + 1]);
    float* o = out + 6 * (size_t)i;
    o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = e.x; o[4] = e.y; o[5] = e.z;
}

// one compaction pass into a preallocated output (a pair emits at most its two inputs): counts -> exclusive scan -> scatter,
// then the compacted size is read back (the next pass pairs up exactly that many elements)
template <typename T>
static int compact_pairs(void (*count_k)(const T*, uint32_t, uint32_t*), void (*scatter_k)(const T*, uint32_t, const uint32_t*, T*),
                         const T* in, uint32_t n, T* out, uint32_t* d_cnt, uint32_t* d_tmp, uint32_t* n_out, cudaStream_t st)
{
    const uint32_t n_pairs = n / 2;
    *n_out = 0;
    count_k<<<cdiv((uint64_t)n_pairs + 1, 256), 256, 0, st>>>(in, n_pairs, d_cnt);
    count_launch();
    int rc = exclusive_scan(d_cnt, n_pairs + 1, d_tmp, st);
    if (rc) return rc;
    if (n_pairs) { scatter_k<<<cdiv(n_pairs, 256), 256, 0, st>>>(in, n_pairs, d_cnt, out); count_launch(); }
    uint32_t total = 0;
    VK_CUDA(cudaMemcpyAsync(&total, d_cnt + n_pairs, 4, cudaMemcpyDeviceToHost, st));
    VK_CUDA(cudaStreamSynchronize(st));
    *n_out = total;
    return VKHRT_OK;
}

int apply_lod(DeviceScene& sc, uint32_t split_passes, uint32_t merge_passes, uint32_t curve_merge_passes)
{
    VK_CUDA(cudaSetDevice(sc.device));
    if (sc.built) { set_last_error("vkhrt_scene_apply_lod must be called before vkhrt_scene_build"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (curve_merge_passes && sc.technique != VKHRT_TECHNIQUE_PHANTOM) { set_last_error("curve_merge_passes needs the PHANTOM technique (curves)"); return VKHRT_ERR_UNSUPPORTED; }
    if (!split_passes && !merge_passes && !curve_merge_passes) return VKHRT_OK;
    if (split_passes > 30 || ((uint64_t)sc.n_segments << split_passes) * (sc.technique == VKHRT_TECHNIQUE_DOTS ? 4 : 1) >= 0x3FFFFFFFull) {
        set_last_error("lod: too many segments after splitting"); return VKHRT_ERR_UNSUPPORTED;
    }
    cudaStream_t st = sc.stream;
    MeshIn m{sc.d_positions, sc.d_indices, sc.d_radius_pv, sc.n_segments, sc.radius, sc.d_curves};
    uint32_t n = sc.n_segments;
    VK_CUDA(cudaEventRecord(sc.ev[13], st));

    // ONE scratch allocation for all passes: two ping-pong line buffers of the largest line count, the same for curves, the
    // temporary indexed mesh GenerateCurves reads, pair counts and the scan's block totals
    const size_t N = std::max<size_t>(1, (size_t)n << split_passes);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_l0 = 0, o_l1 = o_l0 + up(N * 32), o_c0 = o_l1 + up(N * 32), o_c1 = o_c0 + (curve_merge_passes ? up(N * 48) : 0),
                 o_pos = o_c1 + (curve_merge_passes ? up(N * 48) : 0), o_idx = o_pos + (curve_merge_passes ? up(N * 24) : 0),
                 o_cnt = o_idx + (curve_merge_passes ? up(N * 8) : 0), o_tmp = o_cnt + up((N / 2 + 1) * 4),
                 total_bytes = o_tmp + up(scan_tmp_elems((uint32_t)(N / 2 + 1)) * 4);
    unsigned char* arena = nullptr;
    VK_CUDA(cudaMalloc((void**)&arena, total_bytes));
    float4* lines = (float4*)(arena + o_l0); float4* lines_alt = (float4*)(arena + o_l1);
    float* curves = (float*)(arena + o_c0); float* curves_alt = (float*)(arena + o_c1);
    uint32_t* d_cnt = (uint32_t*)(arena + o_cnt); uint32_t* d_tmp = (uint32_t*)(arena + o_tmp);
    float* pos = nullptr; uint32_t* idx = nullptr; float* rpv = nullptr; float* curves_out = nullptr;
    auto bail = [&](int code, const char* what, cudaError_t e) {
        set_last_error(std::string("lod: ") + what + (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : std::string()));
        cudaFree(arena); cudaFree(pos); cudaFree(idx); cudaFree(rpv); cudaFree(curves_out);
        return code;
    };
    if (n) { lines_from_mesh_kernel<<<cdiv(n, 256), 256, 0, st>>>(m, n, lines); count_launch(); }
    for (uint32_t k = 0; k < split_passes; ++k) {
        if (n) { split_lines_kernel<<<cdiv(n, 256), 256, 0, st>>>(lines, n, lines_alt); count_launch(); }
        std::swap(lines, lines_alt); n *= 2;
    }
    for (uint32_t k = 0; k < merge_passes; ++k) {
        uint32_t n_out = 0;
        int rc = compact_pairs<float4>(merge_lines_count_kernel, merge_lines_scatter_kernel, lines, n, lines_alt, d_cnt, d_tmp, &n_out, st);
        if (rc) { cudaFree(arena); return rc; }
        std::swap(lines, lines_alt); n = n_out;
    }
    // optionally through the curve merge: GenerateCurves over a temporary indexed mesh, MergeCurvesFast passes, lines = curve ends
    if (curve_merge_passes) {
        float* tpos = (float*)(arena + o_pos); uint32_t* tidx = (uint32_t*)(arena + o_idx);
        if (n) {
            lines_to_mesh_kernel<<<cdiv(n, 256), 256, 0, st>>>(lines, n, tpos, tidx, nullptr);
            MeshIn lm{tpos, tidx, nullptr, n, sc.radius, nullptr};
            curves_materialise_kernel<<<cdiv(n, 256), 256, 0, st>>>(lm, n, curves);
            count_launch(2);
        }
        for (uint32_t k = 0; k < curve_merge_passes; ++k) {
            uint32_t n_out = 0;
            int rc = compact_pairs<float>(merge_curves_count_kernel, merge_curves_scatter_kernel, curves, n, curves_alt, d_cnt, d_tmp, &n_out, st);
            if (rc) { cudaFree(arena); return rc; }
            std::swap(curves, curves_alt); n = n_out;
        }
        if (n) { lines_from_curves_kernel<<<cdiv(n, 256), 256, 0, st>>>(curves, n, sc.radius, lines); count_launch(); }
    }
    // back to the indexed shape the generators consume (2 vertices per line); these arrays stay with the scene
    cudaError_t e = cudaMalloc(&pos, std::max<size_t>(1, (size_t)n * 6) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&idx, std::max<size_t>(1, (size_t)n * 2) * 4);
    if (e == cudaSuccess && sc.d_radius_pv && !curve_merge_passes) e = cudaMalloc(&rpv, std::max<size_t>(1, (size_t)n * 2) * 4);
    if (e == cudaSuccess && curve_merge_passes) e = cudaMalloc(&curves_out, std::max<size_t>(1, (size_t)n * 12) * 4);
    if (e != cudaSuccess) return bail(e == cudaErrorMemoryAllocation ? VKHRT_ERR_OUT_OF_MEMORY : VKHRT_ERR_CUDA, "allocation", e);
    if (n) { lines_to_mesh_kernel<<<cdiv(n, 256), 256, 0, st>>>(lines, n, pos, idx, rpv); count_launch(); }
    if (curve_merge_passes && n) e = cudaMemcpyAsync(curves_out, curves, (size_t)n * 48, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord(sc.ev[14], st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return bail(VKHRT_ERR_CUDA, "passes", e);
    cudaFree(arena);
    cudaFree(sc.d_positions); cudaFree(sc.d_indices); cudaFree(sc.d_radius_pv); cudaFree(sc.d_curves);
    sc.d_positions = pos; sc.d_indices = idx; sc.d_radius_pv = rpv; sc.d_curves = curves_out;
    sc.n_vertices = 2 * n; sc.n_segments = n; sc.n_leaves = n * leaf_split_of(sc.technique);
    sc.n_prims = sc.technique == VKHRT_TECHNIQUE_DOTS ? 4 * n : n;
    sc.lod_applied = true;
    sc.lod_ms = ev_ms(sc.ev[13], sc.ev[14]);
    return VKHRT_OK;
}

int export_lines(DeviceScene& sc, float* host_out, size_t out_floats)
{
    VK_CUDA(cudaSetDevice(sc.device));
    const uint32_t n = sc.n_segments;
    if (out_floats < (size_t)n * 6) { set_last_error("export_lines: output too small"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (n == 0) return VKHRT_OK;
    MeshIn m{sc.d_positions, sc.d_indices, sc.d_radius_pv, sc.n_segments, sc.radius, nullptr};
    float* d_out = nullptr;
    VK_CUDA(cudaMalloc(&d_out, (size_t)n * 24));
    export_lines_kernel<<<cdiv(n, 256), 256, 0, sc.stream>>>(m, n, d_out);
    count_launch();
    cudaError_t e = cudaMemcpyAsync(host_out, d_out, (size_t)n * 24, cudaMemcpyDeviceToHost, sc.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sc.stream);
    cudaFree(d_out);
    if (e != cudaSuccess) { set_last_error(std::string("export_lines: ") + cudaGetErrorString(e)); return VKHRT_ERR_CUDA; }
    return VKHRT_OK;
}

}  // namespace vkhrt
