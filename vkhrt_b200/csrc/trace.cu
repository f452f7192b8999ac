// trace.cu — wavefront ray generator, persistent-thread BVH traversal with the Phantom / LSS / DOTS
// intersectors, and the closest-hit shading kernel.
//
// Replaces (reference paths): shaders/ray_gen.rgen:16-48 (ray generation + traceRayEXT + imageStore),
// the driver/RT-core traversal behind vkCmdTraceRaysKHR (source/renderer.cpp:156-166),
// shaders/hair_intersection.rint:132-150 (candidate test + reportIntersectionEXT),
// shaders/hair_closest_hit.rchit:15-25 and triangle_closest_hit.rchit:43-85 (hit attributes, colour).
//
// Compiled with -fmad=false (see hair_math.cuh): hit records are bit-identical to the CPU oracle.
#include "scene.h"
#include "hair_math.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

#ifndef VKHRT_LINE_PROTOCOL
#define VKHRT_LINE_PROTOCOL 1      // line-wise host delivery: 0 = round 1 (fence + acq_rel count + acquire load), 1 = release-only count
#endif

namespace vkhrt {

constexpr int TR_BLOCK = 128;          // 4 warps per CTA
constexpr int TR_MIN_BLOCKS = 6;       // CTAs per SM the register allocation is held to
constexpr int TR_STACK = 24;           // per-lane shared-memory short stack entries (8 B each)
constexpr int TR_SPILL = 80;           // per-lane local-memory overflow (Karras depth <= 64 + 32)
constexpr uint32_t REF_NONE = 0x7FFFFFFFu;
constexpr uint32_t PRIM_NONE = 0xFFFFFFFFu;
constexpr uint32_t FLAG_HIT = 1u, FLAG_PADDING = 2u;

struct TraceParams {
    const float4* nodes;
    const float4* primA;
    const float4* primB;
    const float2* primR;            // per-vertex radius scenes (TAPER kernels): {r0, r1} per Phantom leaf position
    const uint32_t* sorted_ids;
    uint32_t n_prims;
    float radius;
    // fused ray generation
    Camera cam;
    uint32_t W, H;
    float sx, sy, tmin, tmax;
    uint32_t T, tiles_x, n_tiles, tile_first, tile_stride, compact;
    // several samples of the frame in ONE launch (samples >= 1 of small or sharded frames, so that the persistent kernel does not
    // ramp up and drain once per sample): slot = b * slots_per_sample + slot of the pixel, b-th sample offset (bsx, bsy)[b],
    // record at hits[b * n_out + out].  n_batch <= 1: one sample, offset (sx, sy)
    uint32_t n_batch, slots_per_sample;
    float bsx[16], bsy[16];
    // ambient-occlusion mode (SRC_AO): rays are generated from the primary hit records
    const VkhrtHit* ao_hits;        // indexed like `hits`
    uint32_t* ao_occluded;          // per pixel: += 1 for every occluded AO ray (zeroed by the host before the passes)
    uint32_t ao_index, ao_sample;   // which AO ray of the pixel / which spp sample (seeds of the direction hash)
    float ao_distance, ao_bias;
    // wavefront mode
    const float4* rays;
    uint32_t slot_begin, n_slots;   // this launch traces slots [slot_begin, n_slots)
    VkhrtHit* hits;
    VkhrtHit* hits_mirror;          // optional second destination (pinned host memory), same indexing
    uint32_t hits_aligned32;        // bit 0 / 1: hits / hits_mirror are 32-byte aligned (256-bit record stores)
    uint32_t host_dest;             // a destination is mapped host memory (zero-copy over PCIe)
    // line-wise host delivery (trace_pool_kernel): records go to HBM (`hits`); whoever completes a 128-byte line (4 records) has
    // four lanes copy it to the mapped host buffer in ONE store instruction (see tools/micro/pcie_write.cu for why)
    uint32_t* line_cnt;             // records written per line, zeroed per frame (null = off)
    VkhrtHit* host_lines;           // mapped pinned host buffer, same indexing as `hits`
    uint32_t n_out, line_shift;     // a line = 1 << line_shift records (2: 128 bytes)
    unsigned long long* counters;   // [1] nodes, [2] prims, [3] hits, [4] iterations, [5] rays, [8..11] steps, [12..15] lanes
    // work counters [0] and [6] alternate between launches: a launch pulls slots from `work` and zeroes `work_next` for the launch
    // after it (stream-ordered), so no memset sits between the samples of a frame
    unsigned long long* work;
    unsigned long long* work_next;
    uint32_t refill_threshold;      // lanes waiting for a new ray that trigger a refill step
    uint32_t w_node, w_leaf, w_march;   // scheduler weights (fixed point, 16 = 1.0)
    // trace_pool_kernel
    uint2* pool_overflow;           // stack overflow area: [resident warp][slot][PL_OVF]
    uint32_t pool_node_lanes, pool_batch_lanes, pool_node_min, pool_exit_eighths;
};

// slot (processing order: tiles, inside a tile 8x4-pixel blocks so one warp = one coherent packet)
// -> pixel and output index
struct PixelRef { uint32_t px, py; uint32_t out; bool valid; };
VK_DEV PixelRef slot_to_pixel(const TraceParams& p, uint32_t slot)
{
    const uint32_t tt = p.T * p.T;
    uint32_t tl = slot / tt, r = slot % tt;
    uint32_t blk = r >> 5, ln = r & 31u;
    uint32_t bpr = p.T >> 3;
    uint32_t x = (blk % bpr) * 8u + (ln & 7u), y = (blk / bpr) * 4u + (ln >> 3);
    uint32_t tile = p.tile_first + tl * p.tile_stride;
    PixelRef q;
    q.px = (tile % p.tiles_x) * p.T + x;
    q.py = (tile / p.tiles_x) * p.T + y;
    q.valid = tile < p.n_tiles && q.px < p.W && q.py < p.H;
    q.out = p.compact ? (tl * tt + y * p.T + x) : (q.py * p.W + q.px);
    return q;
}

// One 32-byte record = one 256-bit store (STG.E.256, sm_100) when the buffer is 32-byte aligned: a record that goes out to
// pinned host memory then crosses PCIe as ONE 32-byte write instead of two 16-byte ones.
VK_DEV void store_record(VkhrtHit* dst, bool aligned32, const float4& a, const float4& b)
{
    if (aligned32) {
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)),
                     "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)),
                     "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w))
                     : "memory");
    } else {
        float4* h = reinterpret_cast<float4*>(dst);
        h[0] = a; h[1] = b;
    }
}
template <bool WIDE = true>
VK_DEV void store_hit(const TraceParams& p, size_t i, float t, uint32_t seg, float u, float3 n, uint32_t prim, uint32_t flags)
{
    const float4 a = make_float4(t, __uint_as_float(seg), u, n.x);
    const float4 b = make_float4(n.y, n.z, __uint_as_float(prim), __uint_as_float(flags));
    if (p.hits) store_record(p.hits + i, WIDE && (p.hits_aligned32 & 1u) != 0u, a, b);
    if (p.hits_mirror) store_record(p.hits_mirror + i, WIDE && (p.hits_aligned32 & 2u) != 0u, a, b);
}

// ------------------------------------------------------------------------------------------------
// Persistent-thread traversal with a per-warp majority-state scheduler.
//
// One warp owns 32 ray slots; every lane is in exactly one state:
//   NODE   holds an internal BVH2 node: 64-byte record as 4 x LDG.128, two fused slab tests,
//          nearest-first descent, the far child goes to a short stack in shared memory;
//   LEAF   holds a leaf: Phantom -> the Prhi bounding-cylinder early-out (and, on a pass, the change
//          to ray-centric coordinates); LSS / DOTS -> the full primitive test;
//   MARCH  (Phantom) holds a candidate curve: ONE cone iteration of the 2 x <=8 root finder;
//   REFILL ray finished (or lane empty): write the hit record, pull the next slot from the global
//          counter (warp-aggregated atomic) and generate its primary ray;
//   DONE   no rays left.
// Each scheduling step ballots the four live states and executes the code of the most populated one
// (weighted), staying in it while it keeps >= 3/4 of its lanes.  A ray's own sequence of node visits
// and candidate tests is exactly the CPU oracle's (nothing is speculated or postponed past a cull),
// so the hit records AND the traversal counters are identical; only the interleaving across lanes is
// scheduled for SIMD occupancy.  Justification by ncu counters: DESIGN.md §6.
// ------------------------------------------------------------------------------------------------
enum : uint32_t { ST_NODE = 0, ST_POP = 1, ST_LEAF = 2, ST_MARCH = 3, ST_REFILL = 4, ST_DONE = 5 };   // NODE|POP are scheduled together
// where a lane's next ray comes from
enum : int { SRC_PRIMARY = 0,   // generated from the camera (ray_gen.rgen), fused into the refill step
             SRC_BUFFER = 1,    // wavefront ray buffer
             SRC_AO = 2 };      // ambient-occlusion ray spawned from the pixel's primary hit record

// ANYHIT: gl_RayFlagsTerminateOnFirstHitEXT — the ray retires on its first accepted hit (shadow / occlusion rays).
// TAPER: the scene has per-vertex radii (Phantom: radius(t) linear along the curve, cone slant = r1 - r0; DOTS: per-end offsets)
template <int TECH, bool STATS, int SRC, bool ANYHIT, int MINB, bool TAPER = false>
__global__ void __launch_bounds__(TR_BLOCK, MINB) trace_kernel(const TraceParams p)
{
    constexpr bool WAVEFRONT = SRC == SRC_BUFFER;
    __shared__ uint2 s_stack[TR_STACK][TR_BLOCK];
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.work_next = 0ull;
    uint2 spill[TR_SPILL];
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr bool PH = TECH == VKHRT_TECHNIQUE_PHANTOM;

    uint32_t state = ST_REFILL;
    bool have_ray = false;           // REFILL: a finished ray is waiting for its hit record to be written
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1), id = f3(0, 0, 0), noid = f3(0, 0, 0);
    float tmin = 0.0f, tcur = 0.0f, best_u = 0.0f;
    uint32_t best_prim = PRIM_NONE, best_pos = 0;
    uint32_t last_group = PRIM_NONE;     // one-entry mailbox (VKHRT_MAILBOX_PHANTOM): the curve this ray tested last
    MarchState ms;
    ms.c.p0 = ms.c.p1 = ms.c.p2 = ms.c.p3 = f3(0, 0, 0);
    ms.t = ms.told = ms.dt1 = ms.dt2 = ms.t_start = 0.0f; ms.it = 0u; ms.r0 = ms.dr = 0.0f;
    uint32_t out_idx = 0, mpos = 0;
    int sp = 0;
    uint32_t cur = REF_NONE;
    uint32_t st_nodes = 0, st_prims = 0, st_iters = 0, st_hits = 0, st_rays = 0;
    uint32_t sc_steps[4] = {0, 0, 0, 0}, sc_lanes[4] = {0, 0, 0, 0};   // lane 0 only (STATS)

    auto push = [&](uint32_t ref, float tn) {
        uint2 e = make_uint2(ref, __float_as_uint(tn));
        if (sp < TR_STACK) s_stack[sp][tid] = e; else spill[sp - TR_STACK] = e;
        ++sp;
    };
    // ONE stack entry per call (the cull loop runs across scheduling steps, in parallel over lanes):
    // entries whose box entry distance lies beyond the current closest hit are dropped
    auto pop_one = [&]() {
        if (sp == 0) { state = ST_REFILL; return; }
        --sp;
        const uint2 e = sp < TR_STACK ? s_stack[sp][tid] : spill[sp - TR_STACK];
        if (__uint_as_float(e.y) <= tcur) { cur = e.x; state = (e.x & VKHRT_BVH_LEAF) ? ST_LEAF : ST_NODE; }
    };
    // reportIntersectionEXT interval [tMin, tCurrent] + deterministic tie rule (smaller primitive id)
    auto commit = [&](float t, float u, uint32_t prim, uint32_t pos) {
        if (t >= tmin && (t < tcur || (t == tcur && prim < best_prim))) { tcur = t; best_prim = prim; best_pos = pos; best_u = u; }
    };

    for (;;) {
        // ---------------- scheduler ----------------
        const unsigned mN = __ballot_sync(FULL, state <= ST_POP);
        const unsigned mL = __ballot_sync(FULL, state == ST_LEAF);
        const unsigned mM = PH ? __ballot_sync(FULL, state == ST_MARCH) : 0u;
        const unsigned mR = __ballot_sync(FULL, state == ST_REFILL);
        const uint32_t nN = __popc(mN), nL = __popc(mL), nM = __popc(mM), nR = __popc(mR);
        if ((mN | mL | mM | mR) == 0u) break;
        uint32_t pick;
        if (nR > 0u && (nR >= p.refill_threshold || (mN | mL | mM) == 0u)) pick = ST_REFILL;
        else {
            const uint32_t sN = nN * p.w_node, sL = nL * p.w_leaf, sM = nM * p.w_march;
            pick = (sN >= sL && sN >= sM) ? ST_NODE : (sM >= sL ? ST_MARCH : ST_LEAF);
        }

        if (pick == ST_NODE) {
            // ---------------- internal nodes ----------------
            uint32_t n0 = nN, n1;
            do {
                if (STATS) { sc_steps[0]++; sc_lanes[0] += __popc(__ballot_sync(FULL, state <= ST_POP)); }
                if (state == ST_POP) pop_one();
                if (state == ST_NODE) {
                    const float4* nd = p.nodes + 4 * (size_t)cur;
                    const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3);
                    if (STATS) st_nodes++;
                    float tn0, tn1;
                    const bool h0 = slab_test(xyz(q0), xyz(q1), id, noid, tmin, tcur, &tn0);
                    const bool h1 = slab_test(xyz(q2), xyz(q3), id, noid, tmin, tcur, &tn1);
                    const uint32_t c0 = __float_as_uint(q0.w), c1 = __float_as_uint(q1.w);
                    // nearest-first: descend into the nearer hit child, the other one (if hit) goes to the stack
                    const bool both = h0 && h1;
                    const bool second = both ? (tn1 < tn0) : h1;
                    const uint32_t near_ref = second ? c1 : c0, far_ref = second ? c0 : c1;
                    if (both) push(far_ref, second ? tn0 : tn1);
                    if (h0 || h1) { cur = near_ref; if (near_ref & VKHRT_BVH_LEAF) state = ST_LEAF; }
                    else state = ST_POP;
                }
                n1 = __popc(__ballot_sync(FULL, state <= ST_POP));
            } while (n1 * 4u >= n0 * 3u && n1 > 0u);
        } else if (pick == ST_LEAF) {
            // ---------------- leaves ----------------
            if (STATS) { sc_steps[1]++; sc_lanes[1] += nL; }
            if (state == ST_LEAF) {
                const uint32_t pos = cur & 0x7FFFFFFFu;
                if (PH) {
                    // Prhi early-out (hair_intersection.rint:20-33) with the precomputed rmax
                    const float4 a0 = __ldg(p.primA + 2 * (size_t)pos), a1 = __ldg(p.primA + 2 * (size_t)pos + 1);
                    // mailbox: another piece of the curve tested last gives the same answer; skipped and not counted
                    const bool again = VKHRT_MAILBOX_PHANTOM && __float_as_uint(a1.w) == last_group;
                    if (VKHRT_MAILBOX_PHANTOM) last_group = __float_as_uint(a1.w);
                    if (STATS && !again) st_prims++;
                    if (!again && ray_hits_cylinder(o, d, xyz(a0), xyz(a1), a0.w)) {
                        const float4 b0 = __ldg(p.primB + 2 * (size_t)pos), b1 = __ldg(p.primB + 2 * (size_t)pos + 1);
                        Bezier w;
                        w.p0 = xyz(a0); w.p1 = xyz(b0); w.p2 = xyz(b1); w.p3 = xyz(a1);
                        // the ray-centric frame depends on the ray only; rebuilding it per candidate that survives the
                        // cylinder test (~2 per ray) is cheaper than keeping 9 registers alive through the node loop
                        march_begin(ms, make_ray_frame(d), o, w);
                        mpos = pos;
                        float rfilter = p.radius;
                        if (TAPER) { const float2 rr = __ldg(p.primR + pos); ms.r0 = rr.x; ms.dr = rr.y - rr.x; rfilter = fmaxf(rr.x, rr.y); }
                        // conservative filter (hair_math.cuh): skip marches that cannot report a hit
                        state = quarter_chords_near_ray(ms.c, rfilter, b0.w) ? ST_MARCH : ST_POP;
                    } else state = ST_POP;
                } else if (TECH == VKHRT_TECHNIQUE_LSS) {
                    const float4 a0 = __ldg(p.primA + 2 * (size_t)pos), a1 = __ldg(p.primA + 2 * (size_t)pos + 1);
                    float t, u;
                    if (STATS) st_prims++;
                    if (lss_intersect<false>(o, d, xyz(a0), a0.w, xyz(a1), a1.w, &t, &u, nullptr)) commit(t, u, __ldg(p.sorted_ids + pos) / VKHRT_LEAF_SPLIT_LSS, pos);
                    state = (ANYHIT && best_prim != PRIM_NONE) ? ST_REFILL : ST_POP;
                } else {
                    // one strip = the 4 triangles of a segment (64-byte record); the cheap axis-distance reject first
                    const float4* rec = p.primA + 4 * (size_t)pos;
                    const float4 a0 = __ldg(rec), a1 = __ldg(rec + 1);
                    if (STATS) st_prims += 4;
                    // per-vertex radius: the record's last two float4 are {v0 (unit), r0} {v1 (unit), r1}
                    const float4 a2 = __ldg(rec + 2), a3 = __ldg(rec + 3);
                    if (ray_near_strip_axis(o, d, xyz(a0), xyz(a1), TAPER ? fmaxf(a2.w, a3.w) : p.radius)) {
                        const uint32_t prim0 = __float_as_uint(a0.w) << 2;
#pragma unroll 1
                        for (uint32_t k = 0; k < 4u; ++k) {
                            float3 v0, v1, v2;
                            const float3 fv = (k & 2u) ? xyz(a3) : xyz(a2);
                            if (TAPER) strip_triangle_taper(xyz(a0), xyz(a1), fv * a2.w, fv * a3.w, k & 1u, &v0, &v1, &v2);
                            else strip_triangle(xyz(a0), xyz(a1), fv, k & 1u, &v0, &v1, &v2);
                            float t, u;
                            if (tri_intersect(o, d, v0, v1, v2, k & 1u, &t, &u)) commit(t, u, prim0 + k, pos);
                            if (ANYHIT && best_prim != PRIM_NONE) break;
                        }
                    }
                    state = (ANYHIT && best_prim != PRIM_NONE) ? ST_REFILL : ST_POP;
                }
            }
        } else if (PH && pick == ST_MARCH) {
            // ---------------- Phantom cone iterations ----------------
            uint32_t n0 = nM, n1;
            do {
                if (STATS) { sc_steps[2]++; sc_lanes[2] += __popc(__ballot_sync(FULL, state == ST_MARCH)); }
                if (state == ST_MARCH) {
                    if (STATS) st_iters++;
                    float t = 0.0f, u = 0.0f;
                    const int r = march_step<TAPER>(ms, p.radius, &t, &u);
                    if (r != MARCH_CONTINUE) {
                        // hair_intersection.rint:146-148: report only tHit > 0
                        if (r == MARCH_HIT && t > 0.0f) commit(t, u, __float_as_uint(__ldg(p.primA + 2 * (size_t)mpos + 1).w), mpos);
                        state = (ANYHIT && best_prim != PRIM_NONE) ? ST_REFILL : ST_POP;
                    }
                }
                n1 = __popc(__ballot_sync(FULL, state == ST_MARCH));
            } while (n1 * 4u >= n0 * 3u && n1 > 0u);
        } else {
            // ---------------- retire finished rays, refill ----------------
            const bool want = state == ST_REFILL;
            if (STATS) { sc_steps[3]++; sc_lanes[3] += nR; }
            if (want && have_ray) {
                have_ray = false;
                if (SRC == SRC_AO) {
                    // one owner per pixel and pass: a plain read-modify-write is enough
                    // one owner per pixel and pass; with several passes in one launch (n_batch > 1) a pixel's rays are in flight together
                    if (best_prim != PRIM_NONE) { if (p.n_batch > 1u) atomicAdd(p.ao_occluded + out_idx, 1u); else p.ao_occluded[out_idx] += 1u; if (STATS) st_hits++; }
                } else if (best_prim != PRIM_NONE) {
                    float3 n;
                    uint32_t seg = best_prim;
                    if (PH) {
                        // hair_intersection.rint:74-76 from the committed (t, u)
                        const float4 a0 = __ldg(p.primA + 2 * (size_t)best_pos), a1 = __ldg(p.primA + 2 * (size_t)best_pos + 1);
                        const float4 b0 = __ldg(p.primB + 2 * (size_t)best_pos), b1 = __ldg(p.primB + 2 * (size_t)best_pos + 1);
                        Bezier w;
                        w.p0 = xyz(a0); w.p1 = xyz(b0); w.p2 = xyz(b1); w.p3 = xyz(a1);
                        n = fnormalize3(fmadd3(tcur, d, o) - bezier_point(w, best_u));
                    } else if (TECH == VKHRT_TECHNIQUE_LSS) {
                        const float4 a0 = __ldg(p.primA + 2 * (size_t)best_pos), a1 = __ldg(p.primA + 2 * (size_t)best_pos + 1);
                        float t, u;
                        lss_intersect<true>(o, d, xyz(a0), a0.w, xyz(a1), a1.w, &t, &u, &n);
                    } else {
                        const float4* rec = p.primA + 4 * (size_t)best_pos;
                        const float4 a0 = __ldg(rec), a1 = __ldg(rec + 1), a2 = __ldg(rec + ((best_prim & 2u) ? 3 : 2));
                        float3 v0, v1, v2;
                        if (TAPER) {
                            const float r0 = __ldg(rec + 2).w, r1 = __ldg(rec + 3).w;
                            strip_triangle_taper(xyz(a0), xyz(a1), xyz(a2) * r0, xyz(a2) * r1, best_prim & 1u, &v0, &v1, &v2);
                        } else strip_triangle(xyz(a0), xyz(a1), xyz(a2), best_prim & 1u, &v0, &v1, &v2);
                        n = tri_normal(d, v0, v1, v2);
                        seg = best_prim >> 2;
                    }
                    store_hit(p, out_idx, tcur, seg, best_u, n, best_prim, FLAG_HIT);
                    if (STATS) st_hits++;
                } else {
                    store_hit(p, out_idx, __int_as_float(0x7f800000), VKHRT_MISS_SEGMENT, 0.0f, f3(0, 0, 0), PRIM_NONE, 0u);
                }
            }
            // warp-aggregated fetch of the next slots
            const unsigned idle = __ballot_sync(FULL, want);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(p.work, (unsigned long long)__popc(idle));
            base = __shfl_sync(FULL, base, 0);
            if (want) {
                const unsigned long long slot64 = p.slot_begin + base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (slot64 >= (unsigned long long)p.n_slots) state = ST_DONE;
                else {
                    const uint32_t slot = (uint32_t)slot64;
                    bool valid = true;
                    if (WAVEFRONT) {
                        const float4 r0 = __ldg(p.rays + 2 * (size_t)slot), r1 = __ldg(p.rays + 2 * (size_t)slot + 1);
                        o = f3(r0.x, r0.y, r0.z); tmin = r0.w; d = f3(r1.x, r1.y, r1.z); tcur = r1.w;
                        out_idx = slot;
                    } else if (SRC == SRC_AO) {
                        // several AO passes per launch: slot = pass * slots_per_sample + pixel slot
                        uint32_t a_idx = p.ao_index, ls = slot;
                        if (p.n_batch > 1u) { const uint32_t b = slot / p.slots_per_sample; ls = slot - b * p.slots_per_sample; a_idx += b; }
                        const PixelRef q = slot_to_pixel(p, ls);
                        valid = q.valid;
                        out_idx = q.out;
                        if (valid) {
                            const float4* h = reinterpret_cast<const float4*>(p.ao_hits + q.out);
                            const float4 h0 = __ldg(h), h1 = __ldg(h + 1);
                            valid = (__float_as_uint(h1.w) & FLAG_HIT) != 0u;     // a miss pixel spawns nothing
                            if (valid) {
                                float3 po, pd;
                                primary_ray(p.cam, p.W, p.H, q.px, q.py, p.sx, p.sy, &po, &pd);
                                const float3 n = f3(h0.w, h1.x, h1.y);
                                o = fmadd3(p.ao_bias, n, fmadd3(h0.x, pd, po));
                                d = ao_direction(n, q.py * p.W + q.px, p.ao_sample, a_idx);
                                tmin = VKHRT_AO_T_MIN; tcur = p.ao_distance;
                            }
                        }
                    } else {
                        uint32_t b = 0u, ls = slot;
                        if (p.n_batch > 1u) { b = slot / p.slots_per_sample; ls = slot - b * p.slots_per_sample; }
                        const PixelRef q = slot_to_pixel(p, ls);
                        valid = q.valid;
                        out_idx = q.out + b * p.n_out;
                        if (valid) {
                            primary_ray(p.cam, p.W, p.H, q.px, q.py, p.n_batch > 1u ? p.bsx[b] : p.sx, p.n_batch > 1u ? p.bsy[b] : p.sy, &o, &d);
                            tmin = p.tmin; tcur = p.tmax;
                        } else if (p.compact) {
                            store_hit(p, out_idx, __int_as_float(0x7f800000), VKHRT_MISS_SEGMENT, 0.0f, f3(0, 0, 0), PRIM_NONE, FLAG_PADDING);
                        }
                    }
                    if (valid) {
                        id = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
                        noid = f3(-(o.x * id.x), -(o.y * id.y), -(o.z * id.z));
                        best_prim = PRIM_NONE; best_pos = 0; best_u = 0.0f;
                        last_group = PRIM_NONE;
                        sp = 0;
                        have_ray = true;
                        if (STATS) st_rays++;
                        if (p.n_prims) { cur = 0u; state = ST_NODE; }   // else: stays in REFILL and retires as a miss
                    }
                }
            }
        }
    }

    if (STATS) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) {
            st_nodes += __shfl_xor_sync(FULL, st_nodes, k); st_prims += __shfl_xor_sync(FULL, st_prims, k);
            st_iters += __shfl_xor_sync(FULL, st_iters, k); st_hits += __shfl_xor_sync(FULL, st_hits, k);
            st_rays += __shfl_xor_sync(FULL, st_rays, k);
        }
        if (lane == 0) {
            atomicAdd(p.counters + 1, (unsigned long long)st_nodes); atomicAdd(p.counters + 2, (unsigned long long)st_prims);
            atomicAdd(p.counters + 3, (unsigned long long)st_hits); atomicAdd(p.counters + 4, (unsigned long long)st_iters);
            atomicAdd(p.counters + 5, (unsigned long long)st_rays);
            for (int k = 0; k < 4; ++k) { atomicAdd(p.counters + 8 + k, (unsigned long long)sc_steps[k]); atomicAdd(p.counters + 12 + k, (unsigned long long)sc_lanes[k]); }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Phantom primary rays, closest hit: per-warp RAY POOL.
//
// trace_kernel ties a ray to a lane for its whole life, so a lane whose ray waits for a leaf test or a cone iteration is
// dead weight in every node step (17 of 32 lanes active), and the leaf / march steps run with ~10 / ~9 lanes.  Here a
// warp owns S ray SLOTS in shared memory (direction, 1/d, tcur, best hit, a window of the stack) and rays move between
// queues of slot ids instead of waiting inside a lane:
//   READY  rays that want node steps     -> lanes pull them whenever they are free ("top-up"), traverse until the ray
//                                           reaches a leaf (-> LEAF) or its stack runs empty (-> DONE)
//   LEAF   rays standing at a leaf       -> batch of up to 32: the Prhi bounding-cylinder early-out; reject -> next stack entry, pass -> CAND
//   CAND   rays with a candidate curve   -> set-up batch (ray-centric transform + quarter-chord filter) into the lanes'
//                                           MARCH registers; reject -> next stack entry.  A lane's march persists across phases and is a
//                                           different ray from the one the lane traverses; cone iterations run as a phase of
//                                           their own while enough lanes hold one; a finished march commits into its ray's slot
//                                           (one outstanding candidate per ray: no race) -> next stack entry
//   DONE   finished rays                 -> batch: hit records written, slots refilled with new primary rays -> READY
// Nothing is speculated: a ray's own sequence of node visits and candidate tests is exactly trace_kernel's (and the
// oracle's), so hit records AND traversal counters are identical; only which lane executes which step changes.
// Everything is warp-synchronous (queues are per warp), so there is no inter-warp protocol to get wrong.
// The queues are LIFO stacks of slot ids: the most recently queued rays are served first, while the nodes and curves they
// touched are still in L1 (a FIFO ring measured 4 % slower on C2); every queue drains when the warp runs out of other work.
// The kernel is templated on the technique: for LSS / DOTS the LEAF batch runs the whole primitive test and the CAND / MARCH stages
// compile away.  Parity holds, but those two measure no faster than trace_kernel (profiles/experiments/r02_pool_kernel.txt §8), so only
// Phantom frames are dispatched here by default (VKHRT_POOL_LSS / VKHRT_POOL_DOTS opt in).
// ------------------------------------------------------------------------------------------------
constexpr int PL_OVF = 96;             // stack entries per slot in the global spill area (Karras depth <= 64 + 32)
constexpr int PL_MINB = 8;             // CTAs per SM the register allocation is held to
enum : int { Q_READY = 0, Q_LEAF = 1, Q_CAND = 2, Q_DONE = 3, Q_FREE = 4 };

// ------------------------------------------------------------------------------------------------
// trace_pool_kernel.  Round 2 rewrote the node step of round 1's kernel (same queues, same scheduler, same per-ray sequence of
// node visits and candidate tests: records AND counters identical); what changed is where the instructions go (ncu source
// page of round 1: stack push / pop / spill 15.8 % of the warp instructions at 5-7 lanes, loop control and state
// bookkeeping 8 %):
//   * a lane's state lives in `cur` alone: internal node index | leaf ref (bit 31) | POP | DONE | IDLE;
//   * WRITE-THROUGH stack: every push stores the entry both in the slot's shared-memory ring (the top PL_STK entries) and at
//     its depth in the warp's global spill area ([depth][slot]: lanes of a warp at similar depths share lines, .cg = L2
//     only).  A push therefore never branches on "ring full"; a pop reads the ring when the entry is still there
//     (index >= lo) and the spill area otherwise.  `lo` = lowest index whose ring copy is intact;
//   * a ray whose leaf test / candidate set-up / march is over pops its next stack entry RIGHT THERE, in the batch that
//     finished it (all lanes of the batch do the same thing), and goes straight back to the LEAF queue when that entry is a
//     leaf — consecutive leaves of a ray (pieces of neighbouring curves) no longer pass through a node lane in between.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t R2_POP = 0x7FFFFFFDu, R2_DONE = 0x7FFFFFFEu, R2_IDLE = 0x7FFFFFFFu;   // cur < R2_POP: internal node

template <int PL_S, int PL_STK, bool ORG = false>
struct PoolWarp {
    uint2 stack[PL_STK][PL_S];         // ring: entry k of the stack sits at [k % PL_STK] while index k >= lo
    float dir[6][PL_S];                // d.xyz, 1/d
    float org[ORG ? 3 : 1][ORG ? PL_S : 1];   // ray origin, only for rays that do not start at the camera (ambient-occlusion rays)
    float tcur[PL_S];
    uint32_t cur[PL_S], best_pos[PL_S];
    uint32_t last_group[PL_S];         // Phantom: the curve this ray tested last (one-entry mailbox); LSS / DOTS: primitive id of the best hit
    float best_u[PL_S];
    uint32_t out_idx[PL_S];
    uint8_t sp[PL_S], lo[PL_S];
    uint8_t q[5][PL_S];
    uint2* ovf;                        // this warp's spill area [depth][slot] (kept here: recomputing it from the CTA / warp id in
                                       // every push costs more than one LDS)
};

// AO = the rays are ambient-occlusion rays spawned from the primary hit records (SRC_AO of trace_kernel: origin per slot, first accepted
// hit ends the ray, result = the pixel's occlusion count); Phantom only.
template <int TECH, bool STATS, int PL_S, int PL_STK, int MINB = PL_MINB, bool TAPER = false, bool AO = false>
__global__ void __launch_bounds__(TR_BLOCK, MINB) trace_pool_kernel(const TraceParams p)
{
    constexpr bool PH = TECH == VKHRT_TECHNIQUE_PHANTOM;        // LSS / DOTS: the leaf batch runs the whole primitive test; no CAND / MARCH
    static_assert(!AO || PH, "the ambient-occlusion variant of the pool kernel is Phantom only");
    __shared__ PoolWarp<PL_S, PL_STK, AO> sh_all[TR_BLOCK / 32];
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.work_next = 0ull;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    PoolWarp<PL_S, PL_STK, AO>& sh = sh_all[warp];
    // spill area of this warp: [depth][slot]
    if (lane == 0) sh.ovf = p.pool_overflow + (size_t)(blockIdx.x * (TR_BLOCK / 32) + (uint32_t)warp) * (size_t)(PL_S * PL_OVF);
    const float3 cam_o = f3(p.cam.vi[12], p.cam.vi[13], p.cam.vi[14]);     // ray_gen.rgen:22: every primary ray starts at the camera
    auto org = [&](uint32_t s) -> float3 { return AO ? f3(sh.org[0][s], sh.org[1][s], sh.org[2][s]) : cam_o; };
    const float ray_tmin = AO ? VKHRT_AO_T_MIN : p.tmin;

    // the ray this lane traverses (cur == R2_IDLE: none)
    uint32_t slot = 0, cur = R2_IDLE;
    uint32_t sp = 0, lo = 0;
    float3 id = f3(0, 0, 0), noid = f3(0, 0, 0);
    float tcur = 0.0f;
    // the candidate this lane marches (a different ray)
    bool mhave = false;
    MarchState ms;
    ms.c.p0 = ms.c.p1 = ms.c.p2 = ms.c.p3 = f3(0, 0, 0);
    ms.t = ms.told = ms.dt1 = ms.dt2 = ms.t_start = 0.0f; ms.it = 0u; ms.r0 = ms.dr = 0.0f;
    uint32_t m_slot = 0;
    uint32_t nR = 0, nL = 0, nC = 0, nD = 0, nF = PL_S;      // queue fills (warp-uniform)
    bool exhausted = false;
    uint32_t st_nodes = 0, st_prims = 0, st_iters = 0, st_hits = 0, st_rays = 0;
    uint32_t sc_steps[4] = {0, 0, 0, 0}, sc_lanes[4] = {0, 0, 0, 0};

    for (int k = lane; k < PL_S; k += 32) sh.q[Q_FREE][k] = (uint8_t)k;
    __syncwarp();

    auto enqueue = [&](int qi, uint32_t& n, bool pred, uint32_t s) {
        const unsigned m = __ballot_sync(FULL, pred);
        if (pred) sh.q[qi][n + __popc(m & lt)] = (uint8_t)s;
        n += __popc(m);
    };
    auto dequeue = [&](int qi, uint32_t n, uint32_t k) -> uint32_t { return sh.q[qi][n - 1u - k]; };   // LIFO (see trace_pool_kernel)
    // Write one hit record per calling lane (`wrote`).  With line-wise host delivery the record goes to HBM, the lane counts it
    // on its 128-byte line with a releasing atomic, and for every line that just became complete four lanes copy its four
    // records to the host in one 256-bit store instruction = one 128-byte PCIe write (random 32-byte writes reach 12 GB/s,
    // whole lines from adjacent lanes 44 GB/s: tools/micro/pcie_write.cu).  Called by all lanes of the warp.
    // (Round 2 also tried counting the lines in shared memory — the four rays of a line always fetched by one warp, the
    // counting slot held until its line completes: no global atomic, no device-wide fence — and measured it SLOWER, device 1244
    // vs 1310 and e2e 1148 vs 1180 Mrays/s: refills in groups of four starve the pool and the extra state spills registers.)
    auto emit = [&](bool wrote, uint32_t oi, float t, uint32_t seg, float u, float3 n, uint32_t prim, uint32_t flags) {
        if (!p.line_cnt) { if (wrote) store_hit<false>(p, oi, t, seg, u, n, prim, flags); return; }      // warp-uniform
        bool completes = false;
        uint32_t line = 0;
        if (wrote) {
            store_hit<false>(p, oi, t, seg, u, n, prim, flags);
            line = oi >> p.line_shift;
            uint32_t old;
#if VKHRT_LINE_PROTOCOL == 0
            __threadfence();
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.line_cnt + line) : "memory");
#else
            // RELEASE only: the count publishes this lane's record (one MEMBAR, no L1 invalidation).  Whoever sees the line's last
            // count reads the records with ld.cg — L2 is the coherence point, so there is nothing in L1 to invalidate on that side.
            asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.line_cnt + line) : "memory");
#endif
            completes = old + 1u == min(1u << p.line_shift, p.n_out - (line << p.line_shift));
        }
        unsigned m = __ballot_sync(FULL, completes);
        while (m) {
            const uint32_t src = __fns(m, 0, (lane >> p.line_shift) + 1);
            const uint32_t ln = __shfl_sync(FULL, line, src & 31u);
            const uint32_t rec = (ln << p.line_shift) + ((uint32_t)lane & ((1u << p.line_shift) - 1u));
            if (src != 0xFFFFFFFFu && rec < p.n_out) {
#if VKHRT_LINE_PROTOCOL == 0
                uint32_t c;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(p.line_cnt + ln) : "memory");
#endif
                const float4* sp4 = reinterpret_cast<const float4*>(p.hits + rec);
                const float4 x = __ldcg(sp4), y = __ldcg(sp4 + 1);
                store_record(p.host_lines + rec, true, x, y);
            }
            for (uint32_t j = 0; j < (32u >> p.line_shift) && m; ++j) m &= m - 1u;
        }
    };
    // ---- stack of the ray a lane traverses (registers sp, lo) ----
    auto push = [&](uint2 e) {
        sh.stack[sp % PL_STK][slot] = e;
        __stcg(reinterpret_cast<unsigned long long*>(sh.ovf + (sp * PL_S + slot)), ((unsigned long long)e.y << 32) | e.x);
        lo = max(lo + (PL_STK - 1), sp) - (PL_STK - 1);         // the ring slot written held entry sp - PL_STK
        ++sp;
    };
    // one entry; entries whose box entry distance lies beyond the current closest hit are dropped (cur stays R2_POP: the cull
    // loop runs across steps, in parallel over lanes)
    auto pop = [&]() {
        if (sp == 0) { cur = R2_DONE; return; }
        --sp;
        uint2 e;
        if (sp >= lo) e = sh.stack[sp % PL_STK][slot];
        else { const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long*>(sh.ovf + (sp * PL_S + slot))); e = make_uint2((uint32_t)v, (uint32_t)(v >> 32)); lo = sp; }
        if (__uint_as_float(e.y) <= tcur) cur = e.x;
    };
    // the same pop on a PARKED ray (slot state in shared memory), for the batches that finish a leaf / candidate / march:
    // returns the ray's new `cur` (internal node | leaf | R2_POP when the entry was culled | R2_DONE)
    auto pop_parked = [&](uint32_t s) -> uint32_t {
        uint32_t k = sh.sp[s];
        if (k == 0u) return R2_DONE;
        --k;
        const uint32_t l = sh.lo[s];
        uint2 e;
        if (k >= l) e = sh.stack[k % PL_STK][s];
        else { const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long*>(sh.ovf + (k * PL_S + s))); e = make_uint2((uint32_t)v, (uint32_t)(v >> 32)); sh.lo[s] = (uint8_t)k; }
        sh.sp[s] = (uint8_t)k;
        return __uint_as_float(e.y) <= sh.tcur[s] ? e.x : R2_POP;
    };
    // a parked ray goes where its `cur` says: leaf -> LEAF, DONE -> DONE, node / POP -> READY
    auto route = [&](bool pred, uint32_t s, uint32_t c) {
        if (pred) sh.cur[s] = c;
        enqueue(Q_LEAF, nL, pred && (int)c < 0, s);
        enqueue(Q_DONE, nD, pred && c == R2_DONE, s);
        enqueue(Q_READY, nR, pred && (int)c >= 0 && c != R2_DONE, s);
    };

    uint32_t nHave = 0, nM = 0;               // lanes holding a ray / a march (warp-uniform)
    auto top_up = [&]() {
        if (nR == 0u || nHave == 32u) return;
        const unsigned idle = __ballot_sync(FULL, cur == R2_IDLE);
        const uint32_t rank = __popc(idle & lt);
        if (cur == R2_IDLE && rank < nR) {
            slot = dequeue(Q_READY, nR, rank);
            id = f3(sh.dir[3][slot], sh.dir[4][slot], sh.dir[5][slot]);
            const float3 o = org(slot);
            noid = f3(-(o.x * id.x), -(o.y * id.y), -(o.z * id.z));
            tcur = sh.tcur[slot]; cur = sh.cur[slot]; sp = sh.sp[slot]; lo = sh.lo[slot];
        }
        const uint32_t taken = min(nR, 32u - nHave);
        nR -= taken; nHave += taken;
        __syncwarp();
    };

    for (;;) {
        top_up();
        const uint32_t nRet = exhausted ? nD : nD + nF;
        if ((nHave | nM | nR | nL | nC | nRet) == 0u) break;

        // ---------------- scheduler (as trace_pool_kernel) ----------------
        enum : int { PH_NODE, PH_LEAF, PH_SETUP, PH_MARCH, PH_RETIRE };
        int phase;
        if (!exhausted && nF >= 32u) phase = PH_RETIRE;
        else if (nHave >= p.pool_node_lanes) phase = PH_NODE;
        else {
            const uint32_t sS = min(nC, 32u - nM);
            const uint32_t sR = min(nRet, 32u);
            uint32_t best = nL; int bp = PH_LEAF;
            if (sS > best) { best = sS; bp = PH_SETUP; }
            if (nM > best || (nM == 32u)) { best = nM; bp = PH_MARCH; }
            if (sR > best) { best = sR; bp = PH_RETIRE; }
            if (best >= p.pool_batch_lanes) phase = bp;
            else if (nHave >= p.pool_node_min) phase = PH_NODE;
            else if (best > 0u) phase = bp;
            else phase = PH_NODE;
        }

        if (phase == PH_NODE) {
            // ---------------- internal nodes ----------------
            do {
                const uint32_t thr = max(1u, (nHave * p.pool_exit_eighths + 7u) >> 3);     // leave the loop below this many node lanes
                uint32_t n1;
                do {
                    if (STATS) { sc_steps[0]++; sc_lanes[0] += __popc(__ballot_sync(FULL, cur <= R2_POP)); }
                    bool both = false;
                    uint2 far = make_uint2(0u, 0u);
                    if (cur < R2_POP) {
                        const float4* nd = p.nodes + 4 * (size_t)cur;
                        const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3);
                        if (STATS) st_nodes++;
                        float tn0, tn1;
                        const bool h0 = slab_test(xyz(q0), xyz(q1), id, noid, ray_tmin, tcur, &tn0);
                        const bool h1 = slab_test(xyz(q2), xyz(q3), id, noid, ray_tmin, tcur, &tn1);
                        const uint32_t c0 = __float_as_uint(q0.w), c1 = __float_as_uint(q1.w);
                        both = h0 && h1;
                        const bool second = both ? (tn1 < tn0) : h1;
                        far = make_uint2(second ? c0 : c1, __float_as_uint(second ? tn0 : tn1));
                        cur = (h0 || h1) ? (second ? c1 : c0) : R2_POP;
                    }
                    if (both) push(far);
                    else if (cur == R2_POP) pop();
                    n1 = __popc(__ballot_sync(FULL, cur <= R2_POP));
                } while (n1 >= thr);
                // rays that left the node state go to their queues; the lane is free again
                const bool to_leaf = (int)cur < 0, to_done = cur == R2_DONE;
                if (to_leaf) { sh.cur[slot] = cur; sh.sp[slot] = (uint8_t)sp; sh.lo[slot] = (uint8_t)lo; }
                enqueue(Q_LEAF, nL, to_leaf, slot);
                enqueue(Q_DONE, nD, to_done, slot);
                if (to_leaf || to_done) cur = R2_IDLE;
                nHave = n1;
                __syncwarp();
                top_up();
            } while (nHave >= p.pool_node_lanes);
        } else if (phase == PH_LEAF) {
            // ---------------- leaves: Phantom: Prhi early-out (hair_intersection.rint:20-33, rmax precomputed); LSS / DOTS: the primitive test ----------------
            const uint32_t n = min(nL, 32u);
            const bool act = (uint32_t)lane < n;
            if (STATS) { sc_steps[1]++; sc_lanes[1] += n; }
            uint32_t s = 0, next = 0;
            bool pass = false;
            if (act) {
                s = dequeue(Q_LEAF, nL, (uint32_t)lane);
                const float3 d = f3(sh.dir[0][s], sh.dir[1][s], sh.dir[2][s]);
                const float3 o = org(s);
                const uint32_t pos = sh.cur[s] & 0x7FFFFFFFu;
                if (PH) {
                    const float4 a0 = __ldg(p.primA + 2 * (size_t)pos), a1 = __ldg(p.primA + 2 * (size_t)pos + 1);
                    // mailbox: another piece of the curve tested last gives the same answer; skipped and not counted
                    const bool again = VKHRT_MAILBOX_PHANTOM && __float_as_uint(a1.w) == sh.last_group[s];
                    if (VKHRT_MAILBOX_PHANTOM) sh.last_group[s] = __float_as_uint(a1.w);
                    if (STATS && !again) st_prims++;
                    pass = !again && ray_hits_cylinder(o, d, xyz(a0), xyz(a1), a0.w);
                } else {
                    // reportIntersectionEXT interval [tMin, tCurrent] + tie rule (smaller primitive id), as trace_kernel's `commit`
                    auto commit = [&](float t, float u, uint32_t prim) {
                        const float tc = sh.tcur[s];
                        if (t >= ray_tmin && (t < tc || (t == tc && prim < sh.last_group[s]))) { sh.tcur[s] = t; sh.last_group[s] = prim; sh.best_pos[s] = pos; sh.best_u[s] = u; }
                    };
                    if (TECH == VKHRT_TECHNIQUE_LSS) {
                        const float4 a0 = __ldg(p.primA + 2 * (size_t)pos), a1 = __ldg(p.primA + 2 * (size_t)pos + 1);
                        float t, u;
                        if (STATS) st_prims++;
                        if (lss_intersect<false>(o, d, xyz(a0), a0.w, xyz(a1), a1.w, &t, &u, nullptr)) commit(t, u, __ldg(p.sorted_ids + pos) / VKHRT_LEAF_SPLIT_LSS);
                    } else {
                        const float4* rec = p.primA + 4 * (size_t)pos;
                        const float4 a0 = __ldg(rec), a1 = __ldg(rec + 1);
                        if (STATS) st_prims += 4;
                        if (ray_near_strip_axis(o, d, xyz(a0), xyz(a1), p.radius)) {
                            const float4 a2 = __ldg(rec + 2), a3 = __ldg(rec + 3);
                            const uint32_t prim0 = __float_as_uint(a0.w) << 2;
#pragma unroll 1
                            for (uint32_t k = 0; k < 4u; ++k) {
                                float3 v0, v1, v2;
                                strip_triangle(xyz(a0), xyz(a1), (k & 2u) ? xyz(a3) : xyz(a2), k & 1u, &v0, &v1, &v2);
                                float t, u;
                                if (tri_intersect(o, d, v0, v1, v2, k & 1u, &t, &u)) commit(t, u, prim0 + k);
                            }
                        }
                    }
                }
                if (!pass) next = pop_parked(s);
            }
            nL -= n;
            __syncwarp();
            enqueue(Q_CAND, nC, act && pass, s);
            route(act && !pass, s, next);
            __syncwarp();
        } else if (PH && phase == PH_SETUP) {
            // ---------------- candidates: ray-centric transform + quarter-chord filter into free MARCH registers ----------------
            const unsigned midle = __ballot_sync(FULL, !mhave);
            const uint32_t rank = __popc(midle & lt);
            const bool take = !mhave && rank < nC;
            if (STATS) { sc_steps[1]++; sc_lanes[1] += __popc(__ballot_sync(FULL, take)); }
            bool reject = false;
            uint32_t next = 0;
            if (take) {
                m_slot = dequeue(Q_CAND, nC, rank);
                const float3 d = f3(sh.dir[0][m_slot], sh.dir[1][m_slot], sh.dir[2][m_slot]);
                const uint32_t pos = sh.cur[m_slot] & 0x7FFFFFFFu;
                const float4 a0 = __ldg(p.primA + 2 * (size_t)pos), a1 = __ldg(p.primA + 2 * (size_t)pos + 1);
                const float4 b0 = __ldg(p.primB + 2 * (size_t)pos), b1 = __ldg(p.primB + 2 * (size_t)pos + 1);
                Bezier w;
                w.p0 = xyz(a0); w.p1 = xyz(b0); w.p2 = xyz(b1); w.p3 = xyz(a1);
                march_begin(ms, make_ray_frame(d), org(m_slot), w);
                float rfilter = p.radius;
                if (TAPER) { const float2 rr = __ldg(p.primR + pos); ms.r0 = rr.x; ms.dr = rr.y - rr.x; rfilter = fmaxf(rr.x, rr.y); }   // per-vertex radii (§4.10)
                if (quarter_chords_near_ray(ms.c, rfilter, b0.w)) mhave = true;
                else { reject = true; next = pop_parked(m_slot); }
            }
            nC -= min(nC, (uint32_t)__popc(midle));
            nM = __popc(__ballot_sync(FULL, mhave));
            __syncwarp();
            route(reject, m_slot, next);
            __syncwarp();
        } else if (PH && phase == PH_MARCH) {
            // ---------------- Phantom cone iterations (hair_intersection.rint:56-126) ----------------
            uint32_t n0 = nM, n1;
            do {
                if (STATS) { sc_steps[2]++; sc_lanes[2] += __popc(__ballot_sync(FULL, mhave)); }
                bool fin = false;
                uint32_t next = 0;
                if (mhave) {
                    if (STATS) st_iters++;
                    float t = 0.0f, u = 0.0f;
                    const int r = march_step<TAPER>(ms, p.radius, &t, &u);
                    if (r != MARCH_CONTINUE) {
                        mhave = false; fin = true;
                        const uint32_t pos = sh.cur[m_slot] & 0x7FFFFFFFu;       // the slot still points at the candidate's leaf
                        // hair_intersection.rint:146-148 (report only tHit > 0), reportIntersectionEXT interval, tie rule of `commit`
                        bool took = false;
                        if (r == MARCH_HIT && t > 0.0f && t >= ray_tmin) {
                            const float tc = sh.tcur[m_slot];
                            bool take_it = t < tc;
                            if (t == tc) {
                                const uint32_t bpos = sh.best_pos[m_slot];
                                const uint32_t prim = __float_as_uint(__ldg(p.primA + 2 * (size_t)pos + 1).w);
                                take_it = bpos == PRIM_NONE || prim < __float_as_uint(__ldg(p.primA + 2 * (size_t)bpos + 1).w);
                            }
                            if (take_it) { sh.tcur[m_slot] = t; sh.best_pos[m_slot] = pos; sh.best_u[m_slot] = u; took = true; }
                        }
                        // gl_RayFlagsTerminateOnFirstHitEXT (occlusion rays): the first accepted hit ends the ray
                        next = (AO && took) ? R2_DONE : pop_parked(m_slot);          // (the pop culls against the hit just committed)
                    }
                }
                route(fin, m_slot, next);
                n1 = __popc(__ballot_sync(FULL, mhave));
            } while (n1 * 4u >= n0 * 3u && n1 > 0u);
            nM = n1;
            __syncwarp();
        } else {
            // ---------------- retire finished rays, refill their slots (and free ones) with new primary rays ----------------
            const uint32_t n_d = min(nD, 32u);
            const uint32_t n_f = exhausted ? 0u : min(nF, 32u - n_d);
            const bool act_d = (uint32_t)lane < n_d, act_f = !act_d && (uint32_t)lane - n_d < n_f;
            if (STATS) { sc_steps[3]++; sc_lanes[3] += n_d + n_f; }
            uint32_t s = 0;
            if (act_d) s = dequeue(Q_DONE, nD, (uint32_t)lane);
            else if (act_f) s = dequeue(Q_FREE, nF, (uint32_t)lane - n_d);
            nD -= n_d; nF -= n_f;
            {
                float rt = __int_as_float(0x7f800000), ru = 0.0f;
                float3 rn = f3(0, 0, 0);
                uint32_t rprim = PRIM_NONE, rseg = VKHRT_MISS_SEGMENT, rflags = 0u, oi = 0u;
                if (AO) {
                    // one more occluded ray for the pixel; several passes of a pixel can be in flight together (n_batch > 1): integer atomic
                    if (act_d && sh.best_pos[s] != PRIM_NONE) { atomicAdd(p.ao_occluded + sh.out_idx[s], 1u); if (STATS) st_hits++; }
                } else if (act_d) {
                    const uint32_t pos = sh.best_pos[s];
                    oi = sh.out_idx[s];
                    if (pos != PRIM_NONE) {
                        rt = sh.tcur[s]; ru = sh.best_u[s];
                        const float3 d = f3(sh.dir[0][s], sh.dir[1][s], sh.dir[2][s]);
                        const float3 o = cam_o;
                        if (PH) {
                            // hair_intersection.rint:74-76 from the committed (t, u)
                            const float4 a0 = __ldg(p.primA + 2 * (size_t)pos), a1 = __ldg(p.primA + 2 * (size_t)pos + 1);
                            const float4 b0 = __ldg(p.primB + 2 * (size_t)pos), b1 = __ldg(p.primB + 2 * (size_t)pos + 1);
                            Bezier w;
                            w.p0 = xyz(a0); w.p1 = xyz(b0); w.p2 = xyz(b1); w.p3 = xyz(a1);
                            rprim = rseg = __float_as_uint(a1.w);
                            rn = fnormalize3(fmadd3(rt, d, o) - bezier_point(w, ru));
                        } else if (TECH == VKHRT_TECHNIQUE_LSS) {
                            const float4 a0 = __ldg(p.primA + 2 * (size_t)pos), a1 = __ldg(p.primA + 2 * (size_t)pos + 1);
                            float t, u;
                            lss_intersect<true>(o, d, xyz(a0), a0.w, xyz(a1), a1.w, &t, &u, &rn);
                            rprim = rseg = sh.last_group[s];
                        } else {
                            rprim = sh.last_group[s];
                            const float4* rec = p.primA + 4 * (size_t)pos;
                            const float4 a0 = __ldg(rec), a1 = __ldg(rec + 1), a2 = __ldg(rec + ((rprim & 2u) ? 3 : 2));
                            float3 v0, v1, v2;
                            strip_triangle(xyz(a0), xyz(a1), xyz(a2), rprim & 1u, &v0, &v1, &v2);
                            rn = tri_normal(d, v0, v1, v2);
                            rseg = rprim >> 2;
                        }
                        rflags = FLAG_HIT;
                        if (STATS) st_hits++;
                    }
                }
                if (!AO) emit(act_d, oi, rt, rseg, ru, rn, rprim, rflags);
            }
            const bool want = (act_d || act_f) && !exhausted;
            const unsigned wm = __ballot_sync(FULL, want);
            unsigned long long base = 0;
            if (lane == 0 && wm) base = atomicAdd(p.work, (unsigned long long)__popc(wm));
            base = __shfl_sync(FULL, base, 0);
            bool fresh = false, padded = false;
            uint32_t pad_idx = 0u;
            if (want) {
                const unsigned long long slot64 = p.slot_begin + base + (unsigned)__popc(wm & lt);
                if (AO) {
                    if (slot64 < (unsigned long long)p.n_slots) {
                        // an occlusion ray of pass a_idx from the pixel's primary hit record (the refill step of trace_kernel<.., SRC_AO>)
                        uint32_t a_idx = p.ao_index, ls = (uint32_t)slot64;
                        if (p.n_batch > 1u) { const uint32_t b = ls / p.slots_per_sample; ls -= b * p.slots_per_sample; a_idx += b; }
                        const PixelRef q = slot_to_pixel(p, ls);
                        if (q.valid) {
                            const float4* h = reinterpret_cast<const float4*>(p.ao_hits + q.out);
                            const float4 h0 = __ldg(h), h1 = __ldg(h + 1);
                            if (__float_as_uint(h1.w) & FLAG_HIT) {                 // a miss pixel spawns nothing
                                float3 po, pd;
                                primary_ray(p.cam, p.W, p.H, q.px, q.py, p.sx, p.sy, &po, &pd);
                                const float3 n = f3(h0.w, h1.x, h1.y);
                                const float3 ro = fmadd3(p.ao_bias, n, fmadd3(h0.x, pd, po));
                                const float3 d = ao_direction(n, q.py * p.W + q.px, p.ao_sample, a_idx);
                                sh.org[0][s] = ro.x; sh.org[1][s] = ro.y; sh.org[2][s] = ro.z;
                                sh.dir[0][s] = d.x; sh.dir[1][s] = d.y; sh.dir[2][s] = d.z;
                                sh.dir[3][s] = safe_rcp(d.x); sh.dir[4][s] = safe_rcp(d.y); sh.dir[5][s] = safe_rcp(d.z);
                                sh.tcur[s] = p.ao_distance; sh.cur[s] = 0u; sh.sp[s] = 0; sh.lo[s] = 0;
                                sh.best_pos[s] = PRIM_NONE; sh.best_u[s] = 0.0f; sh.out_idx[s] = q.out;
                                sh.last_group[s] = PRIM_NONE;
                                fresh = true;
                                if (STATS) st_rays++;
                            }
                        }
                    }
                } else
                if (slot64 < (unsigned long long)p.n_slots) {
                    uint32_t b = 0u, ls = (uint32_t)slot64;
                    if (p.n_batch > 1u) { b = ls / p.slots_per_sample; ls -= b * p.slots_per_sample; }
                    PixelRef q = slot_to_pixel(p, ls);
                    q.out += b * p.n_out;
                    if (q.valid) {
                        float3 ro, d;
                        primary_ray(p.cam, p.W, p.H, q.px, q.py, p.n_batch > 1u ? p.bsx[b] : p.sx, p.n_batch > 1u ? p.bsy[b] : p.sy, &ro, &d);
                        sh.dir[0][s] = d.x; sh.dir[1][s] = d.y; sh.dir[2][s] = d.z;
                        sh.dir[3][s] = safe_rcp(d.x); sh.dir[4][s] = safe_rcp(d.y); sh.dir[5][s] = safe_rcp(d.z);
                        sh.tcur[s] = p.tmax; sh.cur[s] = 0u; sh.sp[s] = 0; sh.lo[s] = 0;
                        sh.best_pos[s] = PRIM_NONE; sh.best_u[s] = 0.0f; sh.out_idx[s] = q.out;
                        sh.last_group[s] = PRIM_NONE;
                        fresh = true;
                        if (STATS) st_rays++;
                    } else if (p.compact) { padded = true; pad_idx = q.out; }
                }
            }
            if (!AO && p.compact) emit(padded, pad_idx, __int_as_float(0x7f800000), VKHRT_MISS_SEGMENT, 0.0f, f3(0, 0, 0), PRIM_NONE, FLAG_PADDING);
            if (wm && p.slot_begin + base + (unsigned)__popc(wm) >= (unsigned long long)p.n_slots) exhausted = true;
            __syncwarp();
            enqueue(Q_READY, nR, fresh, s);
            enqueue(Q_FREE, nF, (act_d || act_f) && !fresh, s);
            __syncwarp();
        }
    }

    if (STATS) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) {
            st_nodes += __shfl_xor_sync(FULL, st_nodes, k); st_prims += __shfl_xor_sync(FULL, st_prims, k);
            st_iters += __shfl_xor_sync(FULL, st_iters, k); st_hits += __shfl_xor_sync(FULL, st_hits, k);
            st_rays += __shfl_xor_sync(FULL, st_rays, k);
        }
        if (lane == 0) {
            atomicAdd(p.counters + 1, (unsigned long long)st_nodes); atomicAdd(p.counters + 2, (unsigned long long)st_prims);
            atomicAdd(p.counters + 3, (unsigned long long)st_hits); atomicAdd(p.counters + 4, (unsigned long long)st_iters);
            atomicAdd(p.counters + 5, (unsigned long long)st_rays);
            for (int k = 0; k < 4; ++k) { atomicAdd(p.counters + 8 + k, (unsigned long long)sc_steps[k]); atomicAdd(p.counters + 12 + k, (unsigned long long)sc_lanes[k]); }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// wavefront ray generator: one 32-byte record per ray {o, tmin, d, tmax}, in slot order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raygen_kernel(const TraceParams p, float4* __restrict__ rays)
{
    unsigned long long slot64 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot64 >= p.n_slots) return;
    PixelRef q = slot_to_pixel(p, (uint32_t)slot64);
    float3 o = f3(0, 0, 0), d = f3(0, 0, 1);
    float tmin = p.tmin, tmax = p.tmax;
    if (!q.valid && !p.compact) return;   // row-major layout has no slot for padding pixels
    if (q.valid) primary_ray(p.cam, p.W, p.H, q.px, q.py, p.sx, p.sy, &o, &d);
    else tmax = -1.0f;    // empty interval: padding pixels never hit
    rays[2 * (size_t)q.out] = make_float4(o.x, o.y, o.z, tmin);
    rays[2 * (size_t)q.out + 1] = make_float4(d.x, d.y, d.z, tmax);
}

// ------------------------------------------------------------------------------------------------
// closest-hit colour (shading.glsl / debug.glsl) + imageStore to 8-bit UNORM; spp accumulation in fp32
// ------------------------------------------------------------------------------------------------
// One thread per ray slot of this shard (same slot -> pixel map as the traversal kernel), so that a shard only
// ever touches the pixels it owns whatever the output layout is.
__global__ void __launch_bounds__(256) shade_kernel(const TraceParams p, const VkhrtHit* __restrict__ hits, int mode, float3 miss, float3 albedo,
                                                    float4* __restrict__ accum, uchar4* __restrict__ rgba, uint32_t sample, uint32_t spp,
                                                    const uint32_t* __restrict__ occluded, uint32_t ao_samples,
                                                    const float4* __restrict__ env, uint32_t env_w, uint32_t env_h,
                                                    const float4* __restrict__ meshes, uint32_t n_meshes)
{
    const unsigned long long slot64 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot64 >= p.n_slots) return;
    const PixelRef q = slot_to_pixel(p, (uint32_t)slot64);
    const size_t i = q.out;
    if (!q.valid) {                                         // padding pixel of a partial tile
        if (p.compact && rgba && sample + 1 == spp) rgba[i] = make_uchar4(0, 0, 0, 0);
        return;
    }
    const float4* h = reinterpret_cast<const float4*>(hits + i);
    const float4 h0 = h[0], h1 = h[1];
    const uint32_t flags = __float_as_uint(h1.w);
    float3 c;
    if (flags & FLAG_HIT) {
        c = mode == VKHRT_SHADE_DEBUG_PRIMID ? debug_palette(__float_as_uint(h1.z)) : shade_normal(f3(h0.w, h1.x, h1.y));
        if (mode == VKHRT_SHADE_MATERIAL) {                                          // triangle_closest_hit.rchit:72-83
            if (n_meshes > 1u) {
                // geometryNodes[blasInstances[gl_InstanceCustomIndexEXT].firstGeometryIndex].material: the mesh that owns the hit segment
                const uint32_t seg = __float_as_uint(h0.y);
                uint32_t lo = 0u, hi = n_meshes;                                     // last mesh whose first segment <= seg
                while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (__float_as_uint(__ldg(meshes + mid).w) <= seg) lo = mid; else hi = mid; }
                const float4 m = __ldg(meshes + lo);
                albedo = f3(m.x, m.y, m.z);
            }
            c = f3(c.x * albedo.x, c.y * albedo.y, c.z * albedo.z);
        }
        if (ao_samples) c = c * (1.0f - (float)occluded[i] / (float)ao_samples);    // unoccluded fraction of the AO rays
    } else if (env) {
        // miss.rmiss: the colour depends on the ray of THIS sample, regenerated here (ray_gen.rgen:16-24)
        float3 ro, rd;
        primary_ray(p.cam, p.W, p.H, q.px, q.py, p.sx, p.sy, &ro, &rd);
        c = environment_miss(env, env_w, env_h, rd);
    } else c = miss;
    if (spp > 1) {
        float4 a = sample == 0 ? make_float4(0, 0, 0, 0) : accum[i];
        a.x += c.x; a.y += c.y; a.z += c.z;
        if (sample + 1 < spp) { accum[i] = a; return; }
        float fs = (float)spp;
        c = f3(a.x / fs, a.y / fs, a.z / fs);
    }
    if (rgba) rgba[i] = make_uchar4((unsigned char)to_unorm8(c.x), (unsigned char)to_unorm8(c.y), (unsigned char)to_unorm8(c.z), 255);
}

// gathered compact shards (rank-major) -> row-major image; elem = 4 or 32 bytes
struct alignas(16) Elem32 { uint4 a, b; };
template <typename E>
__global__ void __launch_bounds__(256) untile_kernel(const E* __restrict__ src, E* __restrict__ dst, uint32_t W, uint32_t H, uint32_t T,
                                                     uint32_t tiles_x, uint32_t world, unsigned long long shard_elems)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned long long)W * H) return;
    uint32_t px = (uint32_t)(i % W), py = (uint32_t)(i / W);
    uint32_t tile = (py / T) * tiles_x + px / T;
    uint32_t rank = tile % world, tl = tile / world;
    dst[i] = src[rank * shard_elems + (unsigned long long)tl * T * T + (unsigned long long)(py % T) * T + (px % T)];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Resolved {
    uint32_t W, H, spp, T, tiles_x, tiles_y, n_tiles, tile_first, tile_stride, n_local_tiles;
    bool compact;
    float tmin, tmax;
    unsigned long long n_slots, n_out;
};
static bool resolve(const VkhrtFrameDesc& f, Resolved& r)
{
    r.W = f.width; r.H = f.height;
    if (r.W == 0 || r.H == 0) return false;
    r.spp = f.spp ? f.spp : 1;
    r.T = f.tile_size ? f.tile_size : 64;
    if (r.T % 8 != 0 || r.T > 1024) return false;
    r.tile_stride = f.tile_stride ? f.tile_stride : 1;
    r.tile_first = f.tile_first;
    if (r.tile_first >= r.tile_stride) return false;
    r.tiles_x = (r.W + r.T - 1) / r.T; r.tiles_y = (r.H + r.T - 1) / r.T;
    r.n_tiles = r.tiles_x * r.tiles_y;
    r.compact = r.tile_stride > 1 && !f.row_major_output;
    r.n_local_tiles = (r.n_tiles + r.tile_stride - 1) / r.tile_stride;
    r.n_slots = (unsigned long long)r.n_local_tiles * r.T * r.T;
    r.n_out = r.compact ? r.n_slots : (unsigned long long)r.W * r.H;
    if (r.n_slots >= 0xFFFFFFFFull || (unsigned long long)r.W * r.H >= 0xFFFFFFFFull) return false;   // 32-bit slot / pixel indices
    bool def = f.t_min == 0.0f && f.t_max == 0.0f;
    r.tmin = def ? VKHRT_DEFAULT_T_MIN : f.t_min;
    r.tmax = def ? VKHRT_DEFAULT_T_MAX : f.t_max;
    return true;
}

uint64_t frame_local_pixels(const VkhrtFrameDesc& f)
{
    Resolved r;
    if (!resolve(f, r)) return 0;
    return r.n_out;
}

static void sample_offset(uint32_t s, float* sx, float* sy)
{
    if (s == 0) { *sx = 0.5f; *sy = 0.5f; return; }
    const double a1 = 0.7548776662466927, a2 = 0.5698402909980532;   // R2 sequence
    double x = 0.5 + a1 * (double)s, y = 0.5 + a2 * (double)s;
    *sx = (float)(x - std::floor(x)); *sy = (float)(y - std::floor(y));
}

static void fill_params(const DeviceScene& sc, const VkhrtFrameDesc& f, const Resolved& r, TraceParams& p)
{
    p.nodes = sc.d_nodes; p.primA = sc.d_primA; p.primB = sc.d_primB; p.primR = sc.d_primR; p.sorted_ids = sc.d_sorted_ids;
    p.n_prims = sc.n_leaves; p.radius = sc.radius;
    memcpy(p.cam.vi, f.view_inverse, sizeof(p.cam.vi)); memcpy(p.cam.pi, f.proj_inverse, sizeof(p.cam.pi));
    p.W = r.W; p.H = r.H; p.sx = 0.5f; p.sy = 0.5f; p.tmin = r.tmin; p.tmax = r.tmax;
    p.T = r.T; p.tiles_x = r.tiles_x; p.n_tiles = r.n_tiles; p.tile_first = r.tile_first; p.tile_stride = r.tile_stride; p.compact = r.compact ? 1u : 0u;
    p.rays = nullptr; p.slot_begin = 0; p.n_slots = (uint32_t)r.n_slots; p.hits = nullptr; p.hits_mirror = nullptr; p.counters = sc.d_counters;
    p.host_dest = 0u; p.hits_aligned32 = 0u; p.pool_overflow = nullptr; p.line_cnt = nullptr; p.host_lines = nullptr; p.n_out = (uint32_t)r.n_out; p.n_batch = 1u; p.slots_per_sample = (uint32_t)r.n_slots; p.line_shift = 2u;
    p.ao_hits = nullptr; p.ao_occluded = nullptr; p.ao_index = p.ao_sample = 0u; p.ao_distance = 0.0f; p.ao_bias = 0.0f;
}

// Experiment switches: environment variables read ONCE per process, on first use (thread-safe static
// initialisation); the defaults are the shipped configuration.
struct Tunables {
    int refill_threshold, min_blocks, blocks_per_sm, w_node, w_leaf, w_march;
    int pool, pool_stats, pool_min_ratio, pool_node_lanes, pool_batch_lanes, pool_node_min, pool_exit, pool_cfg, pool_host, carveout;
    int store256, zero_copy, linewise, line_shift, sample_batch, pool_lss, pool_dots, early_copy, pool_taper, pool_ao, batch_mrays, batch_max;
};
static float bits_to_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static int env_int(const char* name, int def) { const char* v = getenv(name); return v ? atoi(v) : def; }
static const Tunables& tun()
{
    static const Tunables t = [] {
        Tunables x;
        x.refill_threshold = std::max(1, env_int("VKHRT_REFILL_THRESHOLD", 16));
        x.min_blocks = env_int("VKHRT_MIN_BLOCKS", TR_MIN_BLOCKS);
        x.blocks_per_sm = env_int("VKHRT_BLOCKS_PER_SM", 0);
        x.w_node = env_int("VKHRT_W_NODE", 16);
        x.w_leaf = env_int("VKHRT_W_LEAF", 32);
        x.w_march = env_int("VKHRT_W_MARCH", 32);
        x.pool = env_int("VKHRT_POOL", 1);                       // Phantom primary rays: trace_pool_kernel
        x.pool_stats = env_int("VKHRT_POOL_STATS", 1);           // scheduler statistics of the pool kernel instead of trace_kernel's
        x.pool_min_ratio = env_int("VKHRT_POOL_MIN_RATIO", 3);
        x.pool_node_lanes = env_int("VKHRT_POOL_NODE_LANES", 24);
        x.pool_batch_lanes = env_int("VKHRT_POOL_BATCH_LANES", 8);
        x.pool_node_min = env_int("VKHRT_POOL_NODE_MIN", 8);
        x.pool_exit = env_int("VKHRT_POOL_EXIT", 6);
        x.pool_cfg = env_int("VKHRT_POOL_CFG", 0);
        // LSS / DOTS primary rays through the pool kernel as well: records and counters identical, but measured no better (C3 1987 vs
        // 2234 Mrays/s: 18 % fewer warp instructions, yet L1 hits 44 vs 59 % with 4608 instead of 2048 rays in flight per SM and
        // long-scoreboard stalls double; C4 1059 vs 1040, e2e 996 vs 1020), so both default to the lane-bound kernel
        x.pool_lss = env_int("VKHRT_POOL_LSS", 0);
        x.pool_dots = env_int("VKHRT_POOL_DOTS", 0);
        x.pool_ao = env_int("VKHRT_POOL_AO", 1);                // ambient-occlusion passes of Phantom scenes through the pool kernel
        x.pool_taper = env_int("VKHRT_POOL_TAPER", 1);          // Phantom scenes with per-vertex radii through the pool kernel too
        x.pool_host = env_int("VKHRT_POOL_HOST", 0);
        x.carveout = env_int("VKHRT_CARVEOUT", -1);
        x.store256 = env_int("VKHRT_STORE256", 1);
        x.zero_copy = env_int("VKHRT_ZERO_COPY", 1);
        x.linewise = env_int("VKHRT_LINEWISE", 1);
        x.early_copy = env_int("VKHRT_EARLY_COPY", 1);       // multi-sample frames: sample-0 records leave on the copy engine under the other samples
        x.sample_batch = env_int("VKHRT_SAMPLE_BATCH", 1);     // 0 = one launch per sample (round 1)
        // samples per launch: enough for this many million rays, at most this many (<= 16).  A 1/8 shard of the 4K x 64 spp frame
        // (1.04 M rays per sample) on one GPU: 6 M / 8 -> 42.1 ms, 12 M / 16 -> 40.5 ms, 20 M / 16 -> 39.8 ms (same image)
        x.batch_mrays = env_int("VKHRT_BATCH_MRAYS", 20);
        x.batch_max = env_int("VKHRT_BATCH_MAX", 16);
        // 128-byte lines (4 records) measured best: e2e 996 (64 B) / 1077 (128 B) / 1068 (256 B) / 1034 (512 B) Mrays/s on C2
        x.line_shift = std::min(5, std::max(1, env_int("VKHRT_LINE_SHIFT", 2)));
        return x;
    }();
    return t;
}
static void tunables(TraceParams& p)
{
    const Tunables& t = tun();
    p.pool_node_lanes = (uint32_t)t.pool_node_lanes; p.pool_batch_lanes = (uint32_t)t.pool_batch_lanes; p.pool_node_min = (uint32_t)t.pool_node_min; p.pool_exit_eighths = (uint32_t)t.pool_exit;
    p.refill_threshold = (uint32_t)t.refill_threshold;
    p.w_node = (uint32_t)t.w_node; p.w_leaf = (uint32_t)t.w_leaf; p.w_march = (uint32_t)t.w_march;
}

void init_tunables() { (void)tun(); }

// resident CTAs per SM of a kernel: asked once per kernel instantiation and device
template <typename K>
static int blocks_per_sm(K kernel, int device)
{
    static int cached[64];            // 0 = not asked yet; indexed by device ordinal
    const int d = device & 63;
    if (cached[d] == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TR_BLOCK, 0) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
        cached[d] = std::max(1, per_sm);
    }
    return cached[d];
}

// every traversal launch takes the next of the two alternating work counters (TraceParams::work)
static void take_work_counter(DeviceScene& sc, TraceParams& p)
{
    p.work = sc.d_counters + (sc.work_flip ? 6 : 0);
    p.work_next = sc.d_counters + (sc.work_flip ? 0 : 6);
    sc.work_flip = !sc.work_flip;
}

template <int TECH, bool STATS, int SRC, bool ANYHIT, int MINB, bool TAPER = false>
static int launch_trace_t(DeviceScene& sc, TraceParams& p, cudaStream_t st)
{
    int per_sm = blocks_per_sm(trace_kernel<TECH, STATS, SRC, ANYHIT, MINB, TAPER>, sc.device);
    tunables(p);
    if (tun().blocks_per_sm > 0) per_sm = std::min(per_sm, tun().blocks_per_sm);
    unsigned long long want = ((unsigned long long)(p.n_slots - p.slot_begin) + TR_BLOCK - 1) / TR_BLOCK;
    unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)sc.sm_count * per_sm, std::max<unsigned long long>(want, 1ull));
    trace_kernel<TECH, STATS, SRC, ANYHIT, MINB, TAPER><<<grid, TR_BLOCK, 0, st>>>(p);
    count_launch();
    return VKHRT_OK;
}
template <int TECH, bool STATS, int PL_S, int PL_STK, int MINB = PL_MINB, bool TAPER = false, bool AO = false>
static int launch_pool_t(DeviceScene& sc, TraceParams& p, cudaStream_t st)
{
    const int carve = tun().carveout;
    if (carve >= 0) VK_CUDA(cudaFuncSetAttribute(trace_pool_kernel<TECH, STATS, PL_S, PL_STK, MINB, TAPER, AO>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    int per_sm = blocks_per_sm(trace_pool_kernel<TECH, STATS, PL_S, PL_STK, MINB, TAPER, AO>, sc.device);
    if (tun().blocks_per_sm > 0) per_sm = std::min(per_sm, tun().blocks_per_sm);
    unsigned long long want = ((unsigned long long)(p.n_slots - p.slot_begin) + TR_BLOCK - 1) / TR_BLOCK;
    unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)sc.sm_count * per_sm, std::max<unsigned long long>(want, 1ull));
    const size_t ovf = (size_t)grid * (TR_BLOCK / 32) * PL_S * PL_OVF;
    if (sc.pool_overflow_n < ovf) {
        if (sc.d_pool_overflow) cudaFree(sc.d_pool_overflow);
        sc.d_pool_overflow = nullptr; sc.pool_overflow_n = 0;
        VK_CUDA(cudaMalloc(&sc.d_pool_overflow, ovf * sizeof(uint2)));
        sc.pool_overflow_n = ovf;
    }
    p.pool_overflow = sc.d_pool_overflow;
    trace_pool_kernel<TECH, STATS, PL_S, PL_STK, MINB, TAPER, AO><<<grid, TR_BLOCK, 0, st>>>(p);
    sc.last_trace_was_pool = true;
    count_launch();
    return VKHRT_OK;
}
template <int TECH, bool STATS, bool TAPER = false>
static int launch_pool(DeviceScene& sc, TraceParams& p, cudaStream_t st)
{
    // slots per warp x shared-memory stack window (profiles/experiments/r02_pool_kernel.txt): 72 x 4 measured best; 56 x 8 and 64 x 6
    // trade slots for fewer spill reads (-0.5 % / -1 %); 9 / 10 / 12 CTAs per SM at 56 / 48 / 40 registers lose 3 / 16 / 22 %
    if constexpr (TAPER) return launch_pool_t<TECH, STATS, 72, 4, PL_MINB, true>(sc, p, st);          // the experiment configurations are not instantiated for tapered scenes
    else {
        switch (tun().pool_cfg) {
        case 1: return launch_pool_t<TECH, STATS, 56, 8>(sc, p, st);
        case 2: return launch_pool_t<TECH, STATS, 64, 6>(sc, p, st);
        default: return launch_pool_t<TECH, STATS, 72, 4>(sc, p, st);
        }
    }
}
// ambient-occlusion passes through the pool kernel: 64 slots per warp (a slot also carries the ray's origin)
template <bool STATS, bool TAPER>
static int launch_pool_ao(DeviceScene& sc, TraceParams& p, cudaStream_t st)
{
    return launch_pool_t<VKHRT_TECHNIQUE_PHANTOM, STATS, 64, 4, PL_MINB, TAPER, true>(sc, p, st);
}

// does the per-warp ray-pool kernel serve this scene's primary rays? (Phantom, with or without per-vertex radii; LSS / DOTS behind their switches)
static bool pool_serves(const DeviceScene& sc)
{
    if (!tun().pool) return false;
    if (sc.tapered()) return sc.technique == VKHRT_TECHNIQUE_PHANTOM && tun().pool_taper;      // per-vertex radii: Phantom only
    return sc.technique == VKHRT_TECHNIQUE_PHANTOM || (sc.technique == VKHRT_TECHNIQUE_LSS && tun().pool_lss) || (sc.technique == VKHRT_TECHNIQUE_DOTS && tun().pool_dots);
}

template <bool STATS, int SRC, bool ANYHIT>
static int launch_trace(DeviceScene& sc, TraceParams& p, cudaStream_t st)
{
    tunables(p);
    sc.last_trace_was_pool = false;
    p.hits_aligned32 = (((uintptr_t)p.hits & 31u) == 0u ? 1u : 0u) | (((uintptr_t)p.hits_mirror & 31u) == 0u ? 2u : 0u);
    if (tun().store256 == 0) p.hits_aligned32 = 0u;
    // occlusion rays of a Phantom scene, enough of them to fill the pool: the pool kernel's ambient-occlusion variant
    if (SRC == SRC_AO && ANYHIT && sc.technique == VKHRT_TECHNIQUE_PHANTOM && tun().pool && tun().pool_ao && p.n_prims && (!STATS || tun().pool_stats) &&
        (unsigned long long)(p.n_slots - p.slot_begin) >= (unsigned long long)tun().pool_min_ratio * sc.sm_count * 32ull * 56ull)
        return sc.tapered() ? launch_pool_ao<STATS, true>(sc, p, st) : launch_pool_ao<STATS, false>(sc, p, st);
    if (sc.tapered()) {
        // per-vertex radii (Phantom, DOTS): the lane-bound kernel with the taper terms compiled in (LSS always carries its radii)
        if (sc.technique == VKHRT_TECHNIQUE_PHANTOM) {
            // primary rays of large frames: the pool kernel with the taper terms (radius(t), cone slant) in its set-up and march stages
            if (SRC == SRC_PRIMARY && !ANYHIT && pool_serves(sc) && p.n_prims && (!STATS || tun().pool_stats) && !(p.host_dest && tun().pool_host == 0) &&
                (unsigned long long)(p.n_slots - p.slot_begin) >= (unsigned long long)tun().pool_min_ratio * sc.sm_count * 32ull * 56ull)
                return launch_pool<VKHRT_TECHNIQUE_PHANTOM, STATS, true>(sc, p, st);
            return launch_trace_t<VKHRT_TECHNIQUE_PHANTOM, STATS, SRC, ANYHIT, TR_MIN_BLOCKS, true>(sc, p, st);
        }
        return launch_trace_t<VKHRT_TECHNIQUE_DOTS, STATS, SRC, ANYHIT, TR_MIN_BLOCKS, true>(sc, p, st);
    }
    switch (sc.technique) {
    case VKHRT_TECHNIQUE_PHANTOM:
        // records that go straight to pinned host memory keep trace_kernel: its retiring lanes are pixel neighbours, which the
        // PCIe write path combines better (e2e 997 vs 819 Mrays/s on C2)
        // ... and small frames too: the pool needs several times its resident capacity (SMs x 32 warps x 56 slots = 265 k rays
        // on a B200) in rays to reach a steady state; below that it is all ramp-up and drain (C1: 639 vs 801 Mrays/s)
        if (SRC == SRC_PRIMARY && !ANYHIT && tun().pool && p.n_prims && (!STATS || tun().pool_stats) && !(p.host_dest && tun().pool_host == 0) &&
            (unsigned long long)(p.n_slots - p.slot_begin) >= (unsigned long long)tun().pool_min_ratio * sc.sm_count * 32ull * 56ull)
            return launch_pool<VKHRT_TECHNIQUE_PHANTOM, STATS>(sc, p, st);
        if (!STATS && SRC == SRC_PRIMARY && tun().min_blocks == 7) return launch_trace_t<VKHRT_TECHNIQUE_PHANTOM, false, SRC_PRIMARY, false, 7>(sc, p, st);
        if (!STATS && SRC == SRC_PRIMARY && tun().min_blocks == 8) return launch_trace_t<VKHRT_TECHNIQUE_PHANTOM, false, SRC_PRIMARY, false, 8>(sc, p, st);
        return launch_trace_t<VKHRT_TECHNIQUE_PHANTOM, STATS, SRC, ANYHIT, TR_MIN_BLOCKS>(sc, p, st);
    case VKHRT_TECHNIQUE_LSS:
        if (SRC == SRC_PRIMARY && !ANYHIT && pool_serves(sc) && p.n_prims && (!STATS || tun().pool_stats) && !(p.host_dest && tun().pool_host == 0) &&
            (unsigned long long)(p.n_slots - p.slot_begin) >= (unsigned long long)tun().pool_min_ratio * sc.sm_count * 32ull * 56ull)
            return launch_pool<VKHRT_TECHNIQUE_LSS, STATS>(sc, p, st);
        return launch_trace_t<VKHRT_TECHNIQUE_LSS, STATS, SRC, ANYHIT, TR_MIN_BLOCKS>(sc, p, st);
    default:
        if (SRC == SRC_PRIMARY && !ANYHIT && pool_serves(sc) && p.n_prims && (!STATS || tun().pool_stats) && !(p.host_dest && tun().pool_host == 0) &&
            (unsigned long long)(p.n_slots - p.slot_begin) >= (unsigned long long)tun().pool_min_ratio * sc.sm_count * 32ull * 56ull)
            return launch_pool<VKHRT_TECHNIQUE_DOTS, STATS>(sc, p, st);
        return launch_trace_t<VKHRT_TECHNIQUE_DOTS, STATS, SRC, ANYHIT, TR_MIN_BLOCKS>(sc, p, st);
    }
}

template <typename T>
static int grow(T** p, size_t* have, size_t want)
{
    if (*have >= want) return VKHRT_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    VK_CUDA(cudaMalloc((void**)p, want * sizeof(T)));
    *have = want;
    return VKHRT_OK;
}

int render_frame(DeviceScene& sc, const VkhrtFrameDesc& f, VkhrtHit* hits_out, uint8_t* rgba_out, VkhrtTraceStats* stats, const RenderOpts& opts)
{
    VK_CUDA(cudaSetDevice(sc.device));
    Resolved r;
    if (!resolve(f, r)) { set_last_error("vkhrt_render: bad frame description (size, tile_size multiple of 8, tile_first < tile_stride)"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (f.miss_mode == VKHRT_MISS_ENVIRONMENT && !sc.d_env) { set_last_error("vkhrt_render: miss_mode ENVIRONMENT without vkhrt_scene_set_environment"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (f.miss_mode != VKHRT_MISS_CONSTANT && f.miss_mode != VKHRT_MISS_ENVIRONMENT) { set_last_error("vkhrt_render: unknown miss_mode"); return VKHRT_ERR_INVALID_ARGUMENT; }
    const bool want_rgba = rgba_out != nullptr;
    const bool want_hits = hits_out != nullptr;
    if (want_rgba && sc.mesh_table_dirty) {
        // mesh table {albedo.rgb, first segment}: a few entries, uploaded when vkhrt_scene_set_meshes / _set_mesh_material changed it
        const size_t n = sc.mesh_first.size();
        std::vector<float4> t(n);
        for (size_t m = 0; m < n; ++m) t[m] = make_float4(sc.mesh_albedo[4 * m], sc.mesh_albedo[4 * m + 1], sc.mesh_albedo[4 * m + 2], bits_to_float(sc.mesh_first[m]));
        VK_CUDA(cudaStreamSynchronize(f.stream ? (cudaStream_t)f.stream : sc.stream));
        cudaFree(sc.d_mesh_table); sc.d_mesh_table = nullptr;
        if (n) { VK_CUDA(cudaMalloc(&sc.d_mesh_table, n * sizeof(float4))); VK_CUDA(cudaMemcpy(sc.d_mesh_table, t.data(), n * sizeof(float4), cudaMemcpyHostToDevice)); }
        sc.mesh_table_dirty = false;
    }
    // where each output lives (vkhrt_render_multi sends the two to different places: records to the caller's pinned host
    // buffer, pixels to the gathering GPU's frame buffer)
    const bool host_frame = f.output_memory == VKHRT_MEM_HOST;
    const bool hits_host = want_hits && host_frame && !opts.hits_on_device;
    const bool rgba_host = want_rgba && host_frame && !opts.rgba_on_device;
    const bool any_host = hits_host || rgba_host;
    // a shard that writes at row-major positions of a buffer shared with other shards may only touch its own pixels
    const bool shared_frame = r.tile_stride > 1 && f.row_major_output;
    cudaStream_t st = opts.stream ? opts.stream : ((!host_frame && f.stream) ? (cudaStream_t)f.stream : sc.stream);
    int rc;

    // device-side destinations: sample-0 hit records go to the caller's buffer when it is device memory,
    // otherwise to scratch[0, n_out); samples >= 1 (only traced when an image is wanted) use scratch[n_out, 2 n_out)
    const bool multi = r.spp > 1 && want_rgba;
    const uint32_t ao = want_rgba ? f.ao_samples : 0u;     // AO only changes the image
    // with AO the passes re-read the primary hit records: keep them in this GPU's HBM and mirror them to the caller's
    // buffer (which may be a peer GPU's frame buffer, row_major_output) instead of reading them back over NVLink
    const bool direct_hits = want_hits && !hits_host && ao == 0u;
    VkhrtHit* d_hits_mirror = (want_hits && !hits_host && ao != 0u) ? hits_out : nullptr;
    bool direct_to_host = false;
    VkhrtHit* d_hits0 = nullptr;
    VkhrtHit* d_hits_other = nullptr;
    uint8_t* d_rgba = nullptr;
    // Host hit buffer in pinned (page-locked) memory: the traversal kernel stores each record straight into it over
    // PCIe (posted writes, fully overlapped with the traversal) instead of a device->host copy after the
    // kernel.  When an image is wanted too, the records are also kept in HBM for the shading kernel.
    VkhrtHit* h_hits_mapped = nullptr;
    if (hits_host && !stats && tun().zero_copy) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, hits_out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            h_hits_mapped = static_cast<VkhrtHit*>(at.devicePointer);
        cudaGetLastError();
    }
    if (shared_frame && ((hits_host && !h_hits_mapped) || rgba_host)) {
        set_last_error("vkhrt_render: row_major_output with tile_stride > 1 writes into a frame buffer shared by all shards: device output memory, "
                       "or (hit records only) a page-locked host buffer that the kernel can store into directly");
        return VKHRT_ERR_INVALID_ARGUMENT;
    }
    // samples per launch for the samples after the first (see the loop below): enough for ~20 M rays, at most 16, none with AO passes
    uint32_t batch_k = 1u;
    if (multi && ao == 0u && r.n_slots * 16ull < 0xFFFFFFFFull && tun().sample_batch) {
        const unsigned long long target = (unsigned long long)std::max(1, tun().batch_mrays) * 1000000ull, kmax = (unsigned long long)std::min(16, std::max(1, tun().batch_max));
        batch_k = (uint32_t)std::min<unsigned long long>(kmax, std::max<unsigned long long>(1ull, (target + r.n_slots - 1) / r.n_slots));
        batch_k = std::min(batch_k, std::max(1u, r.spp - 1u));
        if (r.n_out * (unsigned long long)batch_k >= 0xFFFFFFFFull) batch_k = 1u;          // 32-bit record indices inside the kernels
    }
    if ((want_hits || want_rgba) && (!direct_hits || multi) && !(h_hits_mapped && !want_rgba)) {
        if ((rc = grow(&sc.d_hits_scratch, &sc.hits_scratch_n, (size_t)r.n_out * (multi ? 1 + batch_k : 1)))) return rc;
    }
    if (want_hits || want_rgba) d_hits0 = direct_hits ? hits_out : sc.d_hits_scratch;
    // Phantom frames large enough for the pool kernel: line-wise delivery (records to HBM, complete 128-byte lines to the host)
    bool linewise = false;
    const uint32_t line_shift = (uint32_t)tun().line_shift;
    // (with an image the records stay in HBM for the shading kernel anyway, which is where line-wise delivery keeps them)
    if (h_hits_mapped && pool_serves(sc) && ao == 0u && sc.n_leaves && tun().linewise &&
        (((uintptr_t)h_hits_mapped) & ((32u << line_shift) - 1u)) == 0u && r.n_slots >= (unsigned long long)tun().pool_min_ratio * sc.sm_count * 32ull * 56ull) {
        if ((rc = grow(&sc.d_hits_scratch, &sc.hits_scratch_n, (size_t)r.n_out))) return rc;
        if ((rc = grow(&sc.d_line_cnt, &sc.line_cnt_n, (size_t)r.n_out / 2 + 1))) return rc;      // enough for the smallest line (2 records)
        linewise = true;
    }
    VkhrtHit* h_lines = nullptr;
    if (linewise) { h_lines = h_hits_mapped; d_hits0 = sc.d_hits_scratch; h_hits_mapped = nullptr; }
    // Several samples per pixel and a page-locked record buffer: the records belong to sample 0, so instead of mirroring every record
    // over PCIe from inside the sample-0 kernel (which then runs at the link's pace: C3 2.1 ms instead of 0.93) they stay in HBM and
    // the copy engine takes them out while the other samples are traced.
    bool early_copy = false;
    if (h_hits_mapped && multi && !shared_frame && tun().early_copy) { early_copy = true; h_hits_mapped = nullptr; }
    if (h_hits_mapped && !want_rgba) { d_hits0 = h_hits_mapped; h_hits_mapped = nullptr; direct_to_host = true; }   // single destination
    if (multi) d_hits_other = sc.d_hits_scratch + r.n_out;
    if (want_rgba) {
        if (rgba_host) { if ((rc = grow(&sc.d_rgba_scratch, &sc.rgba_scratch_n, (size_t)r.n_out * 4))) return rc; d_rgba = sc.d_rgba_scratch; }
        else d_rgba = rgba_out;
        if (multi) { if ((rc = grow(&sc.d_accum, &sc.accum_n, (size_t)r.n_out))) return rc; }
        if (ao) { if ((rc = grow(&sc.d_occluded, &sc.occluded_n, (size_t)r.n_out))) return rc; }
    }

    TraceParams p;
    fill_params(sc, f, r, p);
    const float3 miss = make_float3(f.miss_rgb[0], f.miss_rgb[1], f.miss_rgb[2]);
    cudaEvent_t* ev = sc.ev;
    VK_CUDA(cudaEventRecord(ev[6], st));
    if (stats) VK_CUDA(cudaMemsetAsync(sc.d_counters, 0, 16 * sizeof(unsigned long long), st));
    const uint32_t n_samples = (want_rgba || stats) ? r.spp : 1;    // hits only => sample 0 is all that is observable

    // samples >= 1 of a small frame or shard go several per launch (TraceParams::n_batch): the persistent kernels need a few
    // million rays to reach a steady state, and a launch per sample pays ramp-up and drain 63 times on a 64-spp shard
    const uint32_t batch_max = batch_k;
    for (uint32_t s = 0; s < n_samples;) {
        const uint32_t kb = (s == 0u) ? 1u : std::min(batch_max, n_samples - s);
        sample_offset(s, &p.sx, &p.sy);
        p.n_batch = kb; p.slots_per_sample = (uint32_t)r.n_slots; p.n_slots = (uint32_t)(r.n_slots * kb);
        for (uint32_t b = 0; b < kb; ++b) sample_offset(s + b, &p.bsx[b], &p.bsy[b]);
        p.hits = s == 0 ? d_hits0 : d_hits_other;
        p.hits_mirror = s == 0 ? (h_hits_mapped ? h_hits_mapped : d_hits_mirror) : nullptr;
        p.host_dest = (s == 0 && (direct_to_host || h_hits_mapped)) ? 1u : 0u;
        p.line_cnt = nullptr; p.host_lines = nullptr;
        if (linewise && s == 0) {
            p.line_cnt = sc.d_line_cnt; p.host_lines = h_lines; p.line_shift = line_shift;
            VK_CUDA(cudaMemsetAsync(sc.d_line_cnt, 0, (((size_t)r.n_out >> p.line_shift) + 1) * sizeof(uint32_t), st));
        }
        take_work_counter(sc, p);
        if (s == 0) VK_CUDA(cudaEventRecord(ev[7], st));
        rc = stats ? launch_trace<true, SRC_PRIMARY, false>(sc, p, st) : launch_trace<false, SRC_PRIMARY, false>(sc, p, st);
        if (rc) return rc;
        if (s == 0) VK_CUDA(cudaEventRecord(ev[8], st));
        // line-wise delivery is a feature of the pool kernel: if the dispatch picked the lane-bound kernel after all, the records are
        // in HBM only and go out with a plain copy (whole buffer: not possible for a shard of a shared frame)
        if (linewise && s == 0 && !sc.last_trace_was_pool) {
            if (shared_frame) { set_last_error("vkhrt_render: line-wise host delivery fell back to a whole-buffer copy on a shared frame"); return VKHRT_ERR_UNSUPPORTED; }
            VK_CUDA(cudaMemcpyAsync(hits_out, d_hits0, (size_t)r.n_out * sizeof(VkhrtHit), cudaMemcpyDeviceToHost, st));
        }
        if (ao) {
            // secondary rays: ao passes of one occlusion ray per hit pixel, spawned from the hit records inside the
            // traversal kernel's refill step (no ray buffer), terminate-on-first-hit  (never batched: kb == 1)
            VK_CUDA(cudaMemsetAsync(sc.d_occluded, 0, (size_t)r.n_out * sizeof(uint32_t), st));
            TraceParams q = p;
            q.ao_hits = p.hits; q.ao_occluded = sc.d_occluded; q.ao_sample = s;
            q.ao_distance = f.ao_distance > 0.0f ? f.ao_distance : VKHRT_DEFAULT_AO_DISTANCE;
            q.ao_bias = f.ao_bias > 0.0f ? f.ao_bias : 0.25f * sc.radius;
            q.hits = nullptr; q.hits_mirror = nullptr;
            // the passes of a sample go up to 8 to a launch (the persistent kernel ramps up and drains once instead of `ao` times);
            // the occlusion counts are integer sums, so the order the rays finish in cannot change them
            const uint32_t ao_batch = (tun().sample_batch && r.n_slots * 8ull < 0xFFFFFFFFull) ? 8u : 1u;
            for (uint32_t a = 0; a < ao;) {
                const uint32_t ka = std::min(ao_batch, ao - a);
                q.ao_index = a;
                q.n_batch = ka; q.slots_per_sample = (uint32_t)r.n_slots; q.n_slots = (uint32_t)(r.n_slots * ka);
                take_work_counter(sc, q);
                rc = stats ? launch_trace<true, SRC_AO, true>(sc, q, st) : launch_trace<false, SRC_AO, true>(sc, q, st);
                if (rc) return rc;
                a += ka;
            }
        }
        if (s == 0) VK_CUDA(cudaEventRecord(ev[12], st));
        if (s == 0 && early_copy) {
            VK_CUDA(cudaEventRecord(ev[13], st));
            VK_CUDA(cudaStreamWaitEvent(sc.copy_stream, ev[13], 0));
            VK_CUDA(cudaMemcpyAsync(hits_out, d_hits0, (size_t)r.n_out * sizeof(VkhrtHit), cudaMemcpyDeviceToHost, sc.copy_stream));
            VK_CUDA(cudaEventRecord(ev[14], sc.copy_stream));
        }
        if (want_rgba) {
            // one shading pass per sample, in sample order (the fp32 accumulation order is part of the result)
            const bool env = f.miss_mode == VKHRT_MISS_ENVIRONMENT && sc.d_env;
            TraceParams ps = p;
            ps.n_batch = 1u; ps.n_slots = (uint32_t)r.n_slots;
            for (uint32_t b = 0; b < kb; ++b) {
                ps.sx = p.bsx[b]; ps.sy = p.bsy[b];
                shade_kernel<<<(unsigned)((r.n_slots + 255) / 256), 256, 0, st>>>(ps, p.hits + (size_t)b * r.n_out, f.shade_mode, miss, make_float3(sc.albedo[0], sc.albedo[1], sc.albedo[2]), sc.d_accum, (uchar4*)d_rgba, s + b, r.spp,
                                                                                  sc.d_occluded, ao, env ? sc.d_env : nullptr, sc.env_w, sc.env_h,
                                                                                  sc.d_mesh_table, (uint32_t)sc.mesh_first.size());
                count_launch();
            }
        }
        if (s == 0) VK_CUDA(cudaEventRecord(ev[9], st));
        s += kb;
    }
    VK_CUDA(cudaEventRecord(ev[10], st));
    if (early_copy) VK_CUDA(cudaStreamWaitEvent(st, ev[14], 0));          // the frame is complete on `st` when the records have landed too
    else if (hits_host && !h_hits_mapped && !direct_to_host && !linewise) VK_CUDA(cudaMemcpyAsync(hits_out, d_hits0, (size_t)r.n_out * sizeof(VkhrtHit), cudaMemcpyDeviceToHost, st));
    if (rgba_host) VK_CUDA(cudaMemcpyAsync(rgba_out, d_rgba, (size_t)r.n_out * 4, cudaMemcpyDeviceToHost, st));
    VK_CUDA(cudaEventRecord(ev[11], st));
    if (stats) {
        unsigned long long c[16];
        VK_CUDA(cudaMemcpyAsync(c, sc.d_counters, sizeof(c), cudaMemcpyDeviceToHost, st));
        VK_CUDA(cudaStreamSynchronize(st));
        stats->nodes_visited = c[1]; stats->prims_tested = c[2]; stats->hits = c[3]; stats->phantom_iterations = c[4]; stats->rays = c[5];
        for (int k = 0; k < 4; ++k) { stats->sched_steps[k] = c[8 + k]; stats->sched_lanes[k] = c[12 + k]; }
    }
    // host outputs: the call returns when the copies have landed.  Device outputs on the scene's own stream (no caller
    // stream): the caller has no handle to order against, so the call is synchronous too, like vkhrt_trace_rays.
    // opts.defer_sync: the caller (vkhrt_render_multi) waits for all its shards at once.
    if (!opts.defer_sync && (any_host || host_frame || !f.stream)) VK_CUDA(cudaStreamSynchronize(st));
    VK_CUDA(cudaGetLastError());
    return VKHRT_OK;
}

int trace_ray_buffer(DeviceScene& sc, const float* rays_dev, uint64_t n, VkhrtHit* hits_dev, bool any_hit, cudaStream_t stream)
{
    VK_CUDA(cudaSetDevice(sc.device));
    cudaStream_t st = stream ? stream : sc.stream;
    TraceParams p;
    memset(&p, 0, sizeof(p));
    p.nodes = sc.d_nodes; p.primA = sc.d_primA; p.primB = sc.d_primB; p.primR = sc.d_primR; p.sorted_ids = sc.d_sorted_ids;
    p.n_prims = sc.n_leaves; p.radius = sc.radius;
    p.counters = sc.d_counters;
    p.slot_begin = 0; p.hits_mirror = nullptr;
    p.T = 8; p.tiles_x = 1; p.tile_stride = 1;
    // 32-bit slot indices inside the kernel: long ray buffers go in chunks of 2^30 rays
    const uint64_t chunk = 1ull << 30;
    for (uint64_t first = 0; first < n; first += chunk) {
        p.rays = reinterpret_cast<const float4*>(rays_dev) + 2 * first;
        p.hits = hits_dev + first;
        p.n_slots = (uint32_t)std::min<uint64_t>(chunk, n - first);
        take_work_counter(sc, p);
        int rc = any_hit ? launch_trace<false, SRC_BUFFER, true>(sc, p, st) : launch_trace<false, SRC_BUFFER, false>(sc, p, st);
        if (rc) return rc;
    }
    if (!stream) VK_CUDA(cudaStreamSynchronize(st));
    VK_CUDA(cudaGetLastError());
    return VKHRT_OK;
}

int generate_ray_buffer(const VkhrtFrameDesc& f, uint32_t sample, float* rays_dev, cudaStream_t stream)
{
    Resolved r;
    if (!resolve(f, r)) { set_last_error("vkhrt_generate_rays: bad frame description"); return VKHRT_ERR_INVALID_ARGUMENT; }
    TraceParams p;
    memset(&p, 0, sizeof(p));
    DeviceScene dummy;
    fill_params(dummy, f, r, p);
    sample_offset(sample, &p.sx, &p.sy);
    // padding slots of the non-compact layout have no output position: generate exactly the n_out records
    raygen_kernel<<<(unsigned)((p.n_slots + 255) / 256), 256, 0, stream>>>(p, reinterpret_cast<float4*>(rays_dev));
    count_launch();
    VK_CUDA(cudaGetLastError());
    return VKHRT_OK;
}

int untile_buffer(const VkhrtFrameDesc& f, uint32_t world, const void* gathered, void* row_major, uint32_t elem_bytes, cudaStream_t stream)
{
    VkhrtFrameDesc g = f;
    g.tile_first = 0; g.tile_stride = world;
    Resolved r;
    if (!resolve(g, r) || world == 0) { set_last_error("vkhrt_untile: bad frame description"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (elem_bytes != 4 && elem_bytes != 32) { set_last_error("vkhrt_untile: elem_bytes must be 4 or 32"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (world == 1) {
        // a one-shard render is never compact (tile_stride 1 writes row-major, W*H elements): plain copy, as vkhrt_untile_host
        VK_CUDA(cudaMemcpyAsync(row_major, gathered, (size_t)r.W * r.H * elem_bytes, cudaMemcpyDeviceToDevice, stream));
        return VKHRT_OK;
    }
    unsigned long long shard = (unsigned long long)r.n_local_tiles * r.T * r.T;
    unsigned grid = (unsigned)(((unsigned long long)r.W * r.H + 255) / 256);
    if (elem_bytes == 4) untile_kernel<uint32_t><<<grid, 256, 0, stream>>>((const uint32_t*)gathered, (uint32_t*)row_major, r.W, r.H, r.T, r.tiles_x, world, shard);
    else if (elem_bytes == 32) untile_kernel<Elem32><<<grid, 256, 0, stream>>>((const Elem32*)gathered, (Elem32*)row_major, r.W, r.H, r.T, r.tiles_x, world, shard);
    else { set_last_error("vkhrt_untile: elem_bytes must be 4 or 32"); return VKHRT_ERR_INVALID_ARGUMENT; }
    count_launch();
    VK_CUDA(cudaGetLastError());
    return VKHRT_OK;
}

}  // namespace vkhrt
