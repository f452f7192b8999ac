// scene.h — internal (not part of the ABI): device-resident scene + launch helpers.
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/vkhrt_b200.h"

namespace vkhrt {

// Device buffers of one scene.  Everything is sized for 32-bit primitive indices (< 2^31).
struct DeviceScene {
    int device = 0;
    int technique = 0;
    float radius = VKHRT_DEFAULT_RADIUS;
    uint32_t n_vertices = 0, n_segments = 0, n_prims = 0, n_nodes = 0;
    // BVH leaves: VKHRT_LEAF_SPLIT_* pieces per group (PHANTOM curve, LSS, DOTS strip = the 4 triangles of a segment)
    uint32_t n_leaves = 0;
    bool built = false;

    // input (Assimp-shaped line mesh)
    float* d_positions = nullptr;      // n_vertices * 3
    uint32_t* d_indices = nullptr;     // n_segments * 2
    float* d_radius_pv = nullptr;      // n_vertices or null
    float* d_curves = nullptr;         // null unless vkhrt_scene_apply_lod merged curves: n_segments * 12 floats, replaces GenerateCurves
    float lod_ms = 0.0f;
    bool lod_applied = false;          // the vertex list was rewritten by the LOD passes: refit with caller positions is refused

    // material (bindless.glsl Material): albedoFactor * albedo map at texCoord (0,0), evaluated when the material is set
    float albedo[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    // meshes (vkhrt_scene_set_meshes): mesh m owns segments [mesh_first[m], mesh_first[m + 1]); one LBVH over all of them.
    // d_mesh_table[m] = {albedo.rgb, bits(first segment)}, uploaded when it changed; empty = one mesh with `albedo`
    std::vector<uint32_t> mesh_first;
    std::vector<float> mesh_albedo;    // 4 per mesh
    float4* d_mesh_table = nullptr; bool mesh_table_dirty = false;

    // environment map (RGBA32F, miss.rmiss)
    float4* d_env = nullptr;
    uint32_t env_w = 0, env_h = 0;

    // acceleration structure + sorted primitives: views into ONE allocation (d_arena)
    unsigned char* d_arena = nullptr; size_t arena_bytes = 0;
    float4* d_nodes = nullptr;         // n_nodes * 4 float4 (VkhrtBvhNode)
    uint32_t* d_sorted_ids = nullptr;  // n_leaves: original leaf id (segment) at each Morton-sorted position
    uint64_t* d_sorted_morton = nullptr;
    uint32_t* d_parent_internal = nullptr;  // n_nodes: (parent << 1) | slot
    uint32_t* d_parent_leaf = nullptr;      // n_leaves
    uint32_t* d_refit_flags = nullptr;      // n_nodes
    float4* d_refit_exits = nullptr; uint32_t* d_refit_exit_count = nullptr; uint32_t refit_exit_cap = 0;   // walkers that leave their refit tile
    uint8_t* d_node_local = nullptr;        // n_nodes: bit 0 = the node's leaf range lies inside one refit tile (handled in shared memory), bit 1 = child 1 is a leaf
    uint32_t* d_child0 = nullptr;           // n_nodes: child 0 (child 1 = its position + 1): the refit writes whole records of tile-local nodes
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
    uint32_t h_bounds[6] = {};            // centroid bounds as ordered uints, copied back at the end of a build
    unsigned char* d_build_scratch = nullptr; size_t build_scratch_bytes = 0;   // centroids, sort ping-pong, histograms, look-back status

    // primitives in Morton-sorted order
    //   PHANTOM: primA[2p] = {B0.xyz, rmax}, primA[2p+1] = {B3.xyz, bits(prim id)}; primB[2p] = {B1.xyz, quarter-chord deviation}, primB[2p+1] = {B2.xyz,0}
    //   LSS:     primA[2p] = {p0.xyz, r0},   primA[2p+1] = {p1.xyz, r1}
    //   DOTS:    one 64-byte strip record per segment: primA[4p] = {start.xyz, bits(segment id)}, primA[4p+1] = {end.xyz, 0},
    //            primA[4p+2] = {v0*r, 0}, primA[4p+3] = {v1*r, 0}; the 12 strip vertices are start/end -+ these offsets
    float4* d_primA = nullptr;
    float4* d_primB = nullptr;
    //   per-vertex radii (d_radius_pv != null): PHANTOM adds primR[p] = {r0, r1} of the curve's two ends; the DOTS record becomes
    //   primA[4p+2] = {v0 (unit), r0}, primA[4p+3] = {v1 (unit), r1}; LSS already carries its radii
    float2* d_primR = nullptr;
    bool tapered() const { return d_radius_pv != nullptr && technique != VKHRT_TECHNIQUE_LSS; }

    // per-frame scratch (grown on demand)
    unsigned long long* d_counters = nullptr;   // [0], [6] alternating work counters (a launch zeroes the other one), [1..5] stats, [8..15] scheduler
    bool work_flip = false;
    VkhrtHit* d_hits_scratch = nullptr; size_t hits_scratch_n = 0;
    float4* d_accum = nullptr; size_t accum_n = 0;
    uint32_t* d_occluded = nullptr; size_t occluded_n = 0;
    bool last_trace_was_pool = false;
    uint32_t* d_line_cnt = nullptr; size_t line_cnt_n = 0;             // line-wise host delivery: records written per 128-byte line
    uint2* d_pool_overflow = nullptr; size_t pool_overflow_n = 0;   // trace_pool_kernel: stack entries beyond the shared-memory slots   // per pixel: occluded AO rays of the current sample
    uint8_t* d_rgba_scratch = nullptr; size_t rgba_scratch_n = 0;
    void* h_pinned = nullptr; size_t h_pinned_bytes = 0;   // staging for host outputs
    // vkhrt_render_multi: full-frame buffers on the gathering GPU (scenes[0]'s device); the other GPUs store into them over NVLink
    VkhrtHit* d_multi_hits = nullptr; size_t multi_hits_n = 0;
    uint8_t* d_multi_rgba = nullptr; size_t multi_rgba_n = 0;
    cudaEvent_t multi_done = nullptr;

    // frames in flight (vkhrt_render_submit / _wait): per slot the device buffers the copy engine reads while the next frame traverses
    struct InFlight {
        VkhrtHit* d_hits = nullptr; size_t hits_n = 0;
        uint8_t* d_rgba = nullptr; size_t rgba_n = 0;
        cudaEvent_t traced = nullptr, copied = nullptr;
        bool pending = false;
    } fl[VKHRT_FRAMES_IN_FLIGHT];
    uint32_t fl_head = 0, fl_count = 0;     // oldest outstanding slot, number outstanding

    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev[16] = {};
    VkhrtTiming timing = {};
    int sm_count = 148;
};

void set_last_error(const std::string& s);
void count_launch(uint64_t n = 1);

#define VK_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ::vkhrt::set_last_error(std::string(#call) + ": " + cudaGetErrorString(e_));           \
            return e_ == cudaErrorMemoryAllocation ? VKHRT_ERR_OUT_OF_MEMORY : VKHRT_ERR_CUDA;     \
        }                                                                                          \
    } while (0)

inline uint32_t leaf_split_of(int technique)
{
    return technique == VKHRT_TECHNIQUE_PHANTOM ? VKHRT_LEAF_SPLIT_PHANTOM : (technique == VKHRT_TECHNIQUE_LSS ? VKHRT_LEAF_SPLIT_LSS : VKHRT_LEAF_SPLIT_DOTS);
}

// build.cu
int build_scene(DeviceScene& sc, bool refit_only);
int export_primitives(DeviceScene& sc, float* host_out, size_t out_floats);
int apply_lod(DeviceScene& sc, uint32_t split_passes, uint32_t merge_passes, uint32_t curve_merge_passes);
int export_lines(DeviceScene& sc, float* host_out, size_t out_floats);

// trace.cu
struct RenderOpts {
    bool defer_sync = false;       // enqueue only; the caller synchronises sc.stream itself
    bool hits_on_device = false;   // hits_out is a DEVICE pointer although frame.output_memory is HOST (vkhrt_render_multi)
    bool rgba_on_device = false;   // same for rgba8_out
    cudaStream_t stream = nullptr; // run on this stream instead of the scene's / the frame's
};
int render_frame(DeviceScene& sc, const VkhrtFrameDesc& f, VkhrtHit* hits_out, uint8_t* rgba_out, VkhrtTraceStats* stats, const RenderOpts& opts = RenderOpts());
int trace_ray_buffer(DeviceScene& sc, const float* rays_dev, uint64_t n, VkhrtHit* hits_dev, bool any_hit, cudaStream_t stream);
int generate_ray_buffer(const VkhrtFrameDesc& f, uint32_t sample, float* rays_dev, cudaStream_t stream);
int untile_buffer(const VkhrtFrameDesc& f, uint32_t world, const void* gathered, void* row_major, uint32_t elem_bytes,
                  cudaStream_t stream);
uint64_t frame_local_pixels(const VkhrtFrameDesc& f);
void init_tunables();   // reads the environment switches once (call before launching from several host threads)

}  // namespace vkhrt
