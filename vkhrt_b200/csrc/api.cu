// api.cu — the extern "C" boundary declared in include/vkhrt_b200.h.
// No CPU fallback lives here: every compute entry point needs a CUDA device.
#include "scene.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <cmath>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace vkhrt {

static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

void set_last_error(const std::string& s) { g_last_error = s; }
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int check_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_last_error("no CUDA device available (this library has no CPU path)");
        return VKHRT_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { set_last_error("device ordinal out of range"); return VKHRT_ERR_INVALID_ARGUMENT; }
    return VKHRT_OK;
}

__global__ void validate_indices_kernel(const uint32_t* idx, uint32_t n, uint32_t n_vertices, uint32_t* bad)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && idx[i] >= n_vertices) atomicAdd(bad, 1u);
}

static void free_scene(DeviceScene* sc)
{
    if (!sc) return;
    cudaSetDevice(sc->device);
    cudaFree(sc->d_positions); cudaFree(sc->d_indices); cudaFree(sc->d_radius_pv); cudaFree(sc->d_curves); cudaFree(sc->d_env); cudaFree(sc->d_mesh_table);
    cudaFree(sc->d_build_scratch);
    cudaFree(sc->d_arena);       // nodes, sorted ids / keys, parents, refit flags, primA, primB
    for (auto& f : sc->fl) { cudaFree(f.d_hits); cudaFree(f.d_rgba); if (f.traced) cudaEventDestroy(f.traced); if (f.copied) cudaEventDestroy(f.copied); }
    cudaFree(sc->d_multi_hits); cudaFree(sc->d_multi_rgba);
    if (sc->multi_done) cudaEventDestroy(sc->multi_done);
    cudaFree(sc->d_counters); cudaFree(sc->d_hits_scratch); cudaFree(sc->d_accum); cudaFree(sc->d_rgba_scratch); cudaFree(sc->d_occluded); cudaFree(sc->d_pool_overflow); cudaFree(sc->d_line_cnt);
    if (sc->h_pinned) cudaFreeHost(sc->h_pinned);
    for (auto& e : sc->ev) if (e) cudaEventDestroy(e);
    if (sc->stream) cudaStreamDestroy(sc->stream);
    if (sc->copy_stream) cudaStreamDestroy(sc->copy_stream);
    delete sc;
}

}  // namespace vkhrt

using namespace vkhrt;

struct VkhrtScene { DeviceScene s; };   // opaque handle == DeviceScene

extern "C" {

int vkhrt_abi_version(void) { return VKHRT_ABI_VERSION; }

int vkhrt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* vkhrt_error_string(int status)
{
    switch (status) {
    case VKHRT_OK: return "ok";
    case VKHRT_ERR_INVALID_ARGUMENT: return "invalid argument";
    case VKHRT_ERR_NO_DEVICE: return "no CUDA device (there is no CPU path)";
    case VKHRT_ERR_CUDA: return "CUDA runtime error";
    case VKHRT_ERR_OUT_OF_MEMORY: return "out of device memory";
    case VKHRT_ERR_NOT_BUILT: return "scene not built";
    case VKHRT_ERR_BAD_TOPOLOGY: return "line index out of range";
    case VKHRT_ERR_UNSUPPORTED: return "unsupported";
    case VKHRT_ERR_IO: return "asset or image file could not be read, parsed or written";
    default: return "unknown status";
    }
}

const char* vkhrt_last_error(void) { return g_last_error.c_str(); }
uint64_t vkhrt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int vkhrt_scene_create(const VkhrtSceneDesc* desc, VkhrtScene** out_scene)
{
    if (!desc || !out_scene) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    *out_scene = nullptr;
    if (desc->technique < VKHRT_TECHNIQUE_PHANTOM || desc->technique > VKHRT_TECHNIQUE_DOTS) { set_last_error("unknown technique"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if ((desc->n_vertices && !desc->positions_xyz) || (desc->n_segments && !desc->line_indices)) { set_last_error("null geometry array"); return VKHRT_ERR_INVALID_ARGUMENT; }
    const uint64_t n_prims = desc->technique == VKHRT_TECHNIQUE_DOTS ? (uint64_t)desc->n_segments * 4 : desc->n_segments;
    if (n_prims >= 0xFFFFFFFFull || (uint64_t)desc->n_segments * leaf_split_of(desc->technique) >= 0x7FFFFFFFull) { set_last_error("too many primitives for 31-bit references"); return VKHRT_ERR_UNSUPPORTED; }
    int rc = check_device(desc->device);
    if (rc) return rc;
    VK_CUDA(cudaSetDevice(desc->device));

    DeviceScene* sc = new (std::nothrow) DeviceScene();
    if (!sc) return VKHRT_ERR_OUT_OF_MEMORY;
    sc->device = desc->device;
    sc->technique = desc->technique;
    sc->radius = desc->radius > 0.0f ? desc->radius : VKHRT_DEFAULT_RADIUS;
    sc->n_vertices = desc->n_vertices;
    sc->n_segments = desc->n_segments;
    sc->n_prims = (uint32_t)n_prims;
    sc->n_leaves = desc->n_segments * leaf_split_of(desc->technique);
#define VK_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_last_error(std::string(#call) + ": " + cudaGetErrorString(e_)); free_scene(sc); return e_ == cudaErrorMemoryAllocation ? VKHRT_ERR_OUT_OF_MEMORY : VKHRT_ERR_CUDA; } } while (0)
    cudaDeviceProp prop;
    VK_TRY(cudaGetDeviceProperties(&prop, sc->device));
    sc->sm_count = prop.multiProcessorCount;
    VK_TRY(cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking));
    VK_TRY(cudaStreamCreateWithFlags(&sc->copy_stream, cudaStreamNonBlocking));
    for (auto& e : sc->ev) VK_TRY(cudaEventCreate(&e));
    VK_TRY(cudaMalloc(&sc->d_counters, 16 * sizeof(unsigned long long)));
    VK_TRY(cudaMemsetAsync(sc->d_counters, 0, 16 * sizeof(unsigned long long), sc->stream));
    VK_TRY(cudaMalloc(&sc->d_positions, std::max<size_t>(1, (size_t)sc->n_vertices * 3) * sizeof(float)));
    VK_TRY(cudaMalloc(&sc->d_indices, std::max<size_t>(1, (size_t)sc->n_segments * 2) * sizeof(uint32_t)));
    if (sc->n_vertices) VK_TRY(cudaMemcpyAsync(sc->d_positions, desc->positions_xyz, (size_t)sc->n_vertices * 12, cudaMemcpyHostToDevice, sc->stream));
    if (sc->n_segments) VK_TRY(cudaMemcpyAsync(sc->d_indices, desc->line_indices, (size_t)sc->n_segments * 8, cudaMemcpyHostToDevice, sc->stream));
    if (desc->radius_per_vertex && sc->n_vertices) {
        VK_TRY(cudaMalloc(&sc->d_radius_pv, (size_t)sc->n_vertices * sizeof(float)));
        VK_TRY(cudaMemcpyAsync(sc->d_radius_pv, desc->radius_per_vertex, (size_t)sc->n_vertices * 4, cudaMemcpyHostToDevice, sc->stream));
    }
    // topology check (the reference logs and bails on malformed input, geometry_processor.cpp:606-612)
    if (sc->n_segments) {
        uint32_t* d_bad = reinterpret_cast<uint32_t*>(sc->d_counters + 7);
        uint32_t n_idx = sc->n_segments * 2;
        validate_indices_kernel<<<(n_idx + 255) / 256, 256, 0, sc->stream>>>(sc->d_indices, n_idx, sc->n_vertices, d_bad);
        count_launch();
        uint32_t bad = 0;
        VK_TRY(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, sc->stream));
        VK_TRY(cudaStreamSynchronize(sc->stream));
        if (bad) { set_last_error("line index out of range"); free_scene(sc); return VKHRT_ERR_BAD_TOPOLOGY; }
        VK_TRY(cudaMemsetAsync(sc->d_counters, 0, 16 * sizeof(unsigned long long), sc->stream));
    }
    VK_TRY(cudaStreamSynchronize(sc->stream));   // inputs are copied before return
#undef VK_TRY
    *out_scene = reinterpret_cast<VkhrtScene*>(sc);
    return VKHRT_OK;
}

int vkhrt_scene_build(VkhrtScene* scene)
{
    if (!scene) { set_last_error("null scene"); return VKHRT_ERR_INVALID_ARGUMENT; }
    return build_scene(scene->s, false);
}

int vkhrt_scene_refit(VkhrtScene* scene, const float* positions_xyz)
{
    if (!scene || !positions_xyz) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    DeviceScene& sc = scene->s;
    if (!sc.built) { set_last_error("refit before build"); return VKHRT_ERR_NOT_BUILT; }
    if (sc.lod_applied) { set_last_error("refit after vkhrt_scene_apply_lod: the caller's vertex list no longer describes the scene"); return VKHRT_ERR_UNSUPPORTED; }
    VK_CUDA(cudaSetDevice(sc.device));
    if (sc.n_vertices) VK_CUDA(cudaMemcpyAsync(sc.d_positions, positions_xyz, (size_t)sc.n_vertices * 12, cudaMemcpyHostToDevice, sc.stream));
    return build_scene(sc, true);
}

int vkhrt_scene_get_bvh(VkhrtScene* scene, VkhrtBvhView* view)
{
    if (!scene || !view) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    DeviceScene& sc = scene->s;
    if (!sc.built) { set_last_error("get_bvh before build"); return VKHRT_ERR_NOT_BUILT; }
    VK_CUDA(cudaSetDevice(sc.device));
    view->n_primitives = sc.n_leaves;
    view->n_nodes = sc.n_nodes;
    for (int k = 0; k < 3; ++k) { view->scene_lo[k] = sc.scene_lo[k]; view->scene_hi[k] = sc.scene_hi[k]; }
    if (sc.n_leaves == 0) return VKHRT_OK;
    if (view->nodes) VK_CUDA(cudaMemcpy(view->nodes, sc.d_nodes, (size_t)sc.n_nodes * 64, cudaMemcpyDeviceToHost));
    if (view->sorted_prim_ids) VK_CUDA(cudaMemcpy(view->sorted_prim_ids, sc.d_sorted_ids, (size_t)sc.n_leaves * 4, cudaMemcpyDeviceToHost));
    if (view->sorted_morton) VK_CUDA(cudaMemcpy(view->sorted_morton, sc.d_sorted_morton, (size_t)sc.n_leaves * 8, cudaMemcpyDeviceToHost));
    return VKHRT_OK;
}

int vkhrt_scene_get_primitives(VkhrtScene* scene, float* out, size_t out_floats)
{
    if (!scene || !out) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    return export_primitives(scene->s, out, out_floats);
}

uint32_t vkhrt_scene_primitive_count(const VkhrtScene* scene) { return scene ? scene->s.n_prims : 0; }
uint32_t vkhrt_scene_segment_count(const VkhrtScene* scene) { return scene ? scene->s.n_segments : 0; }

int vkhrt_scene_apply_lod(VkhrtScene* scene, uint32_t line_split_passes, uint32_t line_merge_passes, uint32_t curve_merge_passes)
{
    if (!scene) { set_last_error("null scene"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (scene->s.mesh_first.size() > 1) { set_last_error("vkhrt_scene_apply_lod on a multi-mesh scene: the passes renumber segments across mesh boundaries"); return VKHRT_ERR_UNSUPPORTED; }
    return apply_lod(scene->s, line_split_passes, line_merge_passes, curve_merge_passes);
}

int vkhrt_scene_get_lines(VkhrtScene* scene, float* out, size_t out_floats)
{
    if (!scene || !out) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    return export_lines(scene->s, out, out_floats);
}

int vkhrt_scene_set_environment(VkhrtScene* scene, const float* rgba32f, uint32_t width, uint32_t height)
{
    if (!scene) { set_last_error("null scene"); return VKHRT_ERR_INVALID_ARGUMENT; }
    DeviceScene& sc = scene->s;
    VK_CUDA(cudaSetDevice(sc.device));
    VK_CUDA(cudaStreamSynchronize(sc.stream));
    cudaFree(sc.d_env); sc.d_env = nullptr; sc.env_w = sc.env_h = 0;
    if (!rgba32f || !width || !height) return VKHRT_OK;
    if (width > 32768u || height > 32768u) { set_last_error("environment map larger than 32768 texels on a side"); return VKHRT_ERR_UNSUPPORTED; }
    const size_t bytes = (size_t)width * height * sizeof(float4);
    VK_CUDA(cudaMalloc(&sc.d_env, bytes));
    VK_CUDA(cudaMemcpy(sc.d_env, rgba32f, bytes, cudaMemcpyHostToDevice));
    sc.env_w = width; sc.env_h = height;
    return VKHRT_OK;
}

// albedo = albedoFactor * texture(albedoMap, vec2(0)) (triangle_closest_hit.rchit:77-81 with the zero UVs of hair primitives)
static int evaluate_albedo(const VkhrtMaterial* material, float a[4])
{
    a[0] = a[1] = a[2] = a[3] = 1.0f;
    if (!material) return VKHRT_OK;
    for (int k = 0; k < 4; ++k) a[k] = material->albedo_factor[k];
    if (material->albedo_map_rgba32f) {
        const uint32_t W = material->albedo_map_width, H = material->albedo_map_height;
        if (!W || !H) { set_last_error("albedo map without a size"); return VKHRT_ERR_INVALID_ARGUMENT; }
        // texture(albedoMap, vec2(0)): texel coordinate -0.5 -> texels W-1 | 0 and H-1 | 0 with weights 1/2 (linear, repeat)
        const float* m = material->albedo_map_rgba32f;
        const size_t i0 = W - 1, i1 = 0, j0 = H - 1, j1 = 0;
        for (int k = 0; k < 4; ++k) {
            const float t00 = m[4 * (j0 * W + i0) + k], t10 = m[4 * (j0 * W + i1) + k], t01 = m[4 * (j1 * W + i0) + k], t11 = m[4 * (j1 * W + i1) + k];
            const float top = std::fmaf(0.5f, t10 - t00, t00), bot = std::fmaf(0.5f, t11 - t01, t01);
            a[k] *= std::fmaf(0.5f, bot - top, top);
        }
    }
    return VKHRT_OK;
}

int vkhrt_scene_set_material(VkhrtScene* scene, const VkhrtMaterial* material)
{
    if (!scene) { set_last_error("null scene"); return VKHRT_ERR_INVALID_ARGUMENT; }
    DeviceScene& sc = scene->s;
    float a[4];
    if (int rc = evaluate_albedo(material, a)) return rc;
    std::memcpy(sc.albedo, a, sizeof(a));
    for (size_t m = 0; m < sc.mesh_first.size(); ++m) std::memcpy(&sc.mesh_albedo[4 * m], a, sizeof(a));     // every mesh
    sc.mesh_table_dirty = !sc.mesh_first.empty();
    return VKHRT_OK;
}

int vkhrt_scene_set_meshes(VkhrtScene* scene, const uint32_t* first_segment, uint32_t n_meshes)
{
    if (!scene) { set_last_error("null scene"); return VKHRT_ERR_INVALID_ARGUMENT; }
    DeviceScene& sc = scene->s;
    if (sc.lod_applied) { set_last_error("vkhrt_scene_set_meshes after vkhrt_scene_apply_lod: the segment numbering changed"); return VKHRT_ERR_UNSUPPORTED; }
    if (n_meshes && !first_segment) { set_last_error("null first_segment"); return VKHRT_ERR_INVALID_ARGUMENT; }
    for (uint32_t m = 0; m < n_meshes; ++m) {
        const bool ok = m == 0 ? first_segment[0] == 0u : first_segment[m] >= first_segment[m - 1];
        if (!ok || first_segment[m] > sc.n_segments) { set_last_error("vkhrt_scene_set_meshes: first_segment must start at 0, ascend and stay <= n_segments"); return VKHRT_ERR_INVALID_ARGUMENT; }
    }
    sc.mesh_first.assign(first_segment, first_segment + n_meshes);
    sc.mesh_albedo.resize(4 * (size_t)n_meshes);
    for (uint32_t m = 0; m < n_meshes; ++m) std::memcpy(&sc.mesh_albedo[4 * (size_t)m], sc.albedo, sizeof(sc.albedo));
    sc.mesh_table_dirty = true;
    return VKHRT_OK;
}

int vkhrt_scene_set_mesh_material(VkhrtScene* scene, uint32_t mesh, const VkhrtMaterial* material)
{
    if (!scene) { set_last_error("null scene"); return VKHRT_ERR_INVALID_ARGUMENT; }
    DeviceScene& sc = scene->s;
    if (mesh >= sc.mesh_first.size()) { set_last_error("vkhrt_scene_set_mesh_material: no such mesh (vkhrt_scene_set_meshes first)"); return VKHRT_ERR_INVALID_ARGUMENT; }
    float a[4];
    if (int rc = evaluate_albedo(material, a)) return rc;
    std::memcpy(&sc.mesh_albedo[4 * (size_t)mesh], a, sizeof(a));
    if (sc.mesh_first.size() == 1) std::memcpy(sc.albedo, a, sizeof(a));          // a table of one mesh is the scene's material
    sc.mesh_table_dirty = true;
    return VKHRT_OK;
}

uint32_t vkhrt_scene_mesh_count(const VkhrtScene* scene)
{
    if (!scene) return 0u;
    return scene->s.mesh_first.empty() ? 1u : (uint32_t)scene->s.mesh_first.size();
}

int vkhrt_scene_mesh_of_segments(const VkhrtScene* scene, const uint32_t* segments, uint32_t* mesh_out, size_t n)
{
    if (!scene || (n && (!segments || !mesh_out))) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    const std::vector<uint32_t>& first = scene->s.mesh_first;
    for (size_t i = 0; i < n; ++i) {
        if (segments[i] >= scene->s.n_segments) { mesh_out[i] = VKHRT_MISS_SEGMENT; continue; }      // a miss record, or not a segment
        mesh_out[i] = first.empty() ? 0u : (uint32_t)(std::upper_bound(first.begin(), first.end(), segments[i]) - first.begin()) - 1u;
    }
    return VKHRT_OK;
}

void vkhrt_scene_destroy(VkhrtScene* scene)
{
    if (scene) free_scene(reinterpret_cast<DeviceScene*>(scene));
}

int vkhrt_render(VkhrtScene* scene, const VkhrtFrameDesc* frame, VkhrtHit* hits_out, uint8_t* rgba8_out)
{
    if (!scene || !frame) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (!scene->s.built) { set_last_error("render before build"); return VKHRT_ERR_NOT_BUILT; }
    return render_frame(scene->s, *frame, hits_out, rgba8_out, nullptr);
}

static int wait_oldest(DeviceScene& sc)
{
    DeviceScene::InFlight& f = sc.fl[sc.fl_head];
    VK_CUDA(cudaSetDevice(sc.device));
    VK_CUDA(cudaEventSynchronize(f.copied));
    f.pending = false;
    sc.fl_head = (sc.fl_head + 1u) % VKHRT_FRAMES_IN_FLIGHT;
    sc.fl_count--;
    return VKHRT_OK;
}

int vkhrt_render_submit(VkhrtScene* scene, const VkhrtFrameDesc* frame, VkhrtHit* hits_out, uint8_t* rgba8_out)
{
    if (!scene || !frame) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    DeviceScene& sc = scene->s;
    if (!sc.built) { set_last_error("render before build"); return VKHRT_ERR_NOT_BUILT; }
    if (frame->output_memory != VKHRT_MEM_HOST) { set_last_error("vkhrt_render_submit: host output buffers (device outputs are asynchronous through VkhrtFrameDesc::stream already)"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (frame->tile_stride > 1 && frame->row_major_output) { set_last_error("vkhrt_render_submit: a shard of a shared frame cannot be copied out as a whole buffer"); return VKHRT_ERR_INVALID_ARGUMENT; }
    const uint64_t n_out = frame_local_pixels(*frame);
    if (n_out == 0) { set_last_error("vkhrt_render_submit: bad frame description"); return VKHRT_ERR_INVALID_ARGUMENT; }
    int rc;
    if (sc.fl_count == VKHRT_FRAMES_IN_FLIGHT && (rc = wait_oldest(sc))) return rc;       // back-pressure: the oldest frame's buffers are needed again
    VK_CUDA(cudaSetDevice(sc.device));
    DeviceScene::InFlight& f = sc.fl[(sc.fl_head + sc.fl_count) % VKHRT_FRAMES_IN_FLIGHT];
    if (!f.traced) { VK_CUDA(cudaEventCreateWithFlags(&f.traced, cudaEventDisableTiming)); VK_CUDA(cudaEventCreateWithFlags(&f.copied, cudaEventDisableTiming)); }
    if (hits_out && f.hits_n < n_out) { cudaFree(f.d_hits); f.d_hits = nullptr; f.hits_n = 0; VK_CUDA(cudaMalloc((void**)&f.d_hits, (size_t)n_out * sizeof(VkhrtHit))); f.hits_n = (size_t)n_out; }
    if (rgba8_out && f.rgba_n < n_out * 4) { cudaFree(f.d_rgba); f.d_rgba = nullptr; f.rgba_n = 0; VK_CUDA(cudaMalloc((void**)&f.d_rgba, (size_t)n_out * 4)); f.rgba_n = (size_t)n_out * 4; }
    // the traversal writes into this slot's device buffers (its previous copy has completed: the slot was waited for); the copy
    // engine takes them to the host on its own stream, so the NEXT frame's kernels do not wait for the copy
    RenderOpts o;
    o.defer_sync = true; o.hits_on_device = true; o.rgba_on_device = true;
    rc = render_frame(sc, *frame, hits_out ? f.d_hits : nullptr, rgba8_out ? f.d_rgba : nullptr, nullptr, o);
    if (rc) return rc;
    VK_CUDA(cudaEventRecord(f.traced, sc.stream));
    VK_CUDA(cudaStreamWaitEvent(sc.copy_stream, f.traced, 0));
    if (hits_out) VK_CUDA(cudaMemcpyAsync(hits_out, f.d_hits, (size_t)n_out * sizeof(VkhrtHit), cudaMemcpyDeviceToHost, sc.copy_stream));
    if (rgba8_out) VK_CUDA(cudaMemcpyAsync(rgba8_out, f.d_rgba, (size_t)n_out * 4, cudaMemcpyDeviceToHost, sc.copy_stream));
    VK_CUDA(cudaEventRecord(f.copied, sc.copy_stream));
    f.pending = true;
    sc.fl_count++;
    return VKHRT_OK;
}

int vkhrt_render_wait(VkhrtScene* scene)
{
    if (!scene) { set_last_error("null scene"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (scene->s.fl_count == 0) { set_last_error("vkhrt_render_wait: no frame is outstanding"); return VKHRT_ERR_INVALID_ARGUMENT; }
    return wait_oldest(scene->s);
}

int vkhrt_render_stats(VkhrtScene* scene, const VkhrtFrameDesc* frame, VkhrtHit* hits_out, uint8_t* rgba8_out, VkhrtTraceStats* stats)
{
    if (!scene || !frame || !stats) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (!scene->s.built) { set_last_error("render before build"); return VKHRT_ERR_NOT_BUILT; }
    return render_frame(scene->s, *frame, hits_out, rgba8_out, stats);
}

uint64_t vkhrt_frame_local_pixels(const VkhrtFrameDesc* frame) { return frame ? frame_local_pixels(*frame) : 0; }

int vkhrt_untile(const VkhrtFrameDesc* frame, uint32_t world, const void* gathered, void* row_major, uint32_t elem_bytes, void* stream)
{
    if (!frame || !gathered || !row_major) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    return untile_buffer(*frame, world, gathered, row_major, elem_bytes, (cudaStream_t)stream);
}

// compact (tile, pixel-in-tile) shards concatenated rank-major -> row-major, on the host (same index arithmetic as untile_kernel)
static int untile_host(const VkhrtFrameDesc& f, uint32_t world, const unsigned char* const* shards, unsigned char* row_major, uint32_t elem_bytes)
{
    const uint32_t T = f.tile_size ? f.tile_size : 64;
    if (world == 0 || f.width == 0 || f.height == 0 || T % 8) { set_last_error("untile: bad frame description"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (elem_bytes != 4 && elem_bytes != 32) { set_last_error("untile: elem_bytes must be 4 or 32"); return VKHRT_ERR_INVALID_ARGUMENT; }
    const uint32_t tiles_x = (f.width + T - 1) / T;
    // a tile row is contiguous on both sides: one copy per (image row, tile column)
    for (uint32_t py = 0; py < f.height; ++py)
        for (uint32_t tx = 0; tx < tiles_x; ++tx) {
            const uint32_t px = tx * T, npx = std::min(T, f.width - px);
            const uint64_t tile = (uint64_t)(py / T) * tiles_x + tx;
            const uint64_t local = (tile / world) * T * T + (uint64_t)(py % T) * T;
            std::memcpy(row_major + ((uint64_t)py * f.width + px) * elem_bytes, shards[tile % world] + local * elem_bytes, (size_t)npx * elem_bytes);
        }
    return VKHRT_OK;
}

int vkhrt_untile_host(const VkhrtFrameDesc* frame, uint32_t world, const void* gathered, void* row_major, uint32_t elem_bytes)
{
    if (!frame || !gathered || !row_major || world == 0) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (elem_bytes != 4 && elem_bytes != 32) { set_last_error("untile: elem_bytes must be 4 or 32"); return VKHRT_ERR_INVALID_ARGUMENT; }
    VkhrtFrameDesc shard = *frame;
    shard.tile_first = 0; shard.tile_stride = world; shard.row_major_output = 0;
    const uint64_t per_shard = world > 1 ? frame_local_pixels(shard) : (uint64_t)frame->width * frame->height;
    if (world == 1) { std::memcpy(row_major, gathered, (size_t)(per_shard * elem_bytes)); return VKHRT_OK; }
    std::vector<const unsigned char*> shards(world);
    for (uint32_t r = 0; r < world; ++r) shards[r] = (const unsigned char*)gathered + (uint64_t)r * per_shard * elem_bytes;
    return untile_host(*frame, world, shards.data(), (unsigned char*)row_major, elem_bytes);
}

// One frame on several GPUs from one process.  Two ways to assemble it, both without a CPU re-ordering pass:
//   * hit records, page-locked caller buffer: every GPU's traversal kernel stores its records at their row-major position
//     straight into the caller's buffer over its own PCIe link (the zero-copy / line-wise delivery of vkhrt_render);
//   * everything else (pixels; records into pageable memory): scenes[0]'s GPU holds a full-frame buffer, the other GPUs store
//     into it over NVLink (peer access, VkhrtFrameDesc::row_major_output: the kernels' own stores are the gather), and one
//     copy brings it to the host.
// Shards are launched by persistent worker threads (one per extra scene) so that no GPU waits for another one's launch.
// Without peer access between the devices the shards are staged through pinned buffers and re-ordered on the host.
}  // extern "C"
namespace vkhrt {
namespace {
class Workers {
public:
    // run fn(0) .. fn(n-1): fn(0) on the calling thread, the others on persistent threads
    void run(uint32_t n, const std::function<void(uint32_t)>& fn)
    {
        std::lock_guard<std::mutex> serial(call_);
        while (threads_.size() + 1 < n) { const uint32_t id = (uint32_t)threads_.size() + 1; threads_.emplace_back([this, id] { loop(id); }); }
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn; n_ = n; pending_ = n - 1; ++generation_;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
    ~Workers()
    {
        { std::lock_guard<std::mutex> g(m_); quit_ = true; }
        cv_.notify_all();
        for (std::thread& t : threads_) t.join();
    }
private:
    void loop(uint32_t id)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(uint32_t)>* fn = nullptr;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return quit_ || (generation_ != seen && id < n_); });
                if (quit_) return;
                seen = generation_; fn = fn_;
            }
            (*fn)(id);
            { std::lock_guard<std::mutex> g(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    std::mutex call_, m_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> threads_;
    const std::function<void(uint32_t)>* fn_ = nullptr;
    uint32_t n_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool quit_ = false;
};
Workers& workers() { static Workers w; return w; }

// can every scene's device store into scenes[0]'s device?  (enables peer access on first use)
bool peer_ready(VkhrtScene* const* scenes, uint32_t n)
{
    const int d0 = scenes[0]->s.device;
    for (uint32_t r = 1; r < n; ++r) {
        const int d = scenes[r]->s.device;
        if (d == d0) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, d, d0) != cudaSuccess || !can) { cudaGetLastError(); return false; }
        if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); return false; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(d0, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
        cudaGetLastError();
    }
    return true;
}
template <typename T>
int grow_buffer(T** p, size_t* have, size_t want)
{
    if (*have >= want) return VKHRT_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    VK_CUDA(cudaMalloc((void**)p, want * sizeof(T)));
    *have = want;
    return VKHRT_OK;
}
}  // namespace
}  // namespace vkhrt
extern "C" {

int vkhrt_render_multi(VkhrtScene* const* scenes, uint32_t n_scenes, const VkhrtFrameDesc* frame, VkhrtHit* hits_out, uint8_t* rgba8_out)
{
    if (!scenes || n_scenes == 0 || !frame) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    for (uint32_t r = 0; r < n_scenes; ++r) {
        if (!scenes[r]) { set_last_error("null scene"); return VKHRT_ERR_INVALID_ARGUMENT; }
        if (!scenes[r]->s.built) { set_last_error("render before build"); return VKHRT_ERR_NOT_BUILT; }
        for (uint32_t q = 0; q < r; ++q) if (scenes[q] == scenes[r]) { set_last_error("vkhrt_render_multi: the same scene handle twice (calls on one scene are not re-entrant)"); return VKHRT_ERR_INVALID_ARGUMENT; }
    }
    if (frame->tile_stride > 1 || frame->tile_first != 0 || frame->row_major_output) { set_last_error("vkhrt_render_multi: the frame must describe the whole image"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (frame->output_memory != VKHRT_MEM_HOST) { set_last_error("vkhrt_render_multi: host output buffers only"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (n_scenes == 1) return vkhrt_render(scenes[0], frame, hits_out, rgba8_out);
    VkhrtFrameDesc base = *frame;
    base.stream = nullptr;
    if (base.tile_size == 0) base.tile_size = 64;
    base.tile_stride = n_scenes;
    const uint64_t per_shard = frame_local_pixels(base);
    if (per_shard == 0) { set_last_error("vkhrt_render_multi: bad frame description"); return VKHRT_ERR_INVALID_ARGUMENT; }
    const size_t n_full = (size_t)frame->width * frame->height;
    init_tunables();                       // the environment switches are read once, before the worker threads launch anything
    DeviceScene& s0 = scenes[0]->s;
    std::vector<int> rc(n_scenes, VKHRT_OK);
    std::vector<std::string> err(n_scenes);
    auto first_error = [&]() -> int {
        for (uint32_t r = 0; r < n_scenes; ++r)
            if (rc[r] != VKHRT_OK) { set_last_error("shard " + std::to_string(r) + ": " + err[r]); return rc[r]; }
        return VKHRT_OK;
    };

    if (peer_ready(scenes, n_scenes)) {
        // records: straight into the caller's buffer when the kernels can store into it (page-locked), else via scenes[0]'s GPU
        bool hits_zero_copy = false;
        if (hits_out) {
            cudaPointerAttributes at;
            hits_zero_copy = cudaPointerGetAttributes(&at, hits_out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr;
            cudaGetLastError();
        }
        VK_CUDA(cudaSetDevice(s0.device));
        if (hits_out && !hits_zero_copy) { int g = grow_buffer(&s0.d_multi_hits, &s0.multi_hits_n, n_full); if (g) return g; }
        if (rgba8_out) { int g = grow_buffer(&s0.d_multi_rgba, &s0.multi_rgba_n, n_full * 4); if (g) return g; }
        for (uint32_t r = 0; r < n_scenes; ++r)
            if (!scenes[r]->s.multi_done) { VK_CUDA(cudaSetDevice(scenes[r]->s.device)); VK_CUDA(cudaEventCreateWithFlags(&scenes[r]->s.multi_done, cudaEventDisableTiming)); }
        workers().run(n_scenes, [&](uint32_t r) {
            DeviceScene& sc = scenes[r]->s;
            VkhrtFrameDesc f = base;
            f.tile_first = r;
            f.row_major_output = 1;
            RenderOpts o;
            o.defer_sync = true;
            o.hits_on_device = hits_out && !hits_zero_copy;
            o.rgba_on_device = rgba8_out != nullptr;
            rc[r] = render_frame(sc, f, hits_out ? (hits_zero_copy ? hits_out : s0.d_multi_hits) : nullptr, rgba8_out ? s0.d_multi_rgba : nullptr, nullptr, o);
            if (rc[r] == VKHRT_OK && cudaEventRecord(sc.multi_done, sc.stream) != cudaSuccess) { rc[r] = VKHRT_ERR_CUDA; set_last_error("cudaEventRecord failed"); }
            if (rc[r] != VKHRT_OK) err[r] = vkhrt_last_error();       // the error text is per thread: carry it to the caller's
        });
        // wait for every shard (also after a failure: nothing may still be writing into the buffers when this call returns)
        for (uint32_t r = 0; r < n_scenes; ++r) { cudaSetDevice(scenes[r]->s.device); cudaStreamSynchronize(scenes[r]->s.stream); }
        cudaGetLastError();
        int e = first_error();
        if (e) return e;
        VK_CUDA(cudaSetDevice(s0.device));
        if (hits_out && !hits_zero_copy) VK_CUDA(cudaMemcpyAsync(hits_out, s0.d_multi_hits, n_full * sizeof(VkhrtHit), cudaMemcpyDeviceToHost, s0.stream));
        if (rgba8_out) VK_CUDA(cudaMemcpyAsync(rgba8_out, s0.d_multi_rgba, n_full * 4, cudaMemcpyDeviceToHost, s0.stream));
        VK_CUDA(cudaStreamSynchronize(s0.stream));
        return VKHRT_OK;
    }

    // no peer access between the devices: compact shards through host staging + re-ordering on the host
    std::vector<std::vector<unsigned char>> sh_hits(n_scenes), sh_rgba(n_scenes);
    try {
        for (uint32_t r = 0; r < n_scenes; ++r) {
            if (hits_out) sh_hits[r].resize((size_t)per_shard * sizeof(VkhrtHit));
            if (rgba8_out) sh_rgba[r].resize((size_t)per_shard * 4);
        }
    } catch (const std::bad_alloc&) { set_last_error("out of host memory"); return VKHRT_ERR_OUT_OF_MEMORY; }
    workers().run(n_scenes, [&](uint32_t r) {
        VkhrtFrameDesc f = base;
        f.tile_first = r;
        rc[r] = vkhrt_render(scenes[r], &f, hits_out ? (VkhrtHit*)sh_hits[r].data() : nullptr, rgba8_out ? sh_rgba[r].data() : nullptr);
        if (rc[r] != VKHRT_OK) err[r] = vkhrt_last_error();
    });
    int e = first_error();
    if (e) return e;
    std::vector<const unsigned char*> ptr(n_scenes);
    if (hits_out) {
        for (uint32_t r = 0; r < n_scenes; ++r) ptr[r] = sh_hits[r].data();
        int u = untile_host(*frame, n_scenes, ptr.data(), (unsigned char*)hits_out, (uint32_t)sizeof(VkhrtHit));
        if (u) return u;
    }
    if (rgba8_out) {
        for (uint32_t r = 0; r < n_scenes; ++r) ptr[r] = sh_rgba[r].data();
        int u = untile_host(*frame, n_scenes, ptr.data(), rgba8_out, 4u);
        if (u) return u;
    }
    return VKHRT_OK;
}

int vkhrt_last_timing(const VkhrtScene* scene, VkhrtTiming* timing)
{
    if (!scene || !timing) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    const DeviceScene& sc = scene->s;
    cudaSetDevice(sc.device);
    *timing = sc.timing;
    // frame stages: events 6..11 (recorded by render_frame; zero if no frame has been rendered yet)
    if (cudaEventSynchronize(sc.ev[11]) == cudaSuccess) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sc.ev[7], sc.ev[8]) == cudaSuccess) timing->trace_ms = ms;
        if (cudaEventElapsedTime(&ms, sc.ev[8], sc.ev[12]) == cudaSuccess) timing->ao_ms = ms;
        if (cudaEventElapsedTime(&ms, sc.ev[12], sc.ev[9]) == cudaSuccess) timing->shade_ms = ms;
        if (cudaEventElapsedTime(&ms, sc.ev[6], sc.ev[10]) == cudaSuccess) timing->render_total_ms = ms;
        if (cudaEventElapsedTime(&ms, sc.ev[10], sc.ev[11]) == cudaSuccess) timing->d2h_ms = ms;
    }
    cudaGetLastError();
    return VKHRT_OK;
}

int vkhrt_host_alloc(size_t bytes, void** out_ptr)
{
    if (!out_ptr || bytes == 0) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    *out_ptr = nullptr;
    int rc = check_device(0);
    if (rc) return rc;
    VK_CUDA(cudaHostAlloc(out_ptr, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    return VKHRT_OK;
}

int vkhrt_host_free(void* ptr)
{
    if (!ptr) return VKHRT_OK;
    VK_CUDA(cudaFreeHost(ptr));
    return VKHRT_OK;
}

int vkhrt_shared_buffer_create(int device, size_t bytes, void** dev_ptr_out, uint8_t handle_out[VKHRT_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == VKHRT_IPC_HANDLE_BYTES, "IPC handle size");
    if (!dev_ptr_out || !handle_out || bytes == 0) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    int rc = check_device(device);
    if (rc) return rc;
    VK_CUDA(cudaSetDevice(device));
    void* ptr = nullptr;
    VK_CUDA(cudaMalloc(&ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) { cudaFree(ptr); set_last_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); return VKHRT_ERR_CUDA; }
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr_out = ptr;
    return VKHRT_OK;
}

int vkhrt_shared_buffer_open(int device, const uint8_t handle[VKHRT_IPC_HANDLE_BYTES], void** dev_ptr_out)
{
    if (!handle || !dev_ptr_out) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    int rc = check_device(device);
    if (rc) return rc;
    VK_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* ptr = nullptr;
    VK_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = ptr;
    return VKHRT_OK;
}

int vkhrt_shared_buffer_close(int device, void* opened_ptr)
{
    if (!opened_ptr) return VKHRT_OK;
    VK_CUDA(cudaSetDevice(device));
    VK_CUDA(cudaIpcCloseMemHandle(opened_ptr));
    return VKHRT_OK;
}

int vkhrt_shared_buffer_destroy(int device, void* created_ptr)
{
    if (!created_ptr) return VKHRT_OK;
    VK_CUDA(cudaSetDevice(device));
    VK_CUDA(cudaFree(created_ptr));
    return VKHRT_OK;
}

int vkhrt_generate_rays(const VkhrtFrameDesc* frame, uint32_t sample, float* rays_out_device, int device)
{
    if (!frame || !rays_out_device) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    int rc = check_device(device);
    if (rc) return rc;
    VK_CUDA(cudaSetDevice(device));
    rc = generate_ray_buffer(*frame, sample, rays_out_device, (cudaStream_t)frame->stream);
    if (rc) return rc;
    if (!frame->stream) VK_CUDA(cudaDeviceSynchronize());
    return VKHRT_OK;
}

int vkhrt_trace_rays(VkhrtScene* scene, const float* rays_device, uint64_t n_rays, VkhrtHit* hits_out_device, void* stream)
{
    if (!scene || (n_rays && (!rays_device || !hits_out_device))) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (!scene->s.built) { set_last_error("trace before build"); return VKHRT_ERR_NOT_BUILT; }
    return trace_ray_buffer(scene->s, rays_device, n_rays, hits_out_device, false, (cudaStream_t)stream);
}

int vkhrt_trace_rays_any_hit(VkhrtScene* scene, const float* rays_device, uint64_t n_rays, VkhrtHit* hits_out_device, void* stream)
{
    if (!scene || (n_rays && (!rays_device || !hits_out_device))) { set_last_error("null argument"); return VKHRT_ERR_INVALID_ARGUMENT; }
    if (!scene->s.built) { set_last_error("trace before build"); return VKHRT_ERR_NOT_BUILT; }
    return trace_ray_buffer(scene->s, rays_device, n_rays, hits_out_device, true, (cudaStream_t)stream);
}

}  // extern "C"
