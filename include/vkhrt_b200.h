/*
 * vkhrt_b200.h — C ABI of the B200-native hair ray-tracing hot path.
 *
 * This is the drop-in boundary for the ONE data-parallel path of mmzala/vkhrt
 * that this repository replaces: primary-ray generation -> BVH traversal over
 * per-segment hair primitives -> ray/segment intersection (Phantom / LSS / DOTS)
 * -> closest-hit shading.  The reference exposes no FFI of its own (it is one
 * Windows/Vulkan executable, reference source/main.cpp:3-7); the entry points
 * below sit at the three seams of the reference where data crosses from the
 * host into the Vulkan ray-tracing pipeline.  Each declaration cites the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers + sizes only; no C++/torch types cross this boundary;
 *   - every call returns int: 0 = VKHRT_OK, negative = VkhrtStatus error.
 *     Nothing aborts or throws across the ABI (the reference abort()s on
 *     Vulkan errors, source/vk_common.cpp:5-15);
 *   - calls on one scene are not re-entrant (the reference is single-threaded,
 *     one graphics queue; source/renderer.cpp:83-131);
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns VKHRT_ERR_NO_DEVICE.
 */
#ifndef VKHRT_B200_H
#define VKHRT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKHRT_ABI_VERSION 6

typedef enum VkhrtStatus {
    VKHRT_OK = 0,
    VKHRT_ERR_INVALID_ARGUMENT = -1,
    VKHRT_ERR_NO_DEVICE = -2,      /* no CUDA device / driver: there is no CPU path      */
    VKHRT_ERR_CUDA = -3,           /* a CUDA runtime call failed (see vkhrt_last_error)  */
    VKHRT_ERR_OUT_OF_MEMORY = -4,
    VKHRT_ERR_NOT_BUILT = -5,      /* render/refit/get_bvh before vkhrt_scene_build      */
    VKHRT_ERR_BAD_TOPOLOGY = -6,   /* index out of range (reference logs and returns the
                                      input unchanged, geometry_processor.cpp:606-612)   */
    VKHRT_ERR_UNSUPPORTED = -7,
    VKHRT_ERR_IO = -8              /* asset / image file could not be read, parsed or written (the reference logs
                                      and returns nullptr, source/resources/model/model_loader.cpp:280-284)      */
} VkhrtStatus;

/* Technique = which ProcessHair* generator of the reference builds the primitives
 * (include/resources/model/geometry_processor.hpp:4-7; the reference picks one by a
 * source edit at source/resources/model/model_loader.cpp:334). */
typedef enum VkhrtTechnique {
    VKHRT_TECHNIQUE_PHANTOM = 0,   /* ProcessHairCurves: Catmull-Rom->Bezier curves + AABBs, Phantom
                                      Ray-Hair Intersector (shaders/hair_intersection.rint)          */
    VKHRT_TECHNIQUE_LSS = 1,       /* ProcessHairLSS: linear swept spheres (renderer.cpp:653-692)    */
    VKHRT_TECHNIQUE_DOTS = 2       /* ProcessHairDOTS: disjoint orthogonal triangle strips           */
} VkhrtTechnique;

/* Closest-hit output colour: shaders/shading.glsl:1-11 or shaders/debug.glsl:1-7 (the
 * reference computes Shade() and then displays the debug colour,
 * shaders/hair_closest_hit.rchit:21-24). */
typedef enum VkhrtShadeMode {
    VKHRT_SHADE = 0,
    VKHRT_SHADE_DEBUG_PRIMID = 1,
    VKHRT_SHADE_MATERIAL = 2       /* Shade(normal) * albedo, albedo = material.albedoFactor [* albedo map]: the term
                                      shaders/triangle_closest_hit.rchit:77-81 computes (and the reference then overwrites
                                      with the debug colour); see vkhrt_scene_set_material */
} VkhrtShadeMode;

/* Colour of a ray that hits nothing.  ENVIRONMENT = shaders/miss.rmiss:17-38: equirectangular lookup of
 * normalize(-rayDirection) in the scene's environment map (linear filter, repeat addressing, the sampler defaults of
 * include/resources/gpu_resources.hpp:46-50), then 1 - exp(-c) and gamma 1/2.2. */
typedef enum VkhrtMissMode {
    VKHRT_MISS_CONSTANT = 0,
    VKHRT_MISS_ENVIRONMENT = 1
} VkhrtMissMode;

/* Where the caller's output pointers live. */
typedef enum VkhrtMemory {
    VKHRT_MEM_HOST = 0,
    VKHRT_MEM_DEVICE = 1
} VkhrtMemory;

#define VKHRT_DEFAULT_RADIUS 0.02f   /* shaders/hair_intersection.rint:18, geometry_processor.cpp:635,681,765 */
#define VKHRT_DEFAULT_T_MIN 0.001f   /* shaders/ray_gen.rgen:27 */
#define VKHRT_DEFAULT_T_MAX 10000.0f /* shaders/ray_gen.rgen:28 */
#define VKHRT_MISS_SEGMENT 0xFFFFFFFFu

/* Scene input = the Assimp line mesh as consumed by GenerateLines
 * (source/resources/model/geometry_processor.cpp:45-67; produced by ProcessMesh,
 * source/resources/model/model_loader.cpp:139-206): vertex positions + uint32 index
 * pairs; consecutive segments of one strand share an identical end/start position. */
typedef struct VkhrtSceneDesc {
    const float*    positions_xyz;      /* n_vertices * 3 floats, world space                      */
    uint32_t        n_vertices;
    const uint32_t* line_indices;       /* n_segments * 2 vertex indices                           */
    uint32_t        n_segments;
    const float*    radius_per_vertex;  /* nullable: n_vertices radii, linear along each segment.  PHANTOM: cone radius(t) and slant
                                           (shaders/cone.glsl:27, which the reference's caller leaves at 0); LSS: end-sphere radii;
                                           DOTS: per-end strip half-widths.  NULL => `radius` everywhere (the reference's 0.02)  */
    float           radius;             /* <= 0 => VKHRT_DEFAULT_RADIUS                            */
    int32_t         technique;          /* VkhrtTechnique                                          */
    int32_t         device;             /* CUDA device ordinal                                     */
} VkhrtSceneDesc;

/* Per-frame input = CameraUniformData (include/resources/camera_resource.hpp:7-11) +
 * the vkCmdTraceRaysKHR(width,height,1) launch (source/renderer.cpp:156-166) + the ray
 * interval of shaders/ray_gen.rgen:27-28.  Matrices are column-major float[16] exactly
 * as glm::mat4 is memcpy'd into the UBO (source/resources/camera_resource.cpp:16-22). */
typedef struct VkhrtFrameDesc {
    float    view_inverse[16];
    float    proj_inverse[16];
    uint32_t width, height;
    float    t_min, t_max;          /* 0,0 => defaults 0.001 / 10000                               */
    uint32_t spp;                   /* samples per pixel; 0 => 1. Sample 0 is the pixel centre
                                       (the reference traces exactly that one, ray_gen.rgen:18)    */
    int32_t  shade_mode;            /* VkhrtShadeMode                                              */
    float    miss_rgb[3];           /* constant miss colour (miss_mode == VKHRT_MISS_CONSTANT)     */
    /* sharding (all zero => the whole frame): the frame is cut into tile_size^2-pixel tiles
     * numbered row-major; this call traces tiles tile_first, tile_first+tile_stride, ...
     * and writes its outputs COMPACTLY in (tile, pixel-in-tile) order when tile_stride > 1 */
    uint32_t tile_size;             /* 0 => 64 (must be a multiple of 8)                           */
    uint32_t tile_first;
    uint32_t tile_stride;           /* 0 => 1                                                      */
    uint32_t row_major_output;      /* with tile_stride > 1: 1 => outputs are FULL-FRAME buffers and every pixel
                                       is written at its row-major position (only the tiles of this shard are
                                       touched).  All shards can then share one frame buffer, e.g. the gathering
                                       GPU's, mapped into every rank over NVLink (vkhrt_shared_buffer_*): the
                                       traversal kernel's stores ARE the gather.  Device output memory, or (hit
                                       records only) a PAGE-LOCKED host buffer the kernels store into directly. */
    int32_t  output_memory;         /* VkhrtMemory of hits_out / rgba8_out                         */
    void*    stream;                /* cudaStream_t to run on (device outputs only); NULL => the
                                       scene's own stream                                         */
    /* secondary rays (not in the reference, whose closest-hit shaders never trace,
     * shaders/hair_closest_hit.rchit:15-25).  ao_samples > 0: every primary hit spawns ao_samples
     * ambient-occlusion rays (cosine-weighted about the shading normal, terminate-on-first-hit,
     * through the same traversal kernel) and the pixel colour is multiplied by the unoccluded
     * fraction.  Only the image is affected; hit records stay the primary hits. */
    uint32_t ao_samples;
    float    ao_distance;           /* <= 0 => VKHRT_DEFAULT_AO_DISTANCE                           */
    float    ao_bias;               /* origin offset along the normal; <= 0 => 0.25 * radius       */
    int32_t  miss_mode;             /* VkhrtMissMode; ENVIRONMENT needs vkhrt_scene_set_environment */
} VkhrtFrameDesc;

#define VKHRT_DEFAULT_AO_DISTANCE 2.0f
#define VKHRT_AO_T_MIN 1e-4f

/* Per-ray hit record (sample 0 of each pixel).  The reference exports only t and a
 * normal (hitAttributeEXT, shaders/hair_intersection.rint:13,148) and the primitive id
 * (gl_PrimitiveID); `u` (the converged curve parameter) and `segment` are added here. */
typedef struct VkhrtHit {
    float    t;          /* hit distance; miss: +inf                                               */
    uint32_t segment;    /* input segment id (primitive/4 for DOTS); miss: VKHRT_MISS_SEGMENT      */
    float    u;          /* curve / segment parameter in [0,1]                                     */
    float    nx, ny, nz; /* unit shading normal                                                    */
    uint32_t primitive;  /* gl_PrimitiveID equivalent (curve, LSS or triangle index)               */
    uint32_t flags;      /* bit0 = hit                                                             */
} VkhrtHit;

/* One LBVH node: two children with their boxes, 64 bytes, read as 4 x 16-byte loads.
 * child >> 31 == 1 => leaf; low 31 bits = position in Morton-sorted leaf order
 * (leaf) or internal-node index.  primK = original leaf id when child K is a leaf.
 * A leaf is one PIECE of a primitive group (VKHRT_LEAF_SPLIT_* below): group = PHANTOM curve, LSS, or for DOTS the STRIP = the
 * 4 triangles 4*segment .. 4*segment+3 of a segment; leaf id = group * K + piece, group = segment id. */
typedef struct VkhrtBvhNode {
    float lo0[3]; uint32_t child0;
    float hi0[3]; uint32_t child1;
    float lo1[3]; uint32_t prim0;
    float hi1[3]; uint32_t prim1;
} VkhrtBvhNode;

/* Leaves per primitive group.  A group = one PHANTOM curve, one LSS, or the 4-triangle DOTS strip of a segment.  The AABB of a
 * thin diagonal segment is mostly empty, so a group is cut into K pieces along its length and every piece becomes a BVH leaf of
 * its own (leaf id = group * K + piece) whose box bounds only that piece; the leaf still refers to the whole group, which is
 * tested as before (a group reached through two of its leaves is simply tested twice: closest-hit selection is idempotent).
 * C2: 57.6 -> 52.1 node visits and 5.9 -> 3.9 curve tests per ray; a 4 M-segment DOTS groom: 93 -> 52 and 15.2 -> 4.3 strips. */
#ifndef VKHRT_LEAF_SPLIT_PHANTOM          /* overridable only for experiment builds (tools/variants); the oracle follows the same macros */
#define VKHRT_LEAF_SPLIT_PHANTOM 2
#endif
#ifndef VKHRT_LEAF_SPLIT_LSS
#define VKHRT_LEAF_SPLIT_LSS 2
#endif
#ifndef VKHRT_LEAF_SPLIT_DOTS
#define VKHRT_LEAF_SPLIT_DOTS 4
#endif

/* One-entry mailbox per ray: the group a ray tested LAST is not tested again when the next leaf the ray reaches is another
 * piece of that same group (consecutive pieces of a curve / strip are what a ray along the strand meets).  Result-neutral:
 * the repeated test would return the same (t, u) and closest-hit selection is idempotent; only the traversal counters change
 * (a skipped leaf is not counted as a primitive test).  Part of the traversal definition, so the oracle applies it too.
 * PHANTOM only: there a repeated test is a repeated march (C2, per ray: 3.86 -> 3.70 curve tests and 16.2 -> 15.4 cone iterations
 * with 2 pieces per curve, 3.28 -> 2.75 and 17.7 -> 14.5 with 4).  LSS has no mailbox (its group id is not in the leaf record and
 * the test is cheaper than the extra fetch), DOTS neither (strip pieces are rarely met back to back: 4.28 -> 4.18 strips). */
#ifndef VKHRT_MAILBOX_PHANTOM
#define VKHRT_MAILBOX_PHANTOM 1
#endif
#define VKHRT_MAILBOX_LSS 0
#define VKHRT_MAILBOX_DOTS 0

#define VKHRT_BVH_LEAF 0x80000000u
#define VKHRT_BVH_EMPTY 0xFFFFFFFFu

/* Host copy-out of the acceleration structure for the bit-exact build check (the
 * reference's BLAS is opaque driver state, source/bottom_level_acceleration_structure.cpp:34-78). */
typedef struct VkhrtBvhView {
    uint32_t      n_primitives;     /* number of BVH leaves (= segments * VKHRT_LEAF_SPLIT_*)      */
    uint32_t      n_nodes;          /* max(n_primitives - 1, 1)                                    */
    VkhrtBvhNode* nodes;            /* caller-allocated, n_nodes entries (nullable)                */
    uint32_t*     sorted_prim_ids;  /* caller-allocated, n_primitives entries (nullable): leaf ids */
    uint64_t*     sorted_morton;    /* caller-allocated, n_primitives entries (nullable)           */
    float         scene_lo[3], scene_hi[3];  /* centroid bounds used for Morton quantisation       */
} VkhrtBvhView;

/* CUDA-event milliseconds of the last call, per stage. */
typedef struct VkhrtTiming {
    float geometry_ms;   /* GenerateLines/Curves/AABBs/DOTS/LSS kernels                            */
    float morton_ms, sort_ms, hierarchy_ms, refit_ms;
    float build_total_ms;
    float raygen_ms, trace_ms, shade_ms, render_total_ms;
    float h2d_ms, d2h_ms;
    float ao_ms;         /* ambient-occlusion passes of sample 0 (0 when ao_samples == 0)          */
    float lod_ms;        /* vkhrt_scene_apply_lod: all passes, incl. the host reads of the compacted sizes */
} VkhrtTiming;

/* Per-frame traversal statistics (debug counters; filled only by vkhrt_render_stats). */
typedef struct VkhrtTraceStats {
    uint64_t rays;
    uint64_t nodes_visited;     /* internal 64-byte node records fetched                           */
    uint64_t prims_tested;      /* primitives of the leaves reached (DOTS: 4 per strip fetched)    */
    uint64_t hits;
    uint64_t phantom_iterations;
    /* warp scheduler: steps executed per state {node, leaf, march, refill} and the lanes active in them
     * (lanes / (32 * steps) = SIMD occupancy of that state) */
    uint64_t sched_steps[4];
    uint64_t sched_lanes[4];
} VkhrtTraceStats;

typedef struct VkhrtScene VkhrtScene;

/* ---- library -------------------------------------------------------------------------- */
int         vkhrt_abi_version(void);
int         vkhrt_device_count(void);                 /* 0 when no usable CUDA device      */
const char* vkhrt_error_string(int status);
const char* vkhrt_last_error(void);                   /* detail text of the last failure   */
uint64_t    vkhrt_launch_count(void);                 /* kernels launched by this library  */

/* ---- scene = ModelLoader::LoadFromFile + ProcessHair{Curves,LSS,DOTS} + Model upload ---- */
/* replaces include/resources/model/model_loader.hpp:20, geometry_processor.hpp:4-7,
 * source/resources/model/model.cpp:52-223.  Input arrays are copied before return. */
int  vkhrt_scene_create(const VkhrtSceneDesc* desc, VkhrtScene** out_scene);
/* replaces BottomLevelAccelerationStructure ctor + TopLevelAccelerationStructure ctor
 * (source/bottom_level_acceleration_structure.cpp:34-78, top_level_...cpp:19-113): blocking. */
int  vkhrt_scene_build(VkhrtScene* scene);
/* new vertex positions, same topology: regenerate primitives, keep hierarchy, refit boxes */
int  vkhrt_scene_refit(VkhrtScene* scene, const float* positions_xyz);
int  vkhrt_scene_get_bvh(VkhrtScene* scene, VkhrtBvhView* view);
/* host copy-out of the generated primitive buffers (ModelCreation::curveBuffer /
 * lssPositionBuffer+lssRadiusBuffer / vertexBuffer; include/resources/model/model.hpp:113-126).
 * floats per primitive: PHANTOM 12 (4 control points), LSS 8 (p0,r0,p1,r1), DOTS 9 (3 vertices) */
int  vkhrt_scene_get_primitives(VkhrtScene* scene, float* out, size_t out_floats);
uint32_t vkhrt_scene_primitive_count(const VkhrtScene* scene);
void vkhrt_scene_destroy(VkhrtScene* scene);

/* ---- frame = UpdateCameraResource + traceRaysKHR(W,H,1) + image ------------------------ */
/* replaces source/renderer.cpp:156-166,189-195 and shaders/ray_gen.rgen:16-48.
 * hits_out: n_local_pixels records (nullable); rgba8_out: n_local_pixels*4 bytes (nullable). */
int  vkhrt_render(VkhrtScene* scene, const VkhrtFrameDesc* frame, VkhrtHit* hits_out, uint8_t* rgba8_out);
/* Frames in flight — Renderer::Render keeps MAX_FRAMES_IN_FLIGHT frames submitted and waits on the fence of the oldest one
 * (source/renderer.cpp:85-97,119).  vkhrt_render_submit enqueues a frame and returns; its outputs (HOST memory, page-locked for the
 * copy to be asynchronous: vkhrt_host_alloc) are complete when the matching vkhrt_render_wait returns.  Up to VKHRT_FRAMES_IN_FLIGHT
 * frames may be outstanding per scene (a further submit first waits for the oldest); frames complete in submission order, one wait per
 * submit.  While frame k's records cross PCIe on the copy engine, frame k+1 is already traversing.  Results are vkhrt_render's. */
#define VKHRT_FRAMES_IN_FLIGHT 2
int  vkhrt_render_submit(VkhrtScene* scene, const VkhrtFrameDesc* frame, VkhrtHit* hits_out, uint8_t* rgba8_out);
int  vkhrt_render_wait(VkhrtScene* scene);      /* the oldest outstanding frame; VKHRT_ERR_INVALID_ARGUMENT when none is outstanding */
/* same, and also returns the traversal counters (slower debug kernel variant) */
int  vkhrt_render_stats(VkhrtScene* scene, const VkhrtFrameDesc* frame, VkhrtHit* hits_out,
                        uint8_t* rgba8_out, VkhrtTraceStats* stats);
/* number of pixels (= hit records) a frame/shard produces, incl. padding of partial tiles */
uint64_t vkhrt_frame_local_pixels(const VkhrtFrameDesc* frame);
/* compact (tile,pixel-in-tile) shards of all ranks, concatenated rank-major -> row-major image.
 * device pointers; elem_bytes = 32 (hits) or 4 (rgba8). */
int  vkhrt_untile(const VkhrtFrameDesc* frame, uint32_t world, const void* gathered, void* row_major,
                  uint32_t elem_bytes, void* stream);
/* the same re-ordering on HOST buffers (plain loops, no device needed) */
int  vkhrt_untile_host(const VkhrtFrameDesc* frame, uint32_t world, const void* gathered, void* row_major, uint32_t elem_bytes);
/* One frame on several GPUs from ONE process (the shape of the reference's single executable, one queue: source/renderer.cpp:83-131;
 * SURVEY.md §8(b)).  scenes[r] is the same geometry built on device r (deterministic build => identical BVHs; several scenes may
 * also share a device).  The frame is cut into tile_size^2 tiles dealt round-robin (scene r traces tiles r, r + n, ...; no exchange
 * step) and assembled without any CPU re-ordering: hit records go straight from every GPU's traversal kernel into
 * the caller's buffer when it is page-locked (vkhrt_host_alloc), everything else (pixels; records into pageable memory) is stored
 * by every GPU at its row-major position of a frame buffer on scenes[0]'s device over NVLink (peer access) and copied out once.
 * Shards are launched by persistent worker threads.  `frame` must describe the whole frame (tile_stride <= 1) with
 * output_memory = VKHRT_MEM_HOST; results are identical to vkhrt_render's.  Without peer access between the devices the shards
 * are staged and re-ordered on the host.  The one-process-per-GPU path used by bench.py is vkhrt_b200/multi.py. */
int  vkhrt_render_multi(VkhrtScene* const* scenes, uint32_t n_scenes, const VkhrtFrameDesc* frame, VkhrtHit* hits_out, uint8_t* rgba8_out);
int  vkhrt_last_timing(const VkhrtScene* scene, VkhrtTiming* timing);

/* ---- page-locked host output buffers --------------------------------------------------------------------------------
 * Host memory every GPU of the box can store into directly (cudaHostAlloc, portable + mapped).  A hit-record buffer from here
 * takes vkhrt_render's zero-copy path: the traversal kernel writes the records straight into it over PCIe instead of a
 * device->host copy after the kernel, and vkhrt_render_multi lets every GPU deliver its own shard.  Any other
 * page-locked memory (cudaHostAlloc / cudaHostRegister by the caller, torch pin_memory) is recognised just the same; plain
 * malloc memory always works too, through a staging copy.  The reference's counterpart is its host-visible staging buffer
 * (source/vk_common.cpp CreateBuffer with HOST_VISIBLE | HOST_COHERENT memory). */
int  vkhrt_host_alloc(size_t bytes, void** out_ptr);
int  vkhrt_host_free(void* ptr);

/* ---- device buffers shareable between the per-GPU processes of one box (CUDA IPC over NVLink/PCIe) ---- */
/* create: plain device allocation on `device` + a 64-byte handle another process can open.
 * open:   map the exporter's allocation into this process (peer access is enabled on first use);
 *         the returned pointer is valid as hits_out / rgba8_out of vkhrt_render with VKHRT_MEM_DEVICE.
 * The reference is single-GPU (one graphics queue, source/vulkan_context.cpp:287-288): nothing replaces these. */
#define VKHRT_IPC_HANDLE_BYTES 64
int  vkhrt_shared_buffer_create(int device, size_t bytes, void** dev_ptr_out, uint8_t handle_out[VKHRT_IPC_HANDLE_BYTES]);
int  vkhrt_shared_buffer_open(int device, const uint8_t handle[VKHRT_IPC_HANDLE_BYTES], void** dev_ptr_out);
int  vkhrt_shared_buffer_close(int device, void* opened_ptr);
int  vkhrt_shared_buffer_destroy(int device, void* created_ptr);

/* ---- wavefront pieces, individually callable (device pointers) -------------------------- */
/* ray buffer entry: 32 bytes {ox,oy,oz,tmin, dx,dy,dz,tmax} */
int  vkhrt_generate_rays(const VkhrtFrameDesc* frame, uint32_t sample, float* rays_out_device, int device);
int  vkhrt_trace_rays(VkhrtScene* scene, const float* rays_device, uint64_t n_rays, VkhrtHit* hits_out_device, void* stream);
/* same with gl_RayFlagsTerminateOnFirstHitEXT semantics (shadow / occlusion rays): the record is the FIRST accepted hit
 * in traversal order (deterministic: the order is the oracle's), not the closest one; flags bit0 = occluded */
int  vkhrt_trace_rays_any_hit(VkhrtScene* scene, const float* rays_device, uint64_t n_rays, VkhrtHit* hits_out_device, void* stream);

/* ---- environment map: Renderer ctor, source/renderer.cpp:45-56 (RGBA32F image, sampled by shaders/miss.rmiss) ---- */
/* rgba32f: width*height texels, row 0 first, HOST memory (copied to the scene's GPU before return).
 * NULL / 0x0 removes the map.  Frames select it with miss_mode = VKHRT_MISS_ENVIRONMENT. */
int  vkhrt_scene_set_environment(VkhrtScene* scene, const float* rgba32f, uint32_t width, uint32_t height);

/* ---- material: shaders/bindless.glsl:6-32 (Material), the fields the closest-hit shaders read ------------------------------
 * ProcessMaterial (source/resources/model/model_loader.cpp:58-136) fills albedoFactor from the glTF base colour and
 * albedoMap from the diffuse texture; triangle_closest_hit.rchit:77-81 forms albedo = albedoFactor * texture(albedoMap, texCoord).
 * Hair primitives carry no texture coordinates (the DOTS vertices are written with zero UVs, geometry_processor.cpp:246-266;
 * curves and LSS have none at all), so the lookup is the sampler's value at (0, 0): linear filter, repeat addressing
 * (include/resources/gpu_resources.hpp:46-50) = the mean of the map's four corner texels, evaluated once per scene.
 * Frames select the term with shade_mode = VKHRT_SHADE_MATERIAL: colour = Shade(normal) * albedo.rgb. */
typedef struct VkhrtMaterial {
    float        albedo_factor[4];       /* Material::albedoFactor                                         */
    const float* albedo_map_rgba32f;     /* nullable (useAlbedoMap = false); width * height texels, row 0 first, HOST memory */
    uint32_t     albedo_map_width, albedo_map_height;
} VkhrtMaterial;
/* NULL => the default material (albedo 1,1,1,1, no map); on a multi-mesh scene: every mesh's material */
int  vkhrt_scene_set_material(VkhrtScene* scene, const VkhrtMaterial* material);

/* ---- multi-mesh scenes --------------------------------------------------------------------------------------------------
 * The reference loads several models, builds one BLAS per mesh / hair and one TLAS instance for each
 * (source/renderer.cpp:33-41, 694-727; source/top_level_acceleration_structure.cpp:19-113: instanceCustomIndex = BLAS
 * ordinal, identity-only transforms), and the closest-hit shaders find the mesh's material through
 * geometryNodes[blasInstances[gl_InstanceCustomIndexEXT].firstGeometryIndex] (shaders/hair_closest_hit.rchit:17-18).
 * Here the meshes' line lists are CONCATENATED into the scene's vertex / index arrays the way GenerateLines addresses them
 * (firstVertex / firstIndex, source/resources/model/geometry_processor.cpp:45-67; node transforms applied by the loader) and ONE
 * LBVH is built over all segments: instances are unique, so a flat hierarchy finds the same closest hit as TLAS -> BLAS with one
 * traversal instead of two.  Mesh m owns segments [first_segment[m], first_segment[m + 1]) (the last one up to n_segments);
 * first_segment[0] must be 0 and the list ascending (an empty mesh repeats its successor's value).  A hit record's `segment` is the
 * scene-wide index; vkhrt_scene_mesh_of_segments maps it to the mesh (= gl_InstanceCustomIndexEXT), VKHRT_MISS_SEGMENT for a miss.
 * shade_mode VKHRT_SHADE_MATERIAL uses the material of the mesh that owns the hit segment.  Strands must not span meshes
 * (GenerateCurves looks at the neighbouring segment only when it shares a vertex, so concatenation keeps every curve as it was).
 * Not combinable with vkhrt_scene_apply_lod (the passes renumber segments).  n_meshes = 0 returns to a single mesh. */
int      vkhrt_scene_set_meshes(VkhrtScene* scene, const uint32_t* first_segment, uint32_t n_meshes);
int      vkhrt_scene_set_mesh_material(VkhrtScene* scene, uint32_t mesh, const VkhrtMaterial* material);
uint32_t vkhrt_scene_mesh_count(const VkhrtScene* scene);
int      vkhrt_scene_mesh_of_segments(const VkhrtScene* scene, const uint32_t* segments, uint32_t* mesh_out, size_t n);

/* ---- strand level of detail on the device, before the build ------------------------------
 * The reference defines but never calls MergeLines / SplitLines / MergeCurvesFast
 * (source/resources/model/geometry_processor.cpp:69-104, 106-121, 158-197).  Applied in this order to the scene's
 * line list: `line_split_passes` x SplitLines, `line_merge_passes` x MergeLines, then (PHANTOM only, after
 * GenerateCurves) `curve_merge_passes` x MergeCurvesFast.  Must be called before vkhrt_scene_build; afterwards
 * segment ids count the processed lines / curves, and vkhrt_scene_refit is refused (the input vertices are gone).
 * Like the reference, MergeLines / MergeCurvesFast drop the last element of an odd-sized list. */
int  vkhrt_scene_apply_lod(VkhrtScene* scene, uint32_t line_split_passes, uint32_t line_merge_passes, uint32_t curve_merge_passes);
/* number of line segments after LOD (= BVH leaves) and a host copy of them: n_segments * 6 floats {start.xyz, end.xyz} */
uint32_t vkhrt_scene_segment_count(const VkhrtScene* scene);
int  vkhrt_scene_get_lines(VkhrtScene* scene, float* out, size_t out_floats);

/* ---- asset ingest and image output (host; no GPU needed): SURVEY.md §8(f) rows 3, 4 --------------------------------
 * Replaces ModelLoader::LoadFromFile + ProcessMesh for line primitives (source/resources/model/model_loader.cpp:139-206,
 * 274-291; Assimp is not vendored) and LoadFloatImageFromFile (source/resources/file_io.cpp:22-37, stbi_loadf). */
typedef struct VkhrtLineAsset {
    float*    positions_xyz;      /* n_vertices * 3, malloc'd by the loader                          */
    uint32_t  n_vertices;
    uint32_t* line_indices;       /* n_segments * 2                                                  */
    uint32_t  n_segments;
    float*    radius_per_vertex;  /* NULL unless the file carries a thickness array (.hair): thickness / 2 */
    uint32_t  n_strands;
    float     base_color[4];      /* Material::albedoFactor as ProcessMaterial reads it (AI_MATKEY_BASE_COLOR, model_loader.cpp:96-99):
                                     glTF pbrMetallicRoughness.baseColorFactor of the first line primitive's material; 1,1,1,1 if none */
} VkhrtLineAsset;
/* by extension: .obj (`v` + `l` polyline records), .hair (Cem Yuksel HAIR format), .gltf / .glb (glTF 2.0 line primitives: modes LINES,
 * LINE_LOOP, LINE_STRIP; external, base64 or GLB buffers; node transforms applied to the positions) — the format of the reference's own
 * scene (source/renderer.cpp:33-37).  Saving: .obj, .hair and .glb (one LINES primitive, loadable by Assimp and therefore by the reference). */
int  vkhrt_asset_load_lines(const char* path, VkhrtLineAsset* out);
int  vkhrt_asset_save_lines(const char* path, const VkhrtLineAsset* in);
void vkhrt_asset_free(VkhrtLineAsset* asset);
/* Radiance .hdr (RGBE, flat or RLE scanlines) -> RGBA32F exactly as stbi_loadf(path, &w, &h, &n, 4); free with vkhrt_image_free */
int  vkhrt_image_load_hdr(const char* path, float** rgba_out, uint32_t* width_out, uint32_t* height_out);
int  vkhrt_image_save_hdr(const char* path, const float* rgba, uint32_t width, uint32_t height);
void vkhrt_image_free(float* rgba);
/* OpenEXR, scanline, uncompressed, four FLOAT channels (A, B, G, R): float images (environment maps, linear frames) for tools that
 * read the reference's HDR inputs */
int  vkhrt_image_save_exr(const char* path, const float* rgba, uint32_t width, uint32_t height);
/* the frame the reference presents to its swap chain (source/renderer.cpp:222-231) as an 8-bit RGBA PNG */
int  vkhrt_image_save_png(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height);
/* procedural equirectangular sky, RGBA32F (the reference's .hdr asset is not in its repository) */
void vkhrt_environment_generate(uint32_t width, uint32_t height, float* rgba_out);

/* ---- host helpers: FlyCamera (source/fly_camera.cpp:25-35) + Renderer::UpdateCameraResource -- */
/* fov in degrees (vertical); yaw/pitch in degrees as FlyCamera (defaults -90, 0) */
void vkhrt_camera_matrices(const float position[3], float yaw_deg, float pitch_deg, float fov_deg,
                           float aspect, float near_plane, float far_plane,
                           float view_inverse_out[16], float proj_inverse_out[16]);

/* ---- synthetic procedural grooms (hair assets are not available offline) ---------------- */
typedef enum VkhrtGroomStyle { VKHRT_GROOM_STRAIGHT = 0, VKHRT_GROOM_CURLY = 1 } VkhrtGroomStyle;
/* writes n_strands*(segs+1) positions and n_strands*segs index pairs; deterministic in seed */
int  vkhrt_groom_generate(uint32_t n_strands, uint32_t segments_per_strand, int32_t style, uint64_t seed,
                          float* positions_xyz_out, uint32_t* line_indices_out);

#ifdef __cplusplus
}
#endif
#endif /* VKHRT_B200_H */
