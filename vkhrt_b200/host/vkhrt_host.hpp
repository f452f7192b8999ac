// vkhrt_host.hpp — C++ host layer over the C ABI (include/vkhrt_b200.h), header-only.
//
// Mirrors the reference's host surface for the ray-tracing path so a vkhrt call site reads the same
// (reference paths relative to the reference root):
//   ModelCreation, ProcessHair{Curves,LSS,DOTS}  include/resources/model/model.hpp:113-126, geometry_processor.hpp:4-7
//   ModelLoader::LoadFromFile                     include/resources/model/model_loader.hpp:20
//   Model (device upload) + BLAS/TLAS build       source/resources/model/model.cpp:52-223, bottom_level_acceleration_structure.cpp:34-78
//   FlyCameraCreation / FlyCamera                 include/fly_camera.hpp:7-52, source/fly_camera.cpp:25-35
//   Renderer::Render / GetModels                  include/renderer.hpp:27-31, source/renderer.cpp:83-166,189-195
// Differences, by design: no window, swap chain, ImGui or Vulkan objects; the technique is a run-time
// field instead of a source edit (model_loader.cpp:334); asset import goes through the ABI's line-asset readers
// (vkhrt_asset_load_lines: OBJ polylines and Cem Yuksel HAIR files; Assimp is not available offline), which yield the
// same "positions + index pairs" line mesh ProcessMesh produces (model_loader.cpp:139-206), plus a `synthetic:` URI
// for the seeded grooms.  The environment map of Renderer's constructor (renderer.cpp:45-56) is a Radiance .hdr file
// or the procedural sky; the unused LOD helpers of geometry_processor.cpp:69-197 are ModelCreation fields.
// Error behaviour follows the reference: asset failures log and return nullptr (model_loader.cpp:280-284);
// device failures are fatal there (abort, vk_common.cpp:5-15) and throw VkhrtError here.
// This layer holds no compute: every number comes from libvkhrt_b200.so (CUDA); there is no CPU path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <array>
#include <vector>

#include "../../include/vkhrt_b200.h"

namespace vkhrt_host {

struct VkhrtError : std::runtime_error {
    int status;
    VkhrtError(int s, const std::string& where)
        : std::runtime_error(where + ": " + vkhrt_error_string(s) + " (" + std::to_string(s) + ") " + vkhrt_last_error()), status(s) {}
};
inline void Check(int status, const char* where) { if (status != VKHRT_OK) throw VkhrtError(status, where); }

// Output buffer in page-locked host memory (vkhrt_host_alloc): what the reference keeps in HOST_VISIBLE staging memory.
// Falls back to plain heap memory when the allocation is refused (then vkhrt_render stages the copy itself).
template <typename T>
class HostBuffer {
public:
    HostBuffer() = default;
    HostBuffer(const HostBuffer&) = delete;
    HostBuffer& operator=(const HostBuffer&) = delete;
    ~HostBuffer() { release(); }
    void resize(size_t n)
    {
        if (n == _n) return;
        release();
        if (n == 0) return;
        void* p = nullptr;
        if (vkhrt_host_alloc(n * sizeof(T), &p) == VKHRT_OK) { _p = static_cast<T*>(p); _pinned = true; }
        else { _p = static_cast<T*>(std::malloc(n * sizeof(T))); _pinned = false; if (!_p) throw std::bad_alloc(); }
        _n = n;
    }
    [[nodiscard]] T* data() { return _p; }
    [[nodiscard]] const T* data() const { return _p; }
    [[nodiscard]] size_t size() const { return _n; }
    [[nodiscard]] bool empty() const { return _n == 0; }
    [[nodiscard]] bool pinned() const { return _pinned; }
    [[nodiscard]] const T& operator[](size_t i) const { return _p[i]; }
    [[nodiscard]] const T* begin() const { return _p; }
    [[nodiscard]] const T* end() const { return _p + _n; }
private:
    void release()
    {
        if (_p) { if (_pinned) vkhrt_host_free(_p); else std::free(_p); }
        _p = nullptr; _n = 0; _pinned = false;
    }
    T* _p = nullptr;
    size_t _n = 0;
    bool _pinned = false;
};

struct vec3 { float x = 0, y = 0, z = 0; };

// The line mesh the reference's importer hands to the geometry processor, reduced to what the hair
// techniques read (positions and 2-index line faces), plus the per-vertex radius extension.
struct ModelCreation {
    std::vector<vec3> vertexBuffer {};          // Mesh::Vertex::position
    std::vector<uint32_t> indexBuffer {};       // pairs (start, end); strand joints repeat the vertex index
    std::vector<float> radiusBuffer {};         // optional, one per vertex (all techniques); empty => `radius`
    float albedoFactor[4] { 1.0f, 1.0f, 1.0f, 1.0f };   // MaterialCreation::albedoFactor (ProcessMaterial, model_loader.cpp:96-99)
    float radius = VKHRT_DEFAULT_RADIUS;
    VkhrtTechnique technique = VKHRT_TECHNIQUE_LSS;   // the reference's own default when the LSS extension exists
    std::string sceneName {};
    // strand level of detail, run on the device before the build (vkhrt_scene_apply_lod):
    // SplitLines / MergeLines / MergeCurvesFast of geometry_processor.cpp:69-197
    uint32_t lineSplitPasses = 0, lineMergePasses = 0, curveMergePasses = 0;
    // several meshes in one model / scene (MergeModels): mesh m = segments [meshFirstSegment[m], meshFirstSegment[m + 1]) with its own
    // albedo factor; empty = one mesh.  The reference keeps one BLAS per mesh under a TLAS (renderer.cpp:694-727).
    std::vector<uint32_t> meshFirstSegment {};
    std::vector<std::array<float, 4>> meshAlbedoFactor {};
};

// The scene of several models as ONE line list, addressed the way GenerateLines addresses a mesh inside the model's buffers
// (firstVertex / firstIndex, geometry_processor.cpp:45-67).  Radii: per vertex as soon as one part has them (the others get their constant).
inline ModelCreation MergeModels(const std::vector<ModelCreation>& parts)
{
    ModelCreation out;
    if (parts.empty()) return out;
    out.radius = parts[0].radius; out.technique = parts[0].technique;
    bool anyRadii = false;
    for (const ModelCreation& p : parts) anyRadii = anyRadii || !p.radiusBuffer.empty();
    for (const ModelCreation& p : parts) {
        const uint32_t firstVertex = (uint32_t)out.vertexBuffer.size();
        out.meshFirstSegment.push_back((uint32_t)(out.indexBuffer.size() / 2));
        out.meshAlbedoFactor.push_back({ p.albedoFactor[0], p.albedoFactor[1], p.albedoFactor[2], p.albedoFactor[3] });
        out.vertexBuffer.insert(out.vertexBuffer.end(), p.vertexBuffer.begin(), p.vertexBuffer.end());
        for (uint32_t i : p.indexBuffer) out.indexBuffer.push_back(firstVertex + i);
        if (anyRadii) {
            if (p.radiusBuffer.empty()) out.radiusBuffer.insert(out.radiusBuffer.end(), p.vertexBuffer.size(), p.radius);
            else out.radiusBuffer.insert(out.radiusBuffer.end(), p.radiusBuffer.begin(), p.radiusBuffer.end());
        }
        out.sceneName += (out.sceneName.empty() ? "" : " + ") + p.sceneName;
    }
    return out;
}

// geometry_processor.hpp:4-7 — in the reference these run the generators on the host; here they only
// select which generator the device build runs (build.cu), keeping the call sites identical.
inline ModelCreation ProcessHairCurves(const ModelCreation& m) { ModelCreation r = m; r.technique = VKHRT_TECHNIQUE_PHANTOM; return r; }
inline ModelCreation ProcessHairLSS(const ModelCreation& m) { ModelCreation r = m; r.technique = VKHRT_TECHNIQUE_LSS; return r; }
inline ModelCreation ProcessHairDOTS(const ModelCreation& m) { ModelCreation r = m; r.technique = VKHRT_TECHNIQUE_DOTS; return r; }

// Device-resident model: primitives + LBVH in HBM.  Construction = Model::Model upload + BLAS/TLAS build (blocking).
class Model {
public:
    Model(const ModelCreation& creation, int device = 0) : _technique(creation.technique)
    {
        VkhrtSceneDesc d {};
        d.positions_xyz = creation.vertexBuffer.empty() ? nullptr : &creation.vertexBuffer[0].x;
        d.n_vertices = (uint32_t)creation.vertexBuffer.size();
        d.line_indices = creation.indexBuffer.empty() ? nullptr : creation.indexBuffer.data();
        d.n_segments = (uint32_t)(creation.indexBuffer.size() / 2);
        d.radius_per_vertex = creation.radiusBuffer.empty() ? nullptr : creation.radiusBuffer.data();
        d.radius = creation.radius;
        d.technique = creation.technique;
        d.device = device;
        Check(vkhrt_scene_create(&d, &_scene), "vkhrt_scene_create");
        VkhrtMaterial mat {};
        std::memcpy(mat.albedo_factor, creation.albedoFactor, sizeof(mat.albedo_factor));
        vkhrt_scene_set_material(_scene, &mat);
        if (creation.meshFirstSegment.size() > 1) {
            Check(vkhrt_scene_set_meshes(_scene, creation.meshFirstSegment.data(), (uint32_t)creation.meshFirstSegment.size()), "vkhrt_scene_set_meshes");
            for (size_t m = 0; m < creation.meshAlbedoFactor.size(); ++m) {
                std::memcpy(mat.albedo_factor, creation.meshAlbedoFactor[m].data(), sizeof(mat.albedo_factor));
                Check(vkhrt_scene_set_mesh_material(_scene, (uint32_t)m, &mat), "vkhrt_scene_set_mesh_material");
            }
        }
        int rc = (creation.lineSplitPasses | creation.lineMergePasses | creation.curveMergePasses) == 0u ? VKHRT_OK : vkhrt_scene_apply_lod(_scene, creation.lineSplitPasses, creation.lineMergePasses, creation.curveMergePasses);
        if (rc == VKHRT_OK) rc = vkhrt_scene_build(_scene);
        if (rc != VKHRT_OK) { vkhrt_scene_destroy(_scene); _scene = nullptr; throw VkhrtError(rc, "vkhrt_scene_build"); }
    }
    ~Model() { if (_scene) vkhrt_scene_destroy(_scene); }
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;

    void Refit(const std::vector<vec3>& positions) { Check(vkhrt_scene_refit(_scene, &positions[0].x), "vkhrt_scene_refit"); }
    [[nodiscard]] uint32_t PrimitiveCount() const { return vkhrt_scene_primitive_count(_scene); }
    [[nodiscard]] uint32_t MeshCount() const { return vkhrt_scene_mesh_count(_scene); }
    // gl_InstanceCustomIndexEXT of a hit record: the mesh that owns its segment
    [[nodiscard]] uint32_t MeshOfSegment(uint32_t segment) const { uint32_t m = 0; Check(vkhrt_scene_mesh_of_segments(_scene, &segment, &m, 1), "vkhrt_scene_mesh_of_segments"); return m; }
    [[nodiscard]] VkhrtTechnique Technique() const { return _technique; }
    [[nodiscard]] VkhrtScene* Handle() const { return _scene; }
    [[nodiscard]] VkhrtTiming Timing() const { VkhrtTiming t {}; Check(vkhrt_last_timing(_scene, &t), "vkhrt_last_timing"); return t; }

private:
    VkhrtScene* _scene = nullptr;
    VkhrtTechnique _technique;
};

class ModelLoader {
public:
    explicit ModelLoader(int device = 0) : _device(device) {}
    // path: a line asset (.gltf / .glb line primitives — what the reference loads, renderer.cpp:33-37 —, .obj with `v` / `l i j k ...`
    // polyline records, or a Cem Yuksel .hair file), or
    // "synthetic:<straight|curly>:<strands>:<segments>[:<seed>]".  Returns nullptr on failure like the reference.
    [[nodiscard]] std::shared_ptr<Model> LoadFromFile(std::string_view path, VkhrtTechnique technique = VKHRT_TECHNIQUE_LSS,
                                                      uint32_t lineSplitPasses = 0, uint32_t lineMergePasses = 0, uint32_t curveMergePasses = 0)
    {
        ModelCreation creation;
        if (!LoadModel(path, creation)) { std::fprintf(stderr, "[MODEL LOADING] Failed to load %.*s\n", (int)path.size(), path.data()); return nullptr; }
        creation.technique = technique;
        creation.lineSplitPasses = lineSplitPasses; creation.lineMergePasses = lineMergePasses; creation.curveMergePasses = curveMergePasses;
        return ProcessModel(creation);
    }
    // The reference's scene is a LIST of models (renderer.cpp:33-41): all of them in one device scene, one mesh each
    [[nodiscard]] std::shared_ptr<Model> LoadFromFiles(const std::vector<std::string>& paths, VkhrtTechnique technique = VKHRT_TECHNIQUE_LSS)
    {
        std::vector<ModelCreation> parts(paths.size());
        for (size_t i = 0; i < paths.size(); ++i)
            if (!LoadModel(paths[i], parts[i])) { std::fprintf(stderr, "[MODEL LOADING] Failed to load %s\n", paths[i].c_str()); return nullptr; }
        ModelCreation creation = MergeModels(parts);
        creation.technique = technique;
        return ProcessModel(creation);
    }
    [[nodiscard]] static bool LoadModel(std::string_view path, ModelCreation& out)
    {
        out = ModelCreation {};
        out.sceneName = std::string(path);
        const std::string p(path);
        if (p.rfind("synthetic:", 0) == 0) {
            std::vector<std::string> f;
            std::stringstream ss(p.substr(10));
            for (std::string tok; std::getline(ss, tok, ':');) f.push_back(tok);
            if (f.size() < 3) return false;
            const int style = f[0] == "straight" ? VKHRT_GROOM_STRAIGHT : (f[0] == "curly" ? VKHRT_GROOM_CURLY : -1);
            const long strands = std::atol(f[1].c_str()), segs = std::atol(f[2].c_str());
            const uint64_t seed = f.size() > 3 ? std::strtoull(f[3].c_str(), nullptr, 0) : 0x5EED0001ull;
            if (style < 0 || strands <= 0 || segs <= 0) return false;
            out.vertexBuffer.resize((size_t)strands * (segs + 1));
            out.indexBuffer.resize((size_t)strands * segs * 2);
            return vkhrt_groom_generate((uint32_t)strands, (uint32_t)segs, style, seed, &out.vertexBuffer[0].x, out.indexBuffer.data()) == VKHRT_OK;
        }
        VkhrtLineAsset a {};
        if (vkhrt_asset_load_lines(p.c_str(), &a) != VKHRT_OK) return false;
        out.vertexBuffer.resize(a.n_vertices);
        if (a.n_vertices) std::memcpy(&out.vertexBuffer[0].x, a.positions_xyz, (size_t)a.n_vertices * 12);
        out.indexBuffer.assign(a.line_indices, a.line_indices + (size_t)a.n_segments * 2);
        if (a.radius_per_vertex) out.radiusBuffer.assign(a.radius_per_vertex, a.radius_per_vertex + a.n_vertices);
        std::memcpy(out.albedoFactor, a.base_color, sizeof(out.albedoFactor));
        vkhrt_asset_free(&a);
        return !out.indexBuffer.empty();
    }

private:
    [[nodiscard]] std::shared_ptr<Model> ProcessModel(const ModelCreation& creation) const
    {
        // model_loader.cpp:314-336 picks ProcessHairLSS or ProcessHairDOTS; here the technique is explicit
        switch (creation.technique) {
        case VKHRT_TECHNIQUE_PHANTOM: return std::make_shared<Model>(ProcessHairCurves(creation), _device);
        case VKHRT_TECHNIQUE_DOTS: return std::make_shared<Model>(ProcessHairDOTS(creation), _device);
        default: return std::make_shared<Model>(ProcessHairLSS(creation), _device);
        }
    }
    int _device;
};

struct FlyCameraCreation {                 // include/fly_camera.hpp:7-18, values of application.cpp:65-73
    vec3 position { 0.0f, 150.0f, 20.0f };
    float fov = 60.0f;
    float aspectRatio = 16.0f / 9.0f;
    float nearPlane = 0.1f;
    float farPlane = 1000.0f;
    float yaw = -90.0f, pitch = 0.0f;      // private members of the reference camera, exposed: there is no input device
};

class FlyCamera {
public:
    explicit FlyCamera(const FlyCameraCreation& creation) : _c(creation) {}
    void SetPose(vec3 position, float yaw, float pitch) { _c.position = position; _c.yaw = yaw; _c.pitch = pitch; }
    void SetAspectRatio(float a) { _c.aspectRatio = a; }
    // CameraUniformData {viewInverse, projInverse} as Renderer::UpdateCameraResource uploads it (renderer.cpp:189-195)
    void CameraUniformData(float viewInverse[16], float projInverse[16]) const
    {
        const float pos[3] = { _c.position.x, _c.position.y, _c.position.z };
        vkhrt_camera_matrices(pos, _c.yaw, _c.pitch, _c.fov, _c.aspectRatio, _c.nearPlane, _c.farPlane, viewInverse, projInverse);
    }

private:
    FlyCameraCreation _c;
};

struct RendererInitInfo {                  // stands in for VulkanInitInfo: the swap-chain extent becomes an image size
    uint32_t width = 1920, height = 1080;
    uint32_t spp = 1;
    VkhrtShadeMode shadeMode = VKHRT_SHADE;
    float missColor[3] = { 0.0f, 0.0f, 0.0f };
    bool wantHits = true, wantImage = true;
    // environment map sampled by the miss shader (renderer.cpp:45-56, miss.rmiss): "" = constant missColor,
    // "procedural" = the built-in sky, anything else = a Radiance .hdr file
    std::string environmentMap {};
    uint32_t aoSamples = 0;                // ambient-occlusion rays per primary hit (not in the reference)
};

class Renderer {
public:
    Renderer(const RendererInitInfo& initInfo, const std::shared_ptr<FlyCamera>& flyCamera) : _info(initInfo), _flyCamera(flyCamera)
    {
        // Initialize scene environment map (renderer.cpp:45-56); a failed load logs and leaves the constant colour
        if (_info.environmentMap == "procedural") {
            _envW = 1024; _envH = 512;
            _env.resize((size_t)_envW * _envH * 4);
            vkhrt_environment_generate(_envW, _envH, _env.data());
        } else if (!_info.environmentMap.empty()) {
            float* data = nullptr;
            if (vkhrt_image_load_hdr(_info.environmentMap.c_str(), &data, &_envW, &_envH) == VKHRT_OK) {
                _env.assign(data, data + (size_t)_envW * _envH * 4);
                vkhrt_image_free(data);
            } else std::fprintf(stderr, "[IMAGE LOADING] Failed to load data for image from path [%s]\n", _info.environmentMap.c_str());
        }
    }
    void AddModel(const std::shared_ptr<Model>& model)
    {
        if (!_env.empty()) Check(vkhrt_scene_set_environment(model->Handle(), _env.data(), _envW, _envH), "vkhrt_scene_set_environment");
        _models.push_back(model);
    }
    [[nodiscard]] const std::vector<std::shared_ptr<Model>>& GetModels() const { return _models; }
    // The same groom built on another GPU (ModelLoader(device).LoadFromFile of the same asset; the build is deterministic, so the
    // replicas are identical).  With replicas, Render() shards the frame over all of them from this one process (vkhrt_render_multi).
    void AddReplica(const std::shared_ptr<Model>& model)
    {
        if (!_env.empty()) Check(vkhrt_scene_set_environment(model->Handle(), _env.data(), _envW, _envH), "vkhrt_scene_set_environment");
        _replicas.push_back(model);
    }

    // One frame: UpdateCameraResource + traceRaysKHR(width, height, 1) + read-back, blocking.
    // (The reference TLAS holds several BLAS; here a scene of several meshes is ONE Model — ModelLoader::LoadFromFiles / MergeModels.)
    void Render()
    {
        if (_models.empty()) throw std::runtime_error("Renderer::Render: no model");
        VkhrtFrameDesc f = FrameDesc();
        const size_t n = (size_t)_info.width * _info.height;
        if (_info.wantHits) _hits.resize(n);
        if (_info.wantImage) _image.resize(n * 4);
        if (_replicas.empty()) {
            Check(vkhrt_render(_models[0]->Handle(), &f, _info.wantHits ? _hits.data() : nullptr, _info.wantImage ? _image.data() : nullptr), "vkhrt_render");
        } else {
            std::vector<VkhrtScene*> handles { _models[0]->Handle() };
            for (const auto& r : _replicas) handles.push_back(r->Handle());
            Check(vkhrt_render_multi(handles.data(), (uint32_t)handles.size(), &f, _info.wantHits ? _hits.data() : nullptr, _info.wantImage ? _image.data() : nullptr),
                  "vkhrt_render_multi");
        }
    }
    // Frames in flight (renderer.cpp:85-97,119: MAX_FRAMES_IN_FLIGHT frames submitted, the fence of the oldest waited on): Submit() enqueues
    // a frame into buffer set (frame % VKHRT_FRAMES_IN_FLIGHT) and returns; Wait() completes the oldest one and returns its buffer index.
    // While frame k's records cross PCIe on the copy engine, frame k+1 is already traversing.  One model, no replicas.
    void Submit()
    {
        if (_models.empty()) throw std::runtime_error("Renderer::Submit: no model");
        if (!_replicas.empty()) throw std::runtime_error("Renderer::Submit: frames in flight are a single-GPU path");
        VkhrtFrameDesc f = FrameDesc();
        const size_t n = (size_t)_info.width * _info.height;
        const int k = (int)(_submitted % VKHRT_FRAMES_IN_FLIGHT);
        if (_info.wantHits) _flightHits[k].resize(n);
        if (_info.wantImage) _flightImage[k].resize(n * 4);
        Check(vkhrt_render_submit(_models[0]->Handle(), &f, _info.wantHits ? _flightHits[k].data() : nullptr, _info.wantImage ? _flightImage[k].data() : nullptr),
              "vkhrt_render_submit");
        ++_submitted;
    }
    int Wait()
    {
        if (_waited == _submitted) throw std::runtime_error("Renderer::Wait: no frame outstanding");
        Check(vkhrt_render_wait(_models[0]->Handle()), "vkhrt_render_wait");
        return (int)(_waited++ % VKHRT_FRAMES_IN_FLIGHT);
    }
    [[nodiscard]] const HostBuffer<VkhrtHit>& GetFlightHits(int k) const { return _flightHits[k]; }
    [[nodiscard]] const HostBuffer<uint8_t>& GetFlightImage(int k) const { return _flightImage[k]; }
    [[nodiscard]] const HostBuffer<VkhrtHit>& GetHits() const { return _hits; }
    [[nodiscard]] const HostBuffer<uint8_t>& GetImage() const { return _image; }     // RGBA8, row-major, row 0 = top
    [[nodiscard]] const RendererInitInfo& GetInitInfo() const { return _info; }

    bool WritePPM(const std::string& path) const
    {
        std::ofstream out(path, std::ios::binary);
        if (!out) return false;
        out << "P6\n" << _info.width << " " << _info.height << "\n255\n";
        for (size_t i = 0; i < (size_t)_info.width * _info.height; ++i) out.write((const char*)&_image[4 * i], 3);
        return (bool)out;
    }

    bool WritePNG(const std::string& path) const
    {
        return !_image.empty() && vkhrt_image_save_png(path.c_str(), _image.data(), _info.width, _info.height) == VKHRT_OK;
    }

private:
    // UpdateCameraResource (renderer.cpp:189-195) + the launch size of traceRaysKHR
    [[nodiscard]] VkhrtFrameDesc FrameDesc() const
    {
        VkhrtFrameDesc f {};
        _flyCamera->CameraUniformData(f.view_inverse, f.proj_inverse);
        f.width = _info.width; f.height = _info.height; f.spp = _info.spp; f.shade_mode = _info.shadeMode;
        std::memcpy(f.miss_rgb, _info.missColor, sizeof(f.miss_rgb));
        f.output_memory = VKHRT_MEM_HOST;
        f.miss_mode = _env.empty() ? VKHRT_MISS_CONSTANT : VKHRT_MISS_ENVIRONMENT;
        f.ao_samples = _info.aoSamples;
        return f;
    }
    HostBuffer<VkhrtHit> _flightHits[VKHRT_FRAMES_IN_FLIGHT];
    HostBuffer<uint8_t> _flightImage[VKHRT_FRAMES_IN_FLIGHT];
    uint64_t _submitted = 0, _waited = 0;
    RendererInitInfo _info;
    std::vector<float> _env;
    uint32_t _envW = 0, _envH = 0;
    std::shared_ptr<FlyCamera> _flyCamera;
    std::vector<std::shared_ptr<Model>> _models;
    std::vector<std::shared_ptr<Model>> _replicas;
    HostBuffer<VkhrtHit> _hits;       // page-locked: the GPUs store hit records straight into it
    HostBuffer<uint8_t> _image;
};

}  // namespace vkhrt_host
