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
// field instead of a source edit (model_loader.cpp:334); asset import is a minimal OBJ polyline reader
// (Assimp is not available offline) that yields the same "positions + index pairs" line mesh
// ProcessMesh produces (model_loader.cpp:139-206), plus a `synthetic:` URI for the seeded grooms.
// Error behaviour follows the reference: asset failures log and return nullptr (model_loader.cpp:280-284);
// device failures are fatal there (abort, vk_common.cpp:5-15) and throw VkhrtError here.
// This layer holds no compute: every number comes from libvkhrt_b200.so (CUDA); there is no CPU path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/vkhrt_b200.h"

namespace vkhrt_host {

struct VkhrtError : std::runtime_error {
    int status;
    VkhrtError(int s, const std::string& where)
        : std::runtime_error(where + ": " + vkhrt_error_string(s) + " (" + std::to_string(s) + ") " + vkhrt_last_error()), status(s) {}
};
inline void Check(int status, const char* where) { if (status != VKHRT_OK) throw VkhrtError(status, where); }

struct vec3 { float x = 0, y = 0, z = 0; };

// The line mesh the reference's importer hands to the geometry processor, reduced to what the hair
// techniques read (positions and 2-index line faces), plus the per-vertex radius extension.
struct ModelCreation {
    std::vector<vec3> vertexBuffer {};          // Mesh::Vertex::position
    std::vector<uint32_t> indexBuffer {};       // pairs (start, end); strand joints repeat the vertex index
    std::vector<float> radiusBuffer {};         // optional, one per vertex (LSS); empty => `radius`
    float radius = VKHRT_DEFAULT_RADIUS;
    VkhrtTechnique technique = VKHRT_TECHNIQUE_LSS;   // the reference's own default when the LSS extension exists
    std::string sceneName {};
};

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
        int rc = vkhrt_scene_build(_scene);
        if (rc != VKHRT_OK) { vkhrt_scene_destroy(_scene); _scene = nullptr; throw VkhrtError(rc, "vkhrt_scene_build"); }
    }
    ~Model() { if (_scene) vkhrt_scene_destroy(_scene); }
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;

    void Refit(const std::vector<vec3>& positions) { Check(vkhrt_scene_refit(_scene, &positions[0].x), "vkhrt_scene_refit"); }
    [[nodiscard]] uint32_t PrimitiveCount() const { return vkhrt_scene_primitive_count(_scene); }
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
    // path: an OBJ file with `v x y z` and `l i j k ...` polyline records (1-based, negative = relative), or
    // "synthetic:<straight|curly>:<strands>:<segments>[:<seed>]".  Returns nullptr on failure like the reference.
    [[nodiscard]] std::shared_ptr<Model> LoadFromFile(std::string_view path, VkhrtTechnique technique = VKHRT_TECHNIQUE_LSS)
    {
        ModelCreation creation;
        if (!LoadModel(path, creation)) { std::fprintf(stderr, "[MODEL LOADING] Failed to load %.*s\n", (int)path.size(), path.data()); return nullptr; }
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
        std::ifstream in(p);
        if (!in) return false;
        for (std::string line; std::getline(in, line);) {
            std::stringstream ls(line);
            std::string tag;
            ls >> tag;
            if (tag == "v") { vec3 v; ls >> v.x >> v.y >> v.z; if (!ls) return false; out.vertexBuffer.push_back(v); }
            else if (tag == "l") {
                long prev = 0; bool have = false;
                for (std::string tok; ls >> tok;) {
                    long i = std::atol(tok.c_str());          // "i" or "i/t"
                    if (i == 0) return false;
                    long idx = i > 0 ? i - 1 : (long)out.vertexBuffer.size() + i;
                    if (idx < 0 || idx >= (long)out.vertexBuffer.size()) return false;
                    if (have) { out.indexBuffer.push_back((uint32_t)prev); out.indexBuffer.push_back((uint32_t)idx); }
                    prev = idx; have = true;
                }
            }
        }
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
};

class Renderer {
public:
    Renderer(const RendererInitInfo& initInfo, const std::shared_ptr<FlyCamera>& flyCamera) : _info(initInfo), _flyCamera(flyCamera) {}
    void AddModel(const std::shared_ptr<Model>& model) { _models.push_back(model); }
    [[nodiscard]] const std::vector<std::shared_ptr<Model>>& GetModels() const { return _models; }

    // One frame: UpdateCameraResource + traceRaysKHR(width, height, 1) + read-back, blocking.
    // (The reference TLAS holds several BLAS; this path renders one groom per Renderer, the first model.)
    void Render()
    {
        if (_models.empty()) throw std::runtime_error("Renderer::Render: no model");
        VkhrtFrameDesc f {};
        _flyCamera->CameraUniformData(f.view_inverse, f.proj_inverse);
        f.width = _info.width; f.height = _info.height; f.spp = _info.spp; f.shade_mode = _info.shadeMode;
        std::memcpy(f.miss_rgb, _info.missColor, sizeof(f.miss_rgb));
        f.output_memory = VKHRT_MEM_HOST;
        const size_t n = (size_t)_info.width * _info.height;
        if (_info.wantHits) _hits.resize(n);
        if (_info.wantImage) _image.resize(n * 4);
        Check(vkhrt_render(_models[0]->Handle(), &f, _info.wantHits ? _hits.data() : nullptr, _info.wantImage ? _image.data() : nullptr), "vkhrt_render");
    }
    [[nodiscard]] const std::vector<VkhrtHit>& GetHits() const { return _hits; }
    [[nodiscard]] const std::vector<uint8_t>& GetImage() const { return _image; }     // RGBA8, row-major, row 0 = top
    [[nodiscard]] const RendererInitInfo& GetInitInfo() const { return _info; }

    bool WritePPM(const std::string& path) const
    {
        std::ofstream out(path, std::ios::binary);
        if (!out) return false;
        out << "P6\n" << _info.width << " " << _info.height << "\n255\n";
        for (size_t i = 0; i < (size_t)_info.width * _info.height; ++i) out.write((const char*)&_image[4 * i], 3);
        return (bool)out;
    }

private:
    RendererInitInfo _info;
    std::shared_ptr<FlyCamera> _flyCamera;
    std::vector<std::shared_ptr<Model>> _models;
    std::vector<VkhrtHit> _hits;
    std::vector<uint8_t> _image;
};

}  // namespace vkhrt_host
