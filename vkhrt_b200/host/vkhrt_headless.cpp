// vkhrt_headless — the reference's `main` (source/main.cpp:3-7 -> Application -> Renderer::Render loop,
// source/application.cpp:101-129) without the window: load or synthesise a groom, build, render N frames,
// write the image and the hit buffer, print per-stage device timings.
//
//   vkhrt_headless --model synthetic:curly:100000:32 --technique phantom --size 1920x1080
//                  [--spp 1] [--debug-primid] [--frames 10] [--ppm out.ppm] [--png out.png] [--hits out.bin] [--device 0]
//                  [--env procedural|sky.hdr] [--ao N] [--lod split,merge,curve_merge]
//
// Host code only: every number comes from libvkhrt_b200.so; without a CUDA device it exits with the ABI's error.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "vkhrt_host.hpp"

using namespace vkhrt_host;

static void usage()
{
    std::puts("usage: vkhrt_headless --model <file.gltf | file.glb | file.obj | file.hair | synthetic:<straight|curly>:<strands>:<segments>[:seed]>  (repeat --model for a multi-mesh scene)\n"
              "                      [--technique phantom|lss|dots] [--size WxH] [--spp N] [--debug-primid | --material]\n"
              "                      [--frames N] [--in-flight] [--no-image | --no-hits] [--ppm out.ppm] [--png out.png] [--hits out.bin] [--device D] [--gpus N]\n"
              "                      [--env procedural|file.hdr] [--ao N] [--lod split,merge,curve_merge]");
}

int main(int argc, char** argv)
{
    std::string model = "synthetic:curly:10000:16", ppm, png, hits_path, technique = "lss";
    std::vector<std::string> models;          // --model may repeat: the reference's scene is a list of models (renderer.cpp:33-41)
    unsigned lod[3] = {0, 0, 0};
    RendererInitInfo info;
    int frames = 1, device = 0, gpus = 1;
    bool inFlight = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(2); } return argv[++i]; };
        if (a == "--help" || a == "-h") { usage(); return 0; }
        else if (a == "--model") { model = next(); models.push_back(model); }
        else if (a == "--technique") technique = next();
        else if (a == "--size") { if (std::sscanf(next(), "%ux%u", &info.width, &info.height) != 2) { usage(); return 2; } }
        else if (a == "--spp") info.spp = (uint32_t)std::atoi(next());
        else if (a == "--debug-primid") info.shadeMode = VKHRT_SHADE_DEBUG_PRIMID;
        else if (a == "--material") info.shadeMode = VKHRT_SHADE_MATERIAL;     // Shade(normal) * the asset's albedo factor
        else if (a == "--no-image") info.wantImage = false;                   // hit records only (BASELINE configs[1])
        else if (a == "--no-hits") info.wantHits = false;
        else if (a == "--frames") frames = std::atoi(next());
        else if (a == "--in-flight") inFlight = true;                        // keep VKHRT_FRAMES_IN_FLIGHT frames submitted (renderer.cpp:85-119)
        else if (a == "--ppm") ppm = next();
        else if (a == "--png") png = next();
        else if (a == "--env") info.environmentMap = next();
        else if (a == "--ao") info.aoSamples = (uint32_t)std::atoi(next());
        else if (a == "--lod") { if (std::sscanf(next(), "%u,%u,%u", &lod[0], &lod[1], &lod[2]) != 3) { usage(); return 2; } }
        else if (a == "--hits") hits_path = next();
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--gpus") gpus = std::atoi(next());            // devices device .. device+N-1 share every frame (one process, vkhrt_render_multi)
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); usage(); return 2; }
    }
    const VkhrtTechnique tech = technique == "phantom" ? VKHRT_TECHNIQUE_PHANTOM : (technique == "dots" ? VKHRT_TECHNIQUE_DOTS : VKHRT_TECHNIQUE_LSS);
    try {
        ModelLoader loader(device);
        auto t0 = std::chrono::steady_clock::now();
        std::shared_ptr<Model> m = models.size() > 1 ? loader.LoadFromFiles(models, tech) : loader.LoadFromFile(model, tech, lod[0], lod[1], lod[2]);
        if (!m) return 1;
        if (models.size() > 1) { model.clear(); for (const std::string& p : models) model += (model.empty() ? "" : " + ") + p; }
        const VkhrtTiming bt = m->Timing();
        std::printf("[MODEL LOADING] %s: %u primitives (%s), load+build %.1f ms wall, device build %.2f ms (sort %.2f, hierarchy %.2f, refit %.2f)\n",
                    model.c_str(), m->PrimitiveCount(), technique.c_str(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), bt.build_total_ms, bt.sort_ms,
                    bt.hierarchy_ms, bt.refit_ms);
        FlyCameraCreation cc;
        cc.aspectRatio = (float)info.width / (float)info.height;
        auto camera = std::make_shared<FlyCamera>(cc);
        Renderer renderer(info, camera);
        renderer.AddModel(m);
        for (int g = 1; g < gpus; ++g) {
            std::shared_ptr<Model> r = models.size() > 1 ? ModelLoader(device + g).LoadFromFiles(models, tech) : ModelLoader(device + g).LoadFromFile(model, tech, lod[0], lod[1], lod[2]);
            if (!r) return 1;
            renderer.AddReplica(r);
        }
        if (inFlight) {
            // the reference's frame loop: submit, and wait for the oldest frame once VKHRT_FRAMES_IN_FLIGHT are outstanding
            for (int k = 0; k < VKHRT_FRAMES_IN_FLIGHT; ++k) renderer.Submit();      // untimed: the page-locked buffer sets are allocated here
            for (int k = 0; k < VKHRT_FRAMES_IN_FLIGHT; ++k) renderer.Wait();
            auto t0 = std::chrono::steady_clock::now();
            int last = 0, outstanding = 0;
            for (int f = 0; f < frames; ++f) {
                if (outstanding == VKHRT_FRAMES_IN_FLIGHT) { last = renderer.Wait(); --outstanding; }
                renderer.Submit(); ++outstanding;
            }
            while (outstanding-- > 0) last = renderer.Wait();
            const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            std::printf("%d frames, %d in flight: %.3f ms per frame wall  (%.1f Mrays/s end to end)\n", frames, VKHRT_FRAMES_IN_FLIGHT, wall / frames,
                        (double)info.width * info.height * info.spp * frames / (wall * 1e3));
            size_t n_hit = 0;
            for (const VkhrtHit& h : renderer.GetFlightHits(last)) n_hit += h.flags & 1u;
            std::printf("hits: %zu of %zu rays\n", n_hit, renderer.GetFlightHits(last).size());
            if (!hits_path.empty()) {
                std::FILE* fp = std::fopen(hits_path.c_str(), "wb");
                if (!fp) { std::fprintf(stderr, "[FILE] cannot write %s\n", hits_path.c_str()); return 1; }
                std::fwrite(renderer.GetFlightHits(last).data(), sizeof(VkhrtHit), renderer.GetFlightHits(last).size(), fp);
                std::fclose(fp);
            }
            return 0;
        }
        for (int f = 0; f < frames; ++f) {
            auto f0 = std::chrono::steady_clock::now();
            renderer.Render();
            const VkhrtTiming t = m->Timing();
            const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - f0).count();
            std::printf("frame %d: %.3f ms wall, trace %.3f ms, shade %.3f ms, d2h %.3f ms  (%.1f Mrays/s device)\n", f, wall, t.trace_ms, t.shade_ms,
                        t.d2h_ms, (double)info.width * info.height / (t.trace_ms * 1e3));
        }
        size_t n_hit = 0;
        for (const VkhrtHit& h : renderer.GetHits()) n_hit += h.flags & 1u;
        std::printf("hits: %zu of %zu rays\n", n_hit, renderer.GetHits().size());
        if (m->MeshCount() > 1) {
            std::vector<size_t> per_mesh(m->MeshCount(), 0);
            for (const VkhrtHit& h : renderer.GetHits()) if (h.flags & 1u) per_mesh[m->MeshOfSegment(h.segment)]++;
            for (size_t k = 0; k < per_mesh.size(); ++k) std::printf("  mesh %zu: %zu rays\n", k, per_mesh[k]);
        }
        if (!ppm.empty() && !renderer.WritePPM(ppm)) { std::fprintf(stderr, "[FILE] cannot write %s\n", ppm.c_str()); return 1; }
        if (!png.empty() && !renderer.WritePNG(png)) { std::fprintf(stderr, "[FILE] cannot write %s\n", png.c_str()); return 1; }
        if (!hits_path.empty()) {
            std::FILE* fp = std::fopen(hits_path.c_str(), "wb");
            if (!fp) { std::fprintf(stderr, "[FILE] cannot write %s\n", hits_path.c_str()); return 1; }
            std::fwrite(renderer.GetHits().data(), sizeof(VkhrtHit), renderer.GetHits().size(), fp);
            std::fclose(fp);
        }
    } catch (const VkhrtError& e) {
        std::fprintf(stderr, "[VKHRT] %s\n", e.what());
        return 3;
    }
    return 0;
}
