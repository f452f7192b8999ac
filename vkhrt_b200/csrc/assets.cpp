// assets.cpp — host-side asset ingest and image output (plain C++, no CUDA): the data formats either side of the hot path.
//
// Replaces (reference paths), SURVEY.md §8(f) rows 3 and 4:
//   source/resources/model/model_loader.cpp:139-206, 274-291  ModelLoader::LoadFromFile / ProcessMesh for LINE primitives
//       (Assimp 6.0.1 is an un-vendored FetchContent dependency, external/CMakeLists.txt:77-97): here two line-asset formats
//       are read directly and produce the contract GenerateLines consumes (geometry_processor.cpp:45-67) — vertex positions +
//       uint32 index pairs, consecutive segments of a strand sharing an identical end/start position:
//         .obj   Wavefront `v x y z` + `l i j k ...` polyline records (what Assimp's OBJ importer turns into 2-index faces)
//         .hair  Cem Yuksel's HAIR binary format (128-byte header, per-strand segment counts, points, optional thickness)
//   source/resources/file_io.cpp:22-37  LoadFloatImageFromFile = stbi_loadf(path, .., 4): Radiance .hdr (RGBE, flat or
//       new-style RLE scanlines) -> RGBA32F, alpha 1 (the environment map of source/renderer.cpp:45-56)
//   the swap-chain present of source/renderer.cpp:222-231 -> an 8-bit PNG file (stored deflate blocks; no zlib dependency)
#include "../../include/vkhrt_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

namespace vkhrt {
void set_last_error(const std::string& s);
int load_gltf(const std::string& path, const std::vector<unsigned char>& data, bool glb, VkhrtLineAsset* out);   // gltf.cpp
int save_glb(const char* path, const VkhrtLineAsset* in);
}
using vkhrt::set_last_error;

namespace {

bool ends_with(const std::string& s, const char* suffix)
{
    std::string t(suffix);
    if (s.size() < t.size()) return false;
    for (size_t i = 0; i < t.size(); ++i)
        if (std::tolower((unsigned char)s[s.size() - t.size() + i]) != t[i]) return false;
    return true;
}

bool read_file(const char* path, std::vector<unsigned char>& out)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return false; }
    out.resize((size_t)n);
    size_t got = n ? std::fread(out.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

template <typename T>
T* dup_array(const std::vector<T>& v)
{
    T* p = (T*)std::malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

int fail(int code, const std::string& msg) { set_last_error(msg); return code; }

// ---- Wavefront OBJ: `v` and `l` records ------------------------------------------------------------
int load_obj(const std::vector<unsigned char>& data, VkhrtLineAsset* out)
{
    std::vector<float> pos;
    std::vector<uint32_t> idx;
    uint32_t strands = 0;
    const char* p = (const char*)data.data();
    const char* end = p + data.size();
    size_t line_no = 0;
    while (p < end) {
        const char* eol = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        std::string line(p, eol);
        p = eol + 1;
        ++line_no;
        size_t h = line.find('#');
        if (h != std::string::npos) line.resize(h);
        const char* s = line.c_str();
        while (*s == ' ' || *s == '\t') ++s;
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            char* q = nullptr;
            float v[3];
            const char* c = s + 1;
            for (int k = 0; k < 3; ++k) {
                v[k] = std::strtof(c, &q);
                if (q == c) return fail(VKHRT_ERR_IO, "obj: malformed vertex at line " + std::to_string(line_no));
                c = q;
            }
            pos.insert(pos.end(), v, v + 3);
        } else if (s[0] == 'l' && (s[1] == ' ' || s[1] == '\t')) {
            const char* c = s + 1;
            long prev = -1;
            uint32_t count = 0;
            for (;;) {
                char* q = nullptr;
                long i = std::strtol(c, &q, 10);
                if (q == c) break;
                c = q;
                if (*c == '/') { while (*c && *c != ' ' && *c != '\t') ++c; }   // `l v/vt` form: texture index ignored
                const long nv = (long)(pos.size() / 3);
                long vi = i > 0 ? i - 1 : nv + i;     // 1-based, negative = relative to the vertices read so far
                if (i == 0 || vi < 0 || vi >= nv) return fail(VKHRT_ERR_BAD_TOPOLOGY, "obj: line index out of range at line " + std::to_string(line_no));
                if (prev >= 0) { idx.push_back((uint32_t)prev); idx.push_back((uint32_t)vi); }
                prev = vi;
                ++count;
            }
            if (count >= 2) ++strands;
        }
    }
    out->n_vertices = (uint32_t)(pos.size() / 3);
    out->n_segments = (uint32_t)(idx.size() / 2);
    out->n_strands = strands;
    out->positions_xyz = dup_array(pos);
    out->line_indices = dup_array(idx);
    out->radius_per_vertex = nullptr;
    return VKHRT_OK;
}

// strands = maximal runs of segments (k, k+1) chained by index
void strand_runs(const VkhrtLineAsset* a, std::vector<std::pair<uint32_t, uint32_t>>& runs)
{
    uint32_t s = 0;
    while (s < a->n_segments) {
        uint32_t e = s + 1;
        while (e < a->n_segments && a->line_indices[2 * e] == a->line_indices[2 * (e - 1) + 1]) ++e;
        runs.emplace_back(s, e);
        s = e;
    }
}

int save_obj(const char* path, const VkhrtLineAsset* a)
{
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(VKHRT_ERR_IO, std::string("cannot open for writing: ") + path);
    std::fprintf(f, "# vkhrt_b200 line asset: %u vertices, %u segments\n", a->n_vertices, a->n_segments);
    for (uint32_t i = 0; i < a->n_vertices; ++i)
        std::fprintf(f, "v %.9g %.9g %.9g\n", a->positions_xyz[3 * i], a->positions_xyz[3 * i + 1], a->positions_xyz[3 * i + 2]);   // %.9g round-trips fp32
    std::vector<std::pair<uint32_t, uint32_t>> runs;
    strand_runs(a, runs);
    for (auto& r : runs) {
        std::fprintf(f, "l %u", a->line_indices[2 * r.first] + 1);
        for (uint32_t s = r.first; s < r.second; ++s) std::fprintf(f, " %u", a->line_indices[2 * s + 1] + 1);
        std::fprintf(f, "\n");
    }
    bool ok = std::fclose(f) == 0;
    return ok ? VKHRT_OK : fail(VKHRT_ERR_IO, std::string("write failed: ") + path);
}

// ---- Cem Yuksel HAIR format ------------------------------------------------------------------------
struct HairHeader {
    char magic[4];
    uint32_t n_strands, n_points, flags, default_segments;
    float default_thickness, default_transparency, default_color[3];
    char info[88];
};
static_assert(sizeof(HairHeader) == 128, "HAIR header is 128 bytes");
enum : uint32_t { HAIR_SEGMENTS = 1u, HAIR_POINTS = 2u, HAIR_THICKNESS = 4u, HAIR_TRANSPARENCY = 8u, HAIR_COLOR = 16u };

int load_hair(const std::vector<unsigned char>& data, VkhrtLineAsset* out)
{
    if (data.size() < sizeof(HairHeader)) return fail(VKHRT_ERR_IO, "hair: file shorter than the 128-byte header");
    HairHeader h;
    std::memcpy(&h, data.data(), sizeof(h));
    if (std::memcmp(h.magic, "HAIR", 4) != 0) return fail(VKHRT_ERR_IO, "hair: bad magic");
    if (!(h.flags & HAIR_POINTS)) return fail(VKHRT_ERR_IO, "hair: file has no points array");
    size_t off = sizeof(HairHeader);
    const size_t need = off + ((h.flags & HAIR_SEGMENTS) ? (size_t)h.n_strands * 2 : 0) + (size_t)h.n_points * 12 +
                        ((h.flags & HAIR_THICKNESS) ? (size_t)h.n_points * 4 : 0);
    if (data.size() < need) return fail(VKHRT_ERR_IO, "hair: truncated file");
    std::vector<uint32_t> idx;
    uint64_t first = 0;
    for (uint32_t s = 0; s < h.n_strands; ++s) {
        uint32_t segs = h.default_segments;
        if (h.flags & HAIR_SEGMENTS) { uint16_t v; std::memcpy(&v, data.data() + off + 2 * (size_t)s, 2); segs = v; }
        if (first + segs + 1 > h.n_points) return fail(VKHRT_ERR_BAD_TOPOLOGY, "hair: segment counts exceed the points array");
        for (uint32_t k = 0; k < segs; ++k) { idx.push_back((uint32_t)(first + k)); idx.push_back((uint32_t)(first + k + 1)); }
        first += (uint64_t)segs + 1;
    }
    if (h.flags & HAIR_SEGMENTS) off += (size_t)h.n_strands * 2;
    std::vector<float> pos((size_t)h.n_points * 3);
    if (h.n_points) std::memcpy(pos.data(), data.data() + off, pos.size() * 4);
    off += pos.size() * 4;
    out->radius_per_vertex = nullptr;
    if (h.flags & HAIR_THICKNESS) {
        // thickness = strand diameter; the pipeline wants a radius per vertex
        std::vector<float> r(h.n_points);
        if (h.n_points) std::memcpy(r.data(), data.data() + off, r.size() * 4);
        for (float& v : r) v *= 0.5f;
        out->radius_per_vertex = dup_array(r);
    }
    out->n_vertices = h.n_points;
    out->n_segments = (uint32_t)(idx.size() / 2);
    out->n_strands = h.n_strands;
    out->positions_xyz = dup_array(pos);
    out->line_indices = dup_array(idx);
    return VKHRT_OK;
}

int save_hair(const char* path, const VkhrtLineAsset* a)
{
    // HAIR stores strands as consecutive point runs: the asset must be in that shape (index pairs (k, k+1))
    std::vector<std::pair<uint32_t, uint32_t>> runs;
    strand_runs(a, runs);
    std::vector<uint16_t> segs;
    uint64_t expect = 0;
    for (auto& r : runs) {
        if (r.second - r.first > 65535u) return fail(VKHRT_ERR_UNSUPPORTED, "hair: more than 65535 segments in one strand");
        for (uint32_t s = r.first; s < r.second; ++s)
            if (a->line_indices[2 * s] != expect + (s - r.first) || a->line_indices[2 * s + 1] != expect + (s - r.first) + 1)
                return fail(VKHRT_ERR_UNSUPPORTED, "hair: strands must be consecutive vertex runs");
        segs.push_back((uint16_t)(r.second - r.first));
        expect += (uint64_t)(r.second - r.first) + 1;
    }
    if (expect != a->n_vertices) return fail(VKHRT_ERR_UNSUPPORTED, "hair: vertices not referenced by any strand");
    HairHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, "HAIR", 4);
    h.n_strands = (uint32_t)runs.size();
    h.n_points = a->n_vertices;
    h.flags = HAIR_SEGMENTS | HAIR_POINTS | (a->radius_per_vertex ? HAIR_THICKNESS : 0u);
    h.default_segments = 0;
    h.default_thickness = 2.0f * VKHRT_DEFAULT_RADIUS;
    h.default_transparency = 0.0f;
    h.default_color[0] = h.default_color[1] = h.default_color[2] = 1.0f;
    std::snprintf(h.info, sizeof(h.info), "vkhrt_b200");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(VKHRT_ERR_IO, std::string("cannot open for writing: ") + path);
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    if (ok && !segs.empty()) ok = std::fwrite(segs.data(), 2, segs.size(), f) == segs.size();
    if (ok && a->n_vertices) ok = std::fwrite(a->positions_xyz, 12, a->n_vertices, f) == a->n_vertices;
    if (ok && a->radius_per_vertex) {
        std::vector<float> t(a->radius_per_vertex, a->radius_per_vertex + a->n_vertices);
        for (float& v : t) v *= 2.0f;
        ok = std::fwrite(t.data(), 4, t.size(), f) == t.size();
    }
    ok = (std::fclose(f) == 0) && ok;
    return ok ? VKHRT_OK : fail(VKHRT_ERR_IO, std::string("write failed: ") + path);
}

// ---- Radiance .hdr (RGBE) -----------------------------------------------------------------------------
// decode one RGBE texel the way stbi_loadf does (stb_image.h stbi__hdr_convert, req_comp = 4): value = mantissa * 2^(e - 136), alpha 1
void rgbe_to_float(const unsigned char* in, float* out)
{
    if (in[3] != 0) {
        float f1 = std::ldexp(1.0f, (int)in[3] - (128 + 8));
        out[0] = in[0] * f1; out[1] = in[1] * f1; out[2] = in[2] * f1;
    } else out[0] = out[1] = out[2] = 0.0f;
    out[3] = 1.0f;
}

void float_to_rgbe(const float* in, unsigned char* out)
{
    float v = std::max(in[0], std::max(in[1], in[2]));
    if (!(v >= 1e-32f)) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    int e;
    float m = std::frexp(v, &e) * 256.0f / v;
    out[0] = (unsigned char)(in[0] * m); out[1] = (unsigned char)(in[1] * m); out[2] = (unsigned char)(in[2] * m);
    out[3] = (unsigned char)(e + 128);
}

int load_hdr(const std::vector<unsigned char>& d, float** rgba_out, uint32_t* w_out, uint32_t* h_out)
{
    size_t p = 0;
    auto next_line = [&](std::string& s) -> bool {
        if (p >= d.size()) return false;
        size_t e = p;
        while (e < d.size() && d[e] != '\n') ++e;
        s.assign((const char*)d.data() + p, e - p);
        if (!s.empty() && s.back() == '\r') s.pop_back();
        p = e + 1;
        return true;
    };
    std::string line;
    if (!next_line(line) || (line != "#?RADIANCE" && line != "#?RGBE")) return fail(VKHRT_ERR_IO, "hdr: not a Radiance file");
    bool format_ok = false;
    for (;;) {
        if (!next_line(line)) return fail(VKHRT_ERR_IO, "hdr: truncated header");
        if (line.empty()) break;
        if (line == "FORMAT=32-bit_rle_rgbe") format_ok = true;
    }
    if (!format_ok) return fail(VKHRT_ERR_UNSUPPORTED, "hdr: only FORMAT=32-bit_rle_rgbe is supported");
    if (!next_line(line)) return fail(VKHRT_ERR_IO, "hdr: missing resolution line");
    int H = 0, W = 0;
    if (std::sscanf(line.c_str(), "-Y %d +X %d", &H, &W) != 2 || H <= 0 || W <= 0) return fail(VKHRT_ERR_UNSUPPORTED, "hdr: only '-Y h +X w' orientation is supported");
    float* img = (float*)std::malloc((size_t)W * H * 4 * sizeof(float));
    if (!img) return fail(VKHRT_ERR_OUT_OF_MEMORY, "hdr: out of host memory");
    std::vector<unsigned char> scan((size_t)W * 4);
    for (int y = 0; y < H; ++y) {
        bool rle = false;
        if (W >= 8 && W < 32768 && p + 4 <= d.size() && d[p] == 2 && d[p + 1] == 2 && !(d[p + 2] & 0x80)) {
            if (((int)d[p + 2] << 8 | d[p + 3]) == W) rle = true;
        }
        if (rle) {
            p += 4;
            for (int c = 0; c < 4; ++c) {
                int x = 0;
                while (x < W) {
                    if (p >= d.size()) { std::free(img); return fail(VKHRT_ERR_IO, "hdr: truncated scanline"); }
                    unsigned char count = d[p++];
                    if (count > 128) {
                        count -= 128;
                        if (p >= d.size() || x + count > W) { std::free(img); return fail(VKHRT_ERR_IO, "hdr: corrupt run"); }
                        unsigned char v = d[p++];
                        for (int k = 0; k < count; ++k) scan[(size_t)(x++) * 4 + c] = v;
                    } else {
                        if (count == 0 || p + count > d.size() || x + count > W) { std::free(img); return fail(VKHRT_ERR_IO, "hdr: corrupt literal run"); }
                        for (int k = 0; k < count; ++k) scan[(size_t)(x++) * 4 + c] = d[p++];
                    }
                }
            }
        } else {
            if (p + (size_t)W * 4 > d.size()) { std::free(img); return fail(VKHRT_ERR_IO, "hdr: truncated flat scanline"); }
            std::memcpy(scan.data(), d.data() + p, (size_t)W * 4);
            p += (size_t)W * 4;
        }
        for (int x = 0; x < W; ++x) rgbe_to_float(&scan[(size_t)x * 4], img + ((size_t)y * W + x) * 4);
    }
    *rgba_out = img; *w_out = (uint32_t)W; *h_out = (uint32_t)H;
    return VKHRT_OK;
}

// ---- PNG (8-bit RGBA, stored deflate blocks) ----------------------------------------------------------------
uint32_t crc_table[256];
bool crc_ready = false;
uint32_t crc32(uint32_t crc, const unsigned char* p, size_t n)
{
    if (!crc_ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            crc_table[i] = c;
        }
        crc_ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = crc_table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return ~crc;
}
void put_be32(std::vector<unsigned char>& v, uint32_t x) { v.push_back(x >> 24); v.push_back((x >> 16) & 255); v.push_back((x >> 8) & 255); v.push_back(x & 255); }
void put_chunk(std::vector<unsigned char>& out, const char* type, const std::vector<unsigned char>& body)
{
    put_be32(out, (uint32_t)body.size());
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), body.begin(), body.end());
    put_be32(out, crc32(0, out.data() + start, out.size() - start));
}

}  // namespace

extern "C" {

int vkhrt_asset_load_lines(const char* path, VkhrtLineAsset* out)
{
    if (!path || !out) return fail(VKHRT_ERR_INVALID_ARGUMENT, "null argument");
    std::memset(out, 0, sizeof(*out));
    for (int k = 0; k < 4; ++k) out->base_color[k] = 1.0f;                 // formats without materials (.obj polylines, .hair)
    std::vector<unsigned char> data;
    if (!read_file(path, data)) return fail(VKHRT_ERR_IO, std::string("cannot read ") + path);
    std::string p(path);
    int rc;
    try {
        if (ends_with(p, ".obj")) rc = load_obj(data, out);
        else if (ends_with(p, ".hair")) rc = load_hair(data, out);
        else if (ends_with(p, ".gltf")) rc = vkhrt::load_gltf(p, data, false, out);
        else if (ends_with(p, ".glb")) rc = vkhrt::load_gltf(p, data, true, out);
        else return fail(VKHRT_ERR_UNSUPPORTED, "unknown line-asset extension (supported: .obj, .hair, .gltf, .glb)");
    } catch (const std::exception& e) {          // nothing may unwind across the C boundary (a hostile file can ask for any amount of memory)
        vkhrt_asset_free(out);
        return fail(VKHRT_ERR_OUT_OF_MEMORY, std::string("asset too large for host memory: ") + e.what());
    }
    if (rc == VKHRT_OK && (!out->positions_xyz || !out->line_indices)) { vkhrt_asset_free(out); return fail(VKHRT_ERR_OUT_OF_MEMORY, "out of host memory"); }
    return rc;
}

int vkhrt_asset_save_lines(const char* path, const VkhrtLineAsset* in)
{
    if (!path || !in || (in->n_vertices && !in->positions_xyz) || (in->n_segments && !in->line_indices)) return fail(VKHRT_ERR_INVALID_ARGUMENT, "null argument");
    for (uint32_t i = 0; i < 2 * in->n_segments; ++i)
        if (in->line_indices[i] >= in->n_vertices) return fail(VKHRT_ERR_BAD_TOPOLOGY, "line index out of range");
    std::string p(path);
    if (ends_with(p, ".obj")) return save_obj(path, in);
    if (ends_with(p, ".hair")) return save_hair(path, in);
    if (ends_with(p, ".glb")) return vkhrt::save_glb(path, in);
    return fail(VKHRT_ERR_UNSUPPORTED, "unknown line-asset extension (supported: .obj, .hair, .glb)");
}

void vkhrt_asset_free(VkhrtLineAsset* a)
{
    if (!a) return;
    std::free(a->positions_xyz); std::free(a->line_indices); std::free(a->radius_per_vertex);
    std::memset(a, 0, sizeof(*a));
}

int vkhrt_image_load_hdr(const char* path, float** rgba_out, uint32_t* width_out, uint32_t* height_out)
{
    if (!path || !rgba_out || !width_out || !height_out) return fail(VKHRT_ERR_INVALID_ARGUMENT, "null argument");
    *rgba_out = nullptr; *width_out = *height_out = 0;
    std::vector<unsigned char> data;
    if (!read_file(path, data)) return fail(VKHRT_ERR_IO, std::string("cannot read ") + path);
    return load_hdr(data, rgba_out, width_out, height_out);
}

int vkhrt_image_save_hdr(const char* path, const float* rgba, uint32_t width, uint32_t height)
{
    if (!path || !rgba || !width || !height) return fail(VKHRT_ERR_INVALID_ARGUMENT, "null argument");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(VKHRT_ERR_IO, std::string("cannot open for writing: ") + path);
    std::fprintf(f, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %u +X %u\n", height, width);
    std::vector<unsigned char> row((size_t)width * 4);
    bool ok = true;
    for (uint32_t y = 0; y < height && ok; ++y) {
        for (uint32_t x = 0; x < width; ++x) float_to_rgbe(rgba + ((size_t)y * width + x) * 4, &row[(size_t)x * 4]);
        // flat scanlines; a first texel that would read as the RLE marker (2,2,hi<128,..) gets a harmless mantissa nudge
        if (width >= 8 && width < 32768 && row[0] == 2 && row[1] == 2 && !(row[2] & 0x80)) row[0] = 3;
        ok = std::fwrite(row.data(), 1, row.size(), f) == row.size();
    }
    ok = (std::fclose(f) == 0) && ok;
    return ok ? VKHRT_OK : fail(VKHRT_ERR_IO, std::string("write failed: ") + path);
}

void vkhrt_image_free(float* rgba) { std::free(rgba); }

// OpenEXR 2.0 single-part scanline file, NO_COMPRESSION, four 32-bit FLOAT channels.  Channels are stored in alphabetical order
// (A, B, G, R), one scanline per block, offsets table in front of the pixel data; everything little-endian.
int vkhrt_image_save_exr(const char* path, const float* rgba, uint32_t width, uint32_t height)
{
    if (!path || !rgba || !width || !height) return fail(VKHRT_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<unsigned char> out;
    auto put32 = [&](uint32_t v) { for (int k = 0; k < 4; ++k) out.push_back((unsigned char)(v >> (8 * k))); };
    auto put64 = [&](uint64_t v) { for (int k = 0; k < 8; ++k) out.push_back((unsigned char)(v >> (8 * k))); };
    auto putf = [&](float f) { uint32_t u; std::memcpy(&u, &f, 4); put32(u); };
    auto puts = [&](const char* z) { while (*z) out.push_back((unsigned char)*z++); out.push_back(0); };
    auto attr = [&](const char* name, const char* type, uint32_t size) { puts(name); puts(type); put32(size); };
    put32(20000630u);                    // magic 76 2f 31 01
    put32(2u);                           // version 2, no flags: scanline, single part
    attr("channels", "chlist", 4 * 18 + 1);
    for (const char* ch : {"A", "B", "G", "R"}) { puts(ch); put32(2u /* FLOAT */); put32(0u /* pLinear + reserved */); put32(1u); put32(1u); }
    out.push_back(0);
    attr("compression", "compression", 1); out.push_back(0);
    attr("dataWindow", "box2i", 16); put32(0); put32(0); put32(width - 1); put32(height - 1);
    attr("displayWindow", "box2i", 16); put32(0); put32(0); put32(width - 1); put32(height - 1);
    attr("lineOrder", "lineOrder", 1); out.push_back(0);
    attr("pixelAspectRatio", "float", 4); putf(1.0f);
    attr("screenWindowCenter", "v2f", 8); putf(0.0f); putf(0.0f);
    attr("screenWindowWidth", "float", 4); putf(1.0f);
    out.push_back(0);                    // end of header
    const uint64_t row_bytes = (uint64_t)width * 16, block = 8 + row_bytes;
    const uint64_t first = out.size() + (uint64_t)height * 8;
    for (uint32_t y = 0; y < height; ++y) put64(first + (uint64_t)y * block);
    const size_t header = out.size();
    out.resize(header + (size_t)(block * height));
    unsigned char* q = out.data() + header;
    static const int order[4] = {3, 2, 1, 0};           // A B G R from RGBA
    for (uint32_t y = 0; y < height; ++y) {
        const uint32_t yy = y, sz = (uint32_t)row_bytes;
        std::memcpy(q, &yy, 4); std::memcpy(q + 4, &sz, 4); q += 8;
        for (int c = 0; c < 4; ++c) {
            const float* src = rgba + (size_t)y * width * 4 + order[c];
            for (uint32_t x = 0; x < width; ++x) { std::memcpy(q, src + 4 * (size_t)x, 4); q += 4; }
        }
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(VKHRT_ERR_IO, std::string("cannot open for writing: ") + path);
    bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    ok = (std::fclose(f) == 0) && ok;
    return ok ? VKHRT_OK : fail(VKHRT_ERR_IO, std::string("write failed: ") + path);
}

int vkhrt_image_save_png(const char* path, const uint8_t* rgba8, uint32_t width, uint32_t height)
{
    if (!path || !rgba8 || !width || !height) return fail(VKHRT_ERR_INVALID_ARGUMENT, "null argument");
    std::vector<unsigned char> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, width); put_be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    put_chunk(png, "IHDR", ihdr);
    // raw scanlines, filter type 0
    const size_t stride = (size_t)width * 4 + 1;
    std::vector<unsigned char> raw(stride * height);
    for (uint32_t y = 0; y < height; ++y) {
        raw[y * stride] = 0;
        std::memcpy(&raw[y * stride + 1], rgba8 + (size_t)y * width * 4, (size_t)width * 4);
    }
    // zlib container around stored (uncompressed) deflate blocks
    std::vector<unsigned char> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;
    for (size_t off = 0; off < raw.size();) {
        size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n == raw.size() ? 1 : 0);
        z.push_back(n & 255); z.push_back((n >> 8) & 255); z.push_back(~n & 255); z.push_back((~n >> 8) & 255);
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = off; i < off + n; ++i) { a += raw[i]; if (a >= 65521u) a -= 65521u; b += a; if (b >= 65521u) b -= 65521u; }
        off += n;
    }
    put_be32(z, (b << 16) | a);
    put_chunk(png, "IDAT", z);
    put_chunk(png, "IEND", {});
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(VKHRT_ERR_IO, std::string("cannot open for writing: ") + path);
    bool ok = std::fwrite(png.data(), 1, png.size(), f) == png.size();
    ok = (std::fclose(f) == 0) && ok;
    return ok ? VKHRT_OK : fail(VKHRT_ERR_IO, std::string("write failed: ") + path);
}

// Procedural equirectangular sky (the reference's qwantani_sunset_puresky_4k.hdr is not in its repository and there is no
// network): horizon-to-zenith gradient, a darker ground half and a sun lobe, linear radiance, alpha 1.  Row 0 is what a ray pointing straight up sees.
void vkhrt_environment_generate(uint32_t width, uint32_t height, float* rgba_out)
{
    if (!rgba_out) return;
    const double pi = 3.14159265358979323846;
    const double sun_az = 0.6 * pi, sun_el = 0.25;
    const double sx = std::cos(sun_el) * std::sin(sun_az), sy = std::sin(sun_el), sz = -std::cos(sun_el) * std::cos(sun_az);
    for (uint32_t y = 0; y < height; ++y) {
        double v = (y + 0.5) / height;
        double el = (v - 0.5) * pi;                 // uv.y = asin(dir.y)/pi + 0.5 (miss.rmiss DirectionToUV)
        for (uint32_t x = 0; x < width; ++x) {
            double u = (x + 0.5) / width;
            double th = (u - 0.5) * 2.0 * pi;       // uv.x = atan(dir.x, -dir.z)/(2 pi) + 0.5
            // (u, v) is where miss.rmiss looks up dir = normalize(-rayDirection); the sky is a function of the RAY direction
            double dx = -(std::cos(el) * std::sin(th)), dy = -std::sin(el), dz = std::cos(el) * std::cos(th);
            double up = std::max(0.0, dy), down = std::max(0.0, -dy);
            double r = 0.9 - 0.55 * up, g = 0.75 - 0.25 * up, b = 0.6 + 0.35 * up;
            if (dy < 0.0) { double k = 1.0 - 0.8 * std::min(1.0, down * 4.0); r = 0.35 * k + 0.05; g = 0.3 * k + 0.05; b = 0.25 * k + 0.05; }
            double c = std::max(0.0, dx * sx + dy * sy + dz * sz);
            double lobe = 6.0 * std::pow(c, 256.0) + 0.6 * std::pow(c, 8.0);
            float* o = rgba_out + ((size_t)y * width + x) * 4;
            o[0] = (float)(r + lobe); o[1] = (float)(g + 0.8 * lobe); o[2] = (float)(b + 0.5 * lobe); o[3] = 1.0f;
        }
    }
}

}  // extern "C"
