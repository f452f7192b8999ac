// glTF 2.0 line-primitive reader (.gltf with external or embedded buffers, .glb) -> the positions + index-pair contract of
// GenerateLines (reference: source/resources/model/geometry_processor.cpp:45-67).
//
// The reference's scene is three glTF files (`assets/claire/Claire_HairMain_less_strands.gltf`, ... source/renderer.cpp:33-37)
// imported by Assimp (`ModelLoader::LoadFromFile`, source/resources/model/model_loader.cpp:274-291; an un-vendored dependency,
// external/CMakeLists.txt:77-97).  What Assimp's importer hands to `ProcessMesh` (model_loader.cpp:139-206) for a line primitive is
// restated here directly on the file format:
//   * every mesh primitive with mode 1 (LINES), 2 (LINE_LOOP) or 3 (LINE_STRIP) becomes one mesh: its POSITION accessor (float
//     VEC3) in file order, and 2-index faces — the index pairs as stored (LINES), or (k, k+1) for consecutive indices (STRIP, plus
//     the closing pair for LOOP); without an index accessor the indices are 0..count-1;
//   * meshes are concatenated with `firstVertex` offsets exactly as ProcessMesh does (model_loader.cpp:143-165);
//   * the node hierarchy (ProcessNode, model_loader.cpp:208-246) carries a transform per mesh instance, which the reference applies
//     as the TLAS instance transform (source/top_level_acceleration_structure.cpp:25-29).  This boundary has one world-space scene,
//     so the node's world matrix is applied to the positions here (in double, rounded once to fp32; the identity leaves the file's
//     floats untouched).  A mesh referenced by several nodes is emitted once per node.
// Other primitive modes (points, triangles) are skipped, like the reference's line path skips them (geometry_processor.cpp:606-612).
// Not supported, reported as VKHRT_ERR_UNSUPPORTED: sparse accessors, required extensions (Draco, meshopt, ...), non-float positions.
// Host code only; every read is bounds-checked (truncated or corrupt files produce an error code, never a fault).
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/vkhrt_b200.h"

namespace vkhrt {
void set_last_error(const std::string& s);
int load_gltf(const std::string& path, const std::vector<unsigned char>& data, bool glb, VkhrtLineAsset* out);
int save_glb(const char* path, const VkhrtLineAsset* in);
}

namespace {

int fail(int code, const std::string& msg) { vkhrt::set_last_error(msg); return code; }

// ---- a small JSON reader (objects, arrays, strings, numbers, true/false/null) ------------------------
struct JVal {
    enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } type = NUL;
    double num = 0.0;
    bool b = false;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* key) const
    {
        if (type != OBJ) return nullptr;
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_num() const { return type == NUM; }
};

struct JParser {
    const char* p; const char* end; int depth = 0; bool ok = true;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool lit(const char* s) { size_t n = std::strlen(s); if ((size_t)(end - p) >= n && std::memcmp(p, s, n) == 0) { p += n; return true; } return false; }
    bool parse_string(std::string& out)
    {
        if (p >= end || *p != '"') return false;
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return false;
                switch (*p) {
                case 'n': out.push_back('\n'); break; case 't': out.push_back('\t'); break; case 'r': out.push_back('\r'); break;
                case 'b': out.push_back('\b'); break; case 'f': out.push_back('\f'); break;
                case 'u': { if (end - p < 5) return false; unsigned v = 0; for (int k = 1; k <= 4; ++k) { char c = p[k]; v <<= 4; if (c >= '0' && c <= '9') v |= (unsigned)(c - '0'); else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10); else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10); else return false; } out.push_back(v < 128 ? (char)v : '?'); p += 4; break; }
                default: out.push_back(*p); break;      // \" \\ \/
                }
                ++p;
            } else out.push_back(*p++);
        }
        if (p >= end) return false;
        ++p;
        return true;
    }
    bool parse(JVal& v)
    {
        if (++depth > 64) return false;
        ws();
        if (p >= end) return false;
        bool r = true;
        if (*p == '{') {
            v.type = JVal::OBJ; ++p; ws();
            if (p < end && *p == '}') ++p;
            else for (;;) {
                ws();
                std::string key;
                if (!parse_string(key)) { r = false; break; }
                ws();
                if (p >= end || *p != ':') { r = false; break; }
                ++p;
                v.obj.emplace_back(std::move(key), JVal());
                if (!parse(v.obj.back().second)) { r = false; break; }
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                r = false; break;
            }
        } else if (*p == '[') {
            v.type = JVal::ARR; ++p; ws();
            if (p < end && *p == ']') ++p;
            else for (;;) {
                v.arr.emplace_back();
                if (!parse(v.arr.back())) { r = false; break; }
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                r = false; break;
            }
        } else if (*p == '"') { v.type = JVal::STR; r = parse_string(v.str); }
        else if (lit("true")) { v.type = JVal::BOOL; v.b = true; }
        else if (lit("false")) { v.type = JVal::BOOL; v.b = false; }
        else if (lit("null")) { v.type = JVal::NUL; }
        else {
            // number: copy the token so strtod cannot run past the buffer
            const char* s = p;
            while (p < end && (std::strchr("+-0123456789.eE", *p) != nullptr)) ++p;
            if (p == s || p - s > 64) r = false;
            else { std::string tok(s, p); char* q = nullptr; v.num = std::strtod(tok.c_str(), &q); v.type = JVal::NUM; r = q && *q == 0; }
        }
        --depth;
        return r;
    }
};

bool as_index(const JVal* v, size_t limit, size_t* out)
{
    if (!v || !v->is_num() || !(v->num >= 0.0) || v->num != std::floor(v->num) || v->num >= (double)limit) return false;
    *out = (size_t)v->num;
    return true;
}
bool as_size(const JVal* v, size_t def, size_t* out)
{
    if (!v) { *out = def; return true; }
    if (!v->is_num() || !(v->num >= 0.0) || v->num != std::floor(v->num) || v->num > 1e15) return false;
    *out = (size_t)v->num;
    return true;
}

bool base64_decode(const char* s, size_t n, std::vector<unsigned char>& out)
{
    uint32_t acc = 0; int bits = 0;
    for (size_t i = 0; i < n; ++i) {
        const char c = s[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A'; else if (c >= 'a' && c <= 'z') v = c - 'a' + 26; else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62; else if (c == '/' || c == '_') v = 63; else if (c == '=' || c == '\n' || c == '\r') continue; else return false;
        acc = (acc << 6) | (uint32_t)v; bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((unsigned char)((acc >> bits) & 0xFFu)); }
    }
    return true;
}

bool read_whole_file(const std::string& path, std::vector<unsigned char>& out)
{
    FILE* f = std::fopen(path.c_str(), "rb");
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

// column-major 4x4 in double
struct M4 { double m[16]; };
M4 m4_identity() { M4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0; return r; }
M4 m4_mul(const M4& a, const M4& b)
{
    M4 r{};
    for (int c = 0; c < 4; ++c) for (int row = 0; row < 4; ++row) {
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s += a.m[k * 4 + row] * b.m[c * 4 + k];
        r.m[c * 4 + row] = s;
    }
    return r;
}
bool m4_is_identity(const M4& a) { const M4 i = m4_identity(); return std::memcmp(a.m, i.m, sizeof(a.m)) == 0; }
bool num_array(const JVal* v, size_t n, double* out)
{
    if (!v || v->type != JVal::ARR || v->arr.size() != n) return false;
    for (size_t i = 0; i < n; ++i) { if (!v->arr[i].is_num()) return false; out[i] = v->arr[i].num; }
    return true;
}
// node.matrix, or T * R * S (glTF 2.0 §3.5.3)
bool node_local(const JVal& node, M4* out)
{
    *out = m4_identity();
    if (const JVal* mj = node.get("matrix")) return num_array(mj, 16, out->m);
    double t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
    if (const JVal* v = node.get("translation")) if (!num_array(v, 3, t)) return false;
    if (const JVal* v = node.get("rotation")) if (!num_array(v, 4, q)) return false;
    if (const JVal* v = node.get("scale")) if (!num_array(v, 3, s)) return false;
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double r[9] = {1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w),        // column 0
                         2 * (x * y - z * w), 1 - 2 * (x * x + z * z), 2 * (y * z + x * w),        // column 1
                         2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y)};       // column 2
    for (int c = 0; c < 3; ++c) for (int row = 0; row < 3; ++row) out->m[c * 4 + row] = r[c * 3 + row] * s[c];
    out->m[12] = t[0]; out->m[13] = t[1]; out->m[14] = t[2];
    return true;
}

struct Gltf {
    JVal root;
    std::vector<std::vector<unsigned char>> buffers;
    std::vector<bool> loaded;
    std::string dir;
    const std::vector<unsigned char>* glb_bin = nullptr;
    size_t node_visits = 0;          // a hierarchy that is a DAG (or a cycle) is cut off instead of being expanded forever
};

int get_buffer(Gltf& g, size_t bi, const std::vector<unsigned char>** out)
{
    const JVal* bufs = g.root.get("buffers");
    if (!bufs || bufs->type != JVal::ARR || bi >= bufs->arr.size()) return fail(VKHRT_ERR_IO, "gltf: buffer index out of range");
    if (g.buffers.size() != bufs->arr.size()) { g.buffers.resize(bufs->arr.size()); g.loaded.assign(bufs->arr.size(), false); }
    if (!g.loaded[bi]) {
        const JVal& b = bufs->arr[bi];
        const JVal* uri = b.get("uri");
        if (!uri) {
            if (bi != 0 || !g.glb_bin) return fail(VKHRT_ERR_IO, "gltf: buffer without uri outside a .glb container");
            g.buffers[bi] = *g.glb_bin;
        } else {
            if (uri->type != JVal::STR) return fail(VKHRT_ERR_IO, "gltf: buffer uri is not a string");
            const std::string& u = uri->str;
            if (u.compare(0, 5, "data:") == 0) {
                size_t comma = u.find(',');
                if (comma == std::string::npos || u.find(";base64") == std::string::npos || u.find(";base64") > comma) return fail(VKHRT_ERR_UNSUPPORTED, "gltf: only base64 data URIs are supported");
                if (!base64_decode(u.c_str() + comma + 1, u.size() - comma - 1, g.buffers[bi])) return fail(VKHRT_ERR_IO, "gltf: malformed base64 buffer");
            } else {
                if (u.find("://") != std::string::npos) return fail(VKHRT_ERR_UNSUPPORTED, "gltf: remote buffer uri");
                std::string file;
                for (size_t i = 0; i < u.size(); ++i) {          // percent-decoding of the relative path
                    if (u[i] == '%' && i + 2 < u.size() && std::isxdigit((unsigned char)u[i + 1]) && std::isxdigit((unsigned char)u[i + 2])) { file.push_back((char)std::strtol(u.substr(i + 1, 2).c_str(), nullptr, 16)); i += 2; }
                    else file.push_back(u[i]);
                }
                if (!read_whole_file(g.dir + file, g.buffers[bi])) return fail(VKHRT_ERR_IO, "gltf: cannot read buffer file " + g.dir + file);
            }
        }
        g.loaded[bi] = true;
    }
    *out = &g.buffers[bi];
    return VKHRT_OK;
}

struct AccessorView { const unsigned char* base; size_t stride, count; int component_type; };

// resolves accessor -> bufferView -> buffer with every range checked; `elem_bytes` = size of one element
int view_accessor(Gltf& g, size_t ai, const char* want_type, AccessorView* v)
{
    const JVal* accs = g.root.get("accessors");
    if (!accs || accs->type != JVal::ARR || ai >= accs->arr.size()) return fail(VKHRT_ERR_IO, "gltf: accessor index out of range");
    const JVal& a = accs->arr[ai];
    if (a.get("sparse")) return fail(VKHRT_ERR_UNSUPPORTED, "gltf: sparse accessors are not supported");
    const JVal* type = a.get("type");
    if (!type || type->type != JVal::STR || type->str != want_type) return fail(VKHRT_ERR_UNSUPPORTED, std::string("gltf: accessor is not ") + want_type);
    const JVal* ct = a.get("componentType");
    if (!ct || !ct->is_num()) return fail(VKHRT_ERR_IO, "gltf: accessor without componentType");
    v->component_type = (int)ct->num;
    size_t comp;
    switch (v->component_type) { case 5120: case 5121: comp = 1; break; case 5122: case 5123: comp = 2; break; case 5125: case 5126: comp = 4; break; default: return fail(VKHRT_ERR_UNSUPPORTED, "gltf: unknown componentType"); }
    const size_t elem = comp * (std::strcmp(want_type, "VEC3") == 0 ? 3 : 1);
    size_t count, a_off, bv_i;
    if (!as_size(a.get("count"), 0, &count) || !as_size(a.get("byteOffset"), 0, &a_off)) return fail(VKHRT_ERR_IO, "gltf: malformed accessor");
    const JVal* bvs = g.root.get("bufferViews");
    if (!a.get("bufferView") || !bvs || bvs->type != JVal::ARR || !as_index(a.get("bufferView"), bvs->arr.size(), &bv_i)) return fail(VKHRT_ERR_UNSUPPORTED, "gltf: accessor without a valid bufferView");
    const JVal& bv = bvs->arr[bv_i];
    size_t b_i, bv_off, bv_len, stride;
    const JVal* bufs = g.root.get("buffers");
    if (!bufs || bufs->type != JVal::ARR || !as_index(bv.get("buffer"), bufs->arr.size(), &b_i) || !as_size(bv.get("byteOffset"), 0, &bv_off) || !bv.get("byteLength") || !as_size(bv.get("byteLength"), 0, &bv_len) ||
        !as_size(bv.get("byteStride"), 0, &stride)) return fail(VKHRT_ERR_IO, "gltf: malformed bufferView");
    if (stride == 0) stride = elem;
    if (stride < elem) return fail(VKHRT_ERR_IO, "gltf: byteStride smaller than the element");
    const std::vector<unsigned char>* buf = nullptr;
    int rc = get_buffer(g, b_i, &buf);
    if (rc) return rc;
    if (bv_off > buf->size() || bv_len > buf->size() - bv_off) return fail(VKHRT_ERR_IO, "gltf: bufferView exceeds its buffer");
    if (count) {
        // last element must end inside the view
        if (a_off > bv_len || (count - 1) > (bv_len - a_off) / stride || (count - 1) * stride + elem > bv_len - a_off) return fail(VKHRT_ERR_IO, "gltf: accessor exceeds its bufferView");
    }
    v->base = buf->data() + bv_off + a_off; v->stride = stride; v->count = count;
    return VKHRT_OK;
}

struct Builder { std::vector<float> pos; std::vector<uint32_t> idx; uint32_t strands = 0; float base_color[4] = {1, 1, 1, 1}; bool have_color = false; };

int emit_primitive(Gltf& g, const JVal& prim, const M4& world, Builder& out)
{
    size_t mode = 4;
    if (!as_size(prim.get("mode"), 4, &mode)) return fail(VKHRT_ERR_IO, "gltf: malformed primitive mode");
    if (mode < 1 || mode > 3) return VKHRT_OK;                            // not a line primitive
    if (!out.have_color) {
        // ProcessMaterial (model_loader.cpp:96-99): albedoFactor = AI_MATKEY_BASE_COLOR = pbrMetallicRoughness.baseColorFactor
        out.have_color = true;
        const JVal* mats = g.root.get("materials");
        size_t mi;
        if (mats && mats->type == JVal::ARR && prim.get("material") && as_index(prim.get("material"), mats->arr.size(), &mi))
            if (const JVal* pbr = mats->arr[mi].get("pbrMetallicRoughness"))
                if (const JVal* bc = pbr->get("baseColorFactor"))
                    if (bc->type == JVal::ARR && bc->arr.size() == 4 && bc->arr[0].is_num() && bc->arr[1].is_num() && bc->arr[2].is_num() && bc->arr[3].is_num())
                        for (int k = 0; k < 4; ++k) out.base_color[k] = (float)bc->arr[k].num;
    }
    const JVal* attrs = prim.get("attributes");
    const JVal* accs = g.root.get("accessors");
    size_t pa;
    if (!attrs || !accs || accs->type != JVal::ARR || !as_index(attrs->get("POSITION"), accs->arr.size(), &pa)) return fail(VKHRT_ERR_IO, "gltf: line primitive without POSITION");
    AccessorView pv;
    int rc = view_accessor(g, pa, "VEC3", &pv);
    if (rc) return rc;
    if (pv.component_type != 5126) return fail(VKHRT_ERR_UNSUPPORTED, "gltf: POSITION must be float");
    if (out.pos.size() / 3 + pv.count >= 0xFFFFFFFFull) return fail(VKHRT_ERR_UNSUPPORTED, "gltf: too many vertices");
    const uint32_t first = (uint32_t)(out.pos.size() / 3);
    const bool ident = m4_is_identity(world);
    for (size_t i = 0; i < pv.count; ++i) {
        float p[3];
        std::memcpy(p, pv.base + i * pv.stride, 12);
        if (!ident) {
            const double x = p[0], y = p[1], z = p[2];
            for (int r = 0; r < 3; ++r) p[r] = (float)(world.m[r] * x + world.m[4 + r] * y + world.m[8 + r] * z + world.m[12 + r]);
        }
        out.pos.insert(out.pos.end(), p, p + 3);
    }
    // indices
    std::vector<uint32_t> ind;
    if (prim.get("indices")) {
        size_t ia;
        if (!as_index(prim.get("indices"), accs->arr.size(), &ia)) return fail(VKHRT_ERR_IO, "gltf: index accessor out of range");
        AccessorView iv;
        rc = view_accessor(g, ia, "SCALAR", &iv);
        if (rc) return rc;
        if (iv.component_type != 5121 && iv.component_type != 5123 && iv.component_type != 5125) return fail(VKHRT_ERR_UNSUPPORTED, "gltf: indices must be unsigned");
        ind.resize(iv.count);
        for (size_t i = 0; i < iv.count; ++i) {
            const unsigned char* s = iv.base + i * iv.stride;
            uint32_t v;
            if (iv.component_type == 5121) v = s[0];
            else if (iv.component_type == 5123) { uint16_t h; std::memcpy(&h, s, 2); v = h; }
            else std::memcpy(&v, s, 4);
            if (v >= pv.count) return fail(VKHRT_ERR_BAD_TOPOLOGY, "gltf: line index out of range");
            ind[i] = v;
        }
    } else {
        ind.resize(pv.count);
        for (size_t i = 0; i < pv.count; ++i) ind[i] = (uint32_t)i;
    }
    auto seg = [&](uint32_t a, uint32_t b) { out.idx.push_back(first + a); out.idx.push_back(first + b); };
    const size_t before = out.idx.size();
    if (mode == 1) { for (size_t i = 0; i + 1 < ind.size(); i += 2) seg(ind[i], ind[i + 1]); }
    else {
        for (size_t i = 0; i + 1 < ind.size(); ++i) seg(ind[i], ind[i + 1]);
        if (mode == 2 && ind.size() > 2) seg(ind.back(), ind[0]);
    }
    // strands = maximal runs of segments chained by index (b of one == a of the next)
    for (size_t i = before; i < out.idx.size(); i += 2)
        if (i == before || out.idx[i] != out.idx[i - 1]) ++out.strands;
    return VKHRT_OK;
}

int walk_node(Gltf& g, size_t ni, const M4& parent, int depth, Builder& out)
{
    const JVal* nodes = g.root.get("nodes");
    if (depth > 64 || ++g.node_visits > (1u << 20)) return fail(VKHRT_ERR_IO, "gltf: node hierarchy too deep or too large (cycle?)");
    const JVal& node = nodes->arr[ni];
    M4 local;
    if (!node_local(node, &local)) return fail(VKHRT_ERR_IO, "gltf: malformed node transform");
    const M4 world = m4_mul(parent, local);
    if (node.get("mesh")) {
        const JVal* meshes = g.root.get("meshes");
        size_t mi;
        if (!meshes || meshes->type != JVal::ARR || !as_index(node.get("mesh"), meshes->arr.size(), &mi)) return fail(VKHRT_ERR_IO, "gltf: mesh index out of range");
        const JVal* prims = meshes->arr[mi].get("primitives");
        if (prims && prims->type == JVal::ARR)
            for (const JVal& pr : prims->arr) { int rc = emit_primitive(g, pr, world, out); if (rc) return rc; }
    }
    if (const JVal* ch = node.get("children")) {
        if (ch->type != JVal::ARR) return fail(VKHRT_ERR_IO, "gltf: malformed children");
        for (const JVal& c : ch->arr) {
            size_t ci;
            if (!as_index(&c, nodes->arr.size(), &ci)) return fail(VKHRT_ERR_IO, "gltf: child index out of range");
            int rc = walk_node(g, ci, world, depth + 1, out);
            if (rc) return rc;
        }
    }
    return VKHRT_OK;
}

template <typename T>
T* dup_array(const std::vector<T>& v)
{
    T* p = (T*)std::malloc((v.empty() ? 1 : v.size()) * sizeof(T));
    if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

}  // namespace

namespace vkhrt {

int load_gltf(const std::string& path, const std::vector<unsigned char>& data, bool glb, VkhrtLineAsset* out)
{
    Gltf g;
    size_t slash = path.find_last_of("/\\");
    g.dir = slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
    const char* json = (const char*)data.data();
    size_t json_len = data.size();
    std::vector<unsigned char> bin;
    if (glb) {
        // 12-byte header {magic 'glTF', version, length} + chunks {length, type, data}: JSON first, then optionally BIN
        if (data.size() < 20 || std::memcmp(data.data(), "glTF", 4) != 0) return fail(VKHRT_ERR_IO, "glb: bad header");
        uint32_t version, clen, ctype;
        std::memcpy(&version, data.data() + 4, 4);
        if (version != 2) return fail(VKHRT_ERR_UNSUPPORTED, "glb: only version 2 is supported");
        std::memcpy(&clen, data.data() + 12, 4); std::memcpy(&ctype, data.data() + 16, 4);
        if (ctype != 0x4E4F534Au || clen > data.size() - 20) return fail(VKHRT_ERR_IO, "glb: first chunk is not JSON");
        json = (const char*)data.data() + 20; json_len = clen;
        size_t off = 20 + (size_t)clen;
        off = (off + 3) & ~(size_t)3;
        if (off + 8 <= data.size()) {
            uint32_t blen, btype;
            std::memcpy(&blen, data.data() + off, 4); std::memcpy(&btype, data.data() + off + 4, 4);
            if (btype == 0x004E4942u) {
                if (blen > data.size() - off - 8) return fail(VKHRT_ERR_IO, "glb: BIN chunk exceeds the file");
                bin.assign(data.begin() + (long)off + 8, data.begin() + (long)off + 8 + blen);
                g.glb_bin = &bin;
            }
        }
    }
    JParser jp{json, json + json_len};
    if (!jp.parse(g.root) || g.root.type != JVal::OBJ) return fail(VKHRT_ERR_IO, "gltf: malformed JSON");
    if (const JVal* req = g.root.get("extensionsRequired"))
        if (req->type == JVal::ARR && !req->arr.empty()) return fail(VKHRT_ERR_UNSUPPORTED, "gltf: required extension " + (req->arr[0].type == JVal::STR ? req->arr[0].str : std::string("?")) + " is not supported");
    const JVal* nodes = g.root.get("nodes");
    Builder b;
    if (nodes && nodes->type == JVal::ARR && !nodes->arr.empty()) {
        std::vector<size_t> roots;
        const JVal* scenes = g.root.get("scenes");
        if (scenes && scenes->type == JVal::ARR && !scenes->arr.empty()) {
            size_t si = 0;
            if (g.root.get("scene") && !as_index(g.root.get("scene"), scenes->arr.size(), &si)) return fail(VKHRT_ERR_IO, "gltf: scene index out of range");
            const JVal* sn = scenes->arr[si].get("nodes");
            if (sn && sn->type == JVal::ARR)
                for (const JVal& r : sn->arr) { size_t ri; if (!as_index(&r, nodes->arr.size(), &ri)) return fail(VKHRT_ERR_IO, "gltf: scene node out of range"); roots.push_back(ri); }
        } else {
            std::vector<bool> is_child(nodes->arr.size(), false);
            for (const JVal& n : nodes->arr)
                if (const JVal* ch = n.get("children")) if (ch->type == JVal::ARR) for (const JVal& c : ch->arr) { size_t ci; if (as_index(&c, nodes->arr.size(), &ci)) is_child[ci] = true; }
            for (size_t i = 0; i < nodes->arr.size(); ++i) if (!is_child[i]) roots.push_back(i);
        }
        for (size_t r : roots) { int rc = walk_node(g, r, m4_identity(), 0, b); if (rc) return rc; }
    } else if (const JVal* meshes = g.root.get("meshes")) {
        // no node hierarchy: every mesh once, untransformed
        if (meshes->type == JVal::ARR)
            for (const JVal& m : meshes->arr)
                if (const JVal* prims = m.get("primitives")) if (prims->type == JVal::ARR)
                    for (const JVal& pr : prims->arr) { int rc = emit_primitive(g, pr, m4_identity(), b); if (rc) return rc; }
    }
    out->n_vertices = (uint32_t)(b.pos.size() / 3);
    out->n_segments = (uint32_t)(b.idx.size() / 2);
    out->n_strands = b.strands;
    out->positions_xyz = dup_array(b.pos);
    out->line_indices = dup_array(b.idx);
    out->radius_per_vertex = nullptr;
    for (int k = 0; k < 4; ++k) out->base_color[k] = b.base_color[k];
    return VKHRT_OK;
}

// Writer: one .glb with one mesh, one LINES primitive (float positions, u32 index pairs), one untransformed node — loadable by
// Assimp (i.e. by the reference) and by load_gltf above, which returns the very same arrays.
int save_glb(const char* path, const VkhrtLineAsset* in)
{
    const size_t pos_bytes = (size_t)in->n_vertices * 12, idx_bytes = (size_t)in->n_segments * 8;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (uint32_t i = 0; i < in->n_vertices; ++i)
        for (int k = 0; k < 3; ++k) {
            const float v = in->positions_xyz[3 * (size_t)i + k];
            if (i == 0 || v < lo[k]) lo[k] = v;
            if (i == 0 || v > hi[k]) hi[k] = v;
        }
    char js[1024];
    int n;
    if (in->n_vertices && in->n_segments)
        n = std::snprintf(js, sizeof(js),
            "{\"asset\":{\"version\":\"2.0\",\"generator\":\"vkhrt_b200\"},\"scene\":0,\"scenes\":[{\"nodes\":[0]}],\"nodes\":[{\"mesh\":0}],"
            "\"meshes\":[{\"primitives\":[{\"mode\":1,\"attributes\":{\"POSITION\":0},\"indices\":1}]}],"
            "\"accessors\":[{\"bufferView\":0,\"componentType\":5126,\"count\":%u,\"type\":\"VEC3\",\"min\":[%.9g,%.9g,%.9g],\"max\":[%.9g,%.9g,%.9g]},"
            "{\"bufferView\":1,\"componentType\":5125,\"count\":%zu,\"type\":\"SCALAR\"}],"
            "\"bufferViews\":[{\"buffer\":0,\"byteOffset\":0,\"byteLength\":%zu,\"target\":34962},{\"buffer\":0,\"byteOffset\":%zu,\"byteLength\":%zu,\"target\":34963}],"
            "\"buffers\":[{\"byteLength\":%zu}]}",
            in->n_vertices, (double)lo[0], (double)lo[1], (double)lo[2], (double)hi[0], (double)hi[1], (double)hi[2], (size_t)in->n_segments * 2,
            pos_bytes, pos_bytes, idx_bytes, pos_bytes + idx_bytes);
    else
        n = std::snprintf(js, sizeof(js), "{\"asset\":{\"version\":\"2.0\",\"generator\":\"vkhrt_b200\"},\"scenes\":[{\"nodes\":[]}]}");
    if (n <= 0 || (size_t)n >= sizeof(js)) return fail(VKHRT_ERR_IO, "glb: header formatting failed");
    std::string json(js, (size_t)n);
    while (json.size() % 4) json.push_back(' ');
    const size_t bin_bytes = (in->n_vertices && in->n_segments) ? pos_bytes + idx_bytes : 0;       // both are multiples of 4
    const size_t total = 12 + 8 + json.size() + (bin_bytes ? 8 + bin_bytes : 0);
    if (total > 0xFFFFFFFFull) return fail(VKHRT_ERR_UNSUPPORTED, "glb: asset larger than 4 GiB");
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(VKHRT_ERR_IO, std::string("cannot write ") + path);
    bool ok = true;
    auto put32 = [&](uint32_t v) { ok = ok && std::fwrite(&v, 4, 1, f) == 1; };
    ok = std::fwrite("glTF", 1, 4, f) == 4;
    put32(2u); put32((uint32_t)total);
    put32((uint32_t)json.size()); put32(0x4E4F534Au);
    ok = ok && std::fwrite(json.data(), 1, json.size(), f) == json.size();
    if (bin_bytes) {
        put32((uint32_t)bin_bytes); put32(0x004E4942u);
        ok = ok && std::fwrite(in->positions_xyz, 1, pos_bytes, f) == pos_bytes;
        ok = ok && std::fwrite(in->line_indices, 1, idx_bytes, f) == idx_bytes;
    }
    ok = (std::fclose(f) == 0) && ok;
    return ok ? VKHRT_OK : fail(VKHRT_ERR_IO, std::string("short write to ") + path);
}

}  // namespace vkhrt
