// host.cpp — host-side helpers of the C ABI that need no GPU:
//   * FlyCamera matrices (reference source/fly_camera.cpp:25-35,72-82) and their inverses as
//     Renderer::UpdateCameraResource uploads them (source/renderer.cpp:189-195);
//   * the seeded synthetic groom generator (hair assets are not available offline).
// glm is not available here, so lookAt / perspectiveRH_ZO / inverse are restated from their
// standard definitions (column-major, m[col*4+row]).
#include "../../include/vkhrt_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct V { float x, y, z; };
inline V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dotv(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V crossv(V a, V b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V norm(V a) { float s = 1.0f / std::sqrt(dotv(a, a)); return {a.x * s, a.y * s, a.z * s}; }

// general 4x4 inverse by cofactors (what glm::inverse(mat4) computes), column-major
void inverse4(const float* m, float* out)
{
    float a00 = m[0], a01 = m[1], a02 = m[2], a03 = m[3];
    float a10 = m[4], a11 = m[5], a12 = m[6], a13 = m[7];
    float a20 = m[8], a21 = m[9], a22 = m[10], a23 = m[11];
    float a30 = m[12], a31 = m[13], a32 = m[14], a33 = m[15];
    float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10;
    float b03 = a01 * a12 - a02 * a11, b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
    float b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30, b08 = a20 * a33 - a23 * a30;
    float b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
    float det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
    float id = 1.0f / det;
    out[0] = (a11 * b11 - a12 * b10 + a13 * b09) * id;
    out[1] = (a02 * b10 - a01 * b11 - a03 * b09) * id;
    out[2] = (a31 * b05 - a32 * b04 + a33 * b03) * id;
    out[3] = (a22 * b04 - a21 * b05 - a23 * b03) * id;
    out[4] = (a12 * b08 - a10 * b11 - a13 * b07) * id;
    out[5] = (a00 * b11 - a02 * b08 + a03 * b07) * id;
    out[6] = (a32 * b02 - a30 * b05 - a33 * b01) * id;
    out[7] = (a20 * b05 - a22 * b02 + a23 * b01) * id;
    out[8] = (a10 * b10 - a11 * b08 + a13 * b06) * id;
    out[9] = (a01 * b08 - a00 * b10 - a03 * b06) * id;
    out[10] = (a30 * b04 - a31 * b02 + a33 * b00) * id;
    out[11] = (a21 * b02 - a20 * b04 - a23 * b00) * id;
    out[12] = (a11 * b07 - a10 * b09 - a12 * b06) * id;
    out[13] = (a00 * b09 - a01 * b07 + a02 * b06) * id;
    out[14] = (a31 * b01 - a30 * b03 - a32 * b00) * id;
    out[15] = (a20 * b03 - a21 * b01 + a22 * b00) * id;
}

inline uint64_t splitmix_next(uint64_t& s)
{
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline double u01(uint64_t& s) { return (double)(splitmix_next(s) >> 40) * (1.0 / 16777216.0); }

}  // namespace

extern "C" {

void vkhrt_camera_matrices(const float position[3], float yaw_deg, float pitch_deg, float fov_deg, float aspect,
                           float near_plane, float far_plane, float view_inverse_out[16], float proj_inverse_out[16])
{
    const float rad = 0.01745329251994329576923690768489f;
    // FlyCamera::UpdateCameraVectors (fly_camera.cpp:72-82)
    V front{cosf(yaw_deg * rad) * cosf(pitch_deg * rad), sinf(pitch_deg * rad), sinf(yaw_deg * rad) * cosf(pitch_deg * rad)};
    front = norm(front);
    V world_up{0.0f, 1.0f, 0.0f};
    V right = norm(crossv(front, world_up));
    V up = norm(crossv(right, front));
    // glm::lookAt(pos, pos + front, up), right-handed
    V eye{position[0], position[1], position[2]};
    V centre{eye.x + front.x, eye.y + front.y, eye.z + front.z};
    V f = norm(sub(centre, eye));
    V s = norm(crossv(f, up));
    V u = crossv(s, f);
    float view[16] = {s.x, u.x, -f.x, 0.0f, s.y, u.y, -f.y, 0.0f, s.z, u.z, -f.z, 0.0f, -dotv(s, eye), -dotv(u, eye), dotv(f, eye), 1.0f};
    // glm::perspectiveRH_ZO(radians(fov), aspect, near, far), then [1][1] *= -1 (fly_camera.cpp:30-35)
    float th = tanf(fov_deg * rad / 2.0f);
    float proj[16] = {0};
    proj[0] = 1.0f / (aspect * th);
    proj[5] = -(1.0f / th);
    proj[10] = far_plane / (near_plane - far_plane);
    proj[11] = -1.0f;
    proj[14] = -(far_plane * near_plane) / (far_plane - near_plane);
    inverse4(view, view_inverse_out);
    inverse4(proj, proj_inverse_out);
}

int vkhrt_groom_generate(uint32_t n_strands, uint32_t segs, int32_t style, uint64_t seed, float* pos, uint32_t* idx)
{
    if (!pos || !idx || segs == 0) return VKHRT_ERR_INVALID_ARGUMENT;
    if (style != VKHRT_GROOM_STRAIGHT && style != VKHRT_GROOM_CURLY) return VKHRT_ERR_INVALID_ARGUMENT;
    if ((uint64_t)n_strands * (segs + 1) >= 0xFFFFFFFFull) return VKHRT_ERR_UNSUPPORTED;
    const double two_pi = 6.283185307179586476925286766559;
    const double cx = 0.0, cy = 150.0, cz = 0.0, head_r = 8.0, length = 6.0;
    const double step = length / (double)segs, omega = two_pi / 1.5;
    auto work = [&](uint32_t j0, uint32_t j1) {
        for (uint32_t j = j0; j < j1; ++j) {
            uint64_t st = seed ^ ((uint64_t)j * 0x9E3779B97F4A7C15ull);
            double a = u01(st), b = u01(st), c = u01(st);
            double y = -0.2 + 1.2 * a, th = two_pi * b, phi = two_pi * c;
            double rxy = std::sqrt(std::max(0.0, 1.0 - y * y));
            double nx = rxy * std::cos(th), ny = y, nz = rxy * std::sin(th);
            double rx = cx + head_r * nx, ry = cy + head_r * ny, rz = cz + head_r * nz;
            // (t1, t2) = MakeOrthonormalBasis(n): v first, then u = cross(v, n)
            double vx, vy, vz;
            if (std::fabs(nx) > std::fabs(ny)) { double l = std::sqrt(nz * nz + nx * nx); vx = -nz / l; vy = 0.0; vz = nx / l; }
            else { double l = std::sqrt(nz * nz + ny * ny); vx = 0.0; vy = nz / l; vz = -ny / l; }
            double ux = vy * nz - vz * ny, uy = vz * nx - vx * nz, uz = vx * ny - vy * nx;
            uint32_t base = j * (segs + 1);
            for (uint32_t k = 0; k <= segs; ++k) {
                double s = step * (double)k;
                double px = rx + nx * s, py = ry + ny * s, pz = rz + nz * s;
                if (style == VKHRT_GROOM_CURLY) {
                    double cs = std::cos(omega * s + phi), sn = std::sin(omega * s + phi);
                    py -= 0.03 * s * s;
                    px += 0.25 * (cs * ux + sn * vx); py += 0.25 * (cs * uy + sn * vy); pz += 0.25 * (cs * uz + sn * vz);
                }
                float* o = pos + 3 * (size_t)(base + k);
                o[0] = (float)px; o[1] = (float)py; o[2] = (float)pz;
                if (k < segs) { uint32_t* e = idx + 2 * ((size_t)j * segs + k); e[0] = base + k; e[1] = base + k + 1; }
            }
        }
    };
    unsigned nt = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
    if (n_strands < 4096) nt = 1;
    std::vector<std::thread> th;
    uint32_t per = (n_strands + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        uint32_t j0 = std::min(n_strands, t * per), j1 = std::min(n_strands, j0 + per);
        if (j0 < j1) th.emplace_back(work, j0, j1);
    }
    for (auto& t : th) t.join();
    return VKHRT_OK;
}

}  // extern "C"
