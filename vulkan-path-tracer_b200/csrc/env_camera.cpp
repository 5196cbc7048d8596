// env_camera.cpp -- host-side one-off preprocessing the reference performs in PathTracer::LoadEnvironmentMap
// and the Editor/FlyCamera camera round trip; plus small shared helpers.
#include "host_api.h"
#include <cmath>
#include <cstring>
#include <numeric>

namespace b200pt {

// PathTracer/PathTracer.cpp:1137-1332.  Serial fp32, same evaluation order as the reference:
//   importance = solid angle * max(R,G,B)                  :1176-1199
//   sum = std::accumulate(.., 0.0f); normalise to mean 1   :1215-1230
//   partition + pairing                                    :1239-1283  (quirk Q1: the low-energy cursor is
//                                                          pre-incremented, so slot 0 of the partition table keeps its
//                                                          zero initialiser and the last low entry lands in the first
//                                                          high slot -- reproduced, not repaired)
//   alpha <- max(R,G,B) / sum                              :1288-1296
float build_env_alias(float *px, uint32_t width, uint32_t height, uint2 *alias) {
    const uint32_t size = width * height;
    std::vector<float> imp(size);
    float cosPrev = 1.0F;
    const float stepPhi = (float)2.0F * (float)3.14159265358979323846 / (float)width;
    const float stepTheta = (float)3.14159265358979323846 / (float)height;
    for (uint32_t y = 0; y < height; ++y) {
        const float cosNext = std::cos((float)(y + 1) * stepTheta);
        const float area = (cosPrev - cosNext) * stepPhi;
        cosPrev = cosNext;
        const float *row = px + (size_t)y * width * 4;
        for (uint32_t x = 0; x < width; ++x) imp[(size_t)y * width + x] = area * std::max(row[x * 4], std::max(row[x * 4 + 1], row[x * 4 + 2]));
    }
    const float sum = std::accumulate(imp.begin(), imp.end(), 0.0f);
    const float average = sum / float(size);
    std::vector<float> q(size);
    for (uint32_t i = 0; i < size; i++) { q[i] = (average == 0.0f) ? 0.0f : imp[i] / average; alias[i].x = i; }
    std::vector<uint32_t> table((size_t)size + 1, 0u);   // +1: the reference indexes [size] when every texel is below average
    uint32_t lowCur = 0U, highCur = size;
    for (uint32_t i = 0; i < size; ++i) {
        if (q[i] < 1.F) table[++lowCur] = i;
        else table[--highCur] = i;
    }
    for (lowCur = 0; lowCur < highCur && highCur < size; lowCur++) {
        const uint32_t lo = table[lowCur], hi = table[highCur];
        alias[lo].x = hi;
        q[hi] -= 1.F - q[lo];
        if (q[hi] < 1.0f) highCur++;
    }
    for (uint32_t i = 0; i < size; ++i) {
        memcpy(&alias[i].y, &q[i], 4);
        float *p = px + (size_t)i * 4;
        p[3] = (sum == 0.0f) ? 0.0f : std::max(p[0], std::max(p[1], p[2])) / sum;
    }
    return sum;
}

static void invert4(const float m[16], float out[16]) {   // column-major, Gauss-Jordan in double
    double w[4][8];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { w[r][c] = m[c * 4 + r]; w[r][4 + c] = r == c ? 1.0 : 0.0; }
    for (int c = 0; c < 4; c++) {
        int piv = c; for (int r = c + 1; r < 4; r++) if (std::fabs(w[r][c]) > std::fabs(w[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 8; j++) std::swap(w[c][j], w[piv][j]);
        double d = w[c][c]; if (d == 0.0) d = 1e-300;
        for (int j = 0; j < 8; j++) w[c][j] /= d;
        for (int r = 0; r < 4; r++) if (r != c) { double f = w[r][c]; for (int j = 0; j < 8; j++) w[r][j] -= f * w[c][j]; }
    }
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out[c * 4 + r] = (float)w[r][4 + c];
}

struct V3 { float x, y, z; };
static V3 nrm(V3 a) { float inv = 1.0f / std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return { a.x * inv, a.y * inv, a.z * inv }; }
static V3 crs(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }

// Editor.cpp:45-48 builds a FlyCamera from inverse(CameraViewInverse) / inverse(CameraProjectionInverse) and
// Editor.cpp:1042-1051 pushes ITS matrices back into the path tracer, so the shader sees:
//   view  = lookAt(pos, pos+front(yaw,pitch), up)        FlyCamera.cpp:84-88,96-108,110-126  (roll is lost, Q2)
//   proj  = perspective(fov', aspect', 0.1, 1000) RH_ZO  FlyCamera.cpp:90-94 with fov'/aspect' re-derived (:128-137)
//           from perspective(45deg, aspect, 0.1, 100)     PathTracer.cpp:578                      (yfov ignored, Q2)
void camera_from_view(const float view[16], float aspect, float viewInv[16], float projInv[16]) {
    float iv[16]; invert4(view, iv);
    const V3 pos = { iv[12], iv[13], iv[14] };
    const V3 fwd = nrm({ -view[0 * 4 + 2], -view[1 * 4 + 2], -view[2 * 4 + 2] });
    const float RAD2DEG = 57.295779513082320876798154814105f, DEG2RAD = 0.01745329251994329576923690768489f;
    const float yaw = std::atan2(fwd.z, fwd.x) * RAD2DEG, pitch = std::asin(fwd.y) * RAD2DEG;
    const float tanHalf0 = std::tan((45.0f * DEG2RAD) / 2.0f);
    const float P00 = 1.0f / (aspect * tanHalf0), P11 = 1.0f / tanHalf0;
    const float fov = (2.0f * std::atan(1.0f / P11)) * RAD2DEG, asp = P11 / P00;
    const V3 front = nrm({ std::cos(yaw * DEG2RAD) * std::cos(pitch * DEG2RAD), std::sin(pitch * DEG2RAD), std::sin(yaw * DEG2RAD) * std::cos(pitch * DEG2RAD) });
    const V3 right = nrm(crs(front, { 0.0f, 1.0f, 0.0f }));
    const V3 up = nrm(crs(right, front));
    const V3 center = { pos.x + front.x, pos.y + front.y, pos.z + front.z };
    const V3 f = nrm({ center.x - pos.x, center.y - pos.y, center.z - pos.z });
    const V3 s = nrm(crs(f, up));
    const V3 u = crs(s, f);
    const float vi[16] = { s.x, s.y, s.z, 0.0f, u.x, u.y, u.z, 0.0f, -f.x, -f.y, -f.z, 0.0f, pos.x, pos.y, pos.z, 1.0f };
    memcpy(viewInv, vi, sizeof vi);
    const float tanHalf = std::tan((fov * DEG2RAD) / 2.0f);
    const float zn = 0.1f, zf = 1000.0f;
    const float A = 1.0f / (asp * tanHalf), B = 1.0f / tanHalf, C = zf / (zn - zf), D = -(zf * zn) / (zf - zn);
    const float pi_[16] = { 1.0f / A, 0, 0, 0, 0, 1.0f / B, 0, 0, 0, 0, 0, 1.0f / D, 0, 0, -1.0f, C / D };
    memcpy(projInv, pi_, sizeof pi_);
}

// WorldToObject linear part (the driver supplies it to the hit shader; SH/Surface.slang:49,60): inverse in double, rounded once
void mat3_inverse_from_o2w(const float o[12], float w[9]) {
    double m[3][3]; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) m[r][c] = o[r * 4 + c];
    const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    const double id = 1.0 / det;
    w[0] = (float)((m[1][1] * m[2][2] - m[1][2] * m[2][1]) * id); w[1] = (float)((m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id); w[2] = (float)((m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id);
    w[3] = (float)((m[1][2] * m[2][0] - m[1][0] * m[2][2]) * id); w[4] = (float)((m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id); w[5] = (float)((m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id);
    w[6] = (float)((m[1][0] * m[2][1] - m[1][1] * m[2][0]) * id); w[7] = (float)((m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id); w[8] = (float)((m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id);
}

// PathTracer/PostProcessor.cpp:128-158
uint32_t bloom_mip_sizes(uint32_t W, uint32_t H, uint32_t wh[20]) {
    uint32_t w = W, h = H, n = 0;
    for (uint32_t i = 0; i < 10; i++) {
        wh[2 * n] = w; wh[2 * n + 1] = h; n++;
        if (w % 2 != 0) w -= 1;
        if (h % 2 != 0) h -= 1;
        w /= 2; h /= 2;
        if (w < 2 || h < 2) break;
    }
    return n;
}

uint32_t partition_global_row(uint32_t local_row, uint32_t rank, uint32_t world, uint32_t band) {
    if (!world) world = 1;
    if (!band) band = 1;
    return ((local_row / band) * world + rank) * band + local_row % band;
}
uint32_t partition_local_rows(uint32_t H, uint32_t rank, uint32_t world, uint32_t band) {
    if (!world) world = 1;
    if (!band) band = 1;
    uint32_t n = 0;
    for (uint32_t y0 = rank * band; y0 < H; y0 += world * band) n += std::min(band, H - y0);
    return n;
}

} // namespace b200pt
