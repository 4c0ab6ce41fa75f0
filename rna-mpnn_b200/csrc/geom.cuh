// Geometry helpers shared by the featurizer kernels (fp32 CUDA-core math, no FMA contraction:
// the library is built with -fmad=false so products and sums round separately like the
// reference's ATen ops).  Reference semantics: rdesign/model/feature.py:52-132,
// rdesign/utils/data.py:164-171.
#pragma once
#include "common.cuh"

namespace rmp {

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 ld3(const float* p) { return {p[0], p[1], p[2]}; }

// v / |v| with 0/0 -> 0  (rdesign/utils/data.py:164-171)
__device__ __forceinline__ V3 unit_nan0(V3 v) {
    const float n = sqrtf(dot(v, v));
    V3 r = {v.x / n, v.y / n, v.z / n};
    if (isnan(r.x)) r.x = 0.f;
    if (isnan(r.y)) r.y = 0.f;
    if (isnan(r.z)) r.z = 0.f;
    return r;
}
// F.normalize: v / max(|v|, eps)
__device__ __forceinline__ V3 unit_eps(V3 v, float eps) {
    const float n = fmaxf(sqrtf(dot(v, v)), eps);
    return {v.x / n, v.y / n, v.z / n};
}
__device__ __forceinline__ float sign_of(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

// sqrt(|a-b|^2 + 1e-6)   (rdesign/model/feature.py:62,65)
__device__ __forceinline__ float eps_dist(V3 a, V3 b) {
    const V3 d = a - b;
    return sqrtf(dot(d, d) + 1e-6f);
}

struct RbfMu {
    float mu[16];
};

// 16 Gaussian bumps  exp(-((d - mu)/1.25)^2)   (rdesign/model/feature.py:52-58)
template <typename Store>
__device__ __forceinline__ void rbf16(float d, const RbfMu& c, Store&& st) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const float z = __fdiv_rn(d - c.mu[r], 1.25f);
        st(r, expf(-(z * z)));
    }
}

__device__ __forceinline__ V3 rotate(const float* q, V3 d) {
    return {q[0] * d.x + q[1] * d.y + q[2] * d.z, q[3] * d.x + q[4] * d.y + q[5] * d.z,
            q[6] * d.x + q[7] * d.y + q[8] * d.z};
}

// unit quaternion (x, y, z, w) of R = Qi^T Qj   (rdesign/model/feature.py:70-88, :124)
__device__ __forceinline__ void frame_quaternion(const float* Qi, const float* Qj, float (&q)[4]) {
    float R[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = Qi[0 + r] * Qj[0 + c] + Qi[3 + r] * Qj[3 + c] + Qi[6 + r] * Qj[6 + c];
    const float Rxx = R[0][0], Ryy = R[1][1], Rzz = R[2][2];
    const float qx = sign_of(R[2][1] - R[1][2]) * (0.5f * sqrtf(fabsf(1.f + (Rxx - Ryy - Rzz))));
    const float qy = sign_of(R[0][2] - R[2][0]) * (0.5f * sqrtf(fabsf(1.f + (-Rxx + Ryy - Rzz))));
    const float qz = sign_of(R[1][0] - R[0][1]) * (0.5f * sqrtf(fabsf(1.f + (-Rxx - Ryy + Rzz))));
    const float qw = sqrtf(fmaxf(1.f + (Rxx + Ryy + Rzz), 0.f)) / 2.f;
    const float qn = fmaxf(sqrtf(qx * qx + qy * qy + qz * qz + qw * qw), 1e-12f);
    q[0] = qx / qn; q[1] = qy / qn; q[2] = qz / qn; q[3] = qw / qn;
}

}  // namespace rmp
