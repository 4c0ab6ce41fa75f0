// Raw geometric features of the rdesign featurizer (fp32 CUDA-core math).
//
// Reference semantics: rdesign/model/feature.py:52-154 (_rbf, _get_rbf,
// _quaternions, _orientations_coarse, _dihedrals) and :179-220 (feature order).
// Nothing [B,N,N]-shaped is ever materialised: every node / edge recomputes its
// few distances from the 6x3 coordinates of its one or two residues.
#include "geom.cuh"

namespace rmp {

// --------------------------------------------------------------------------------
// local frames: one thread per dense residue  (feature.py:90-107)
// --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frames_kernel(const float* __restrict__ X, int B, int N,
                                                    float* __restrict__ frames) {
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (size_t)B * N) return;
    const int i = (int)(idx % N);
    float* q = frames + idx * 9;
    if (i == N - 1) {
#pragma unroll
        for (int c = 0; c < 9; ++c) q[c] = 0.f;
        return;
    }
    const float* xi = X + idx * 18;
    const V3 c3 = ld3(xi + 12), o3 = ld3(xi + 15), p1 = ld3(xi + 18);   // C3'_i, O3'_i, P_{i+1}
    const V3 u0 = unit_nan0(o3 - c3), u1 = unit_nan0(p1 - o3);
    const V3 n0 = unit_nan0(cross(u0, u1));
    const V3 b1 = unit_nan0(u0 - u1);
    const V3 t = cross(b1, n0);
    q[0] = b1.x; q[1] = b1.y; q[2] = b1.z;
    q[3] = n0.x; q[4] = n0.y; q[5] = n0.z;
    q[6] = t.x;  q[7] = t.y;  q[8] = t.z;
}

// --------------------------------------------------------------------------------
// node features: 64 nodes per block, staged in smem, written coalesced
// --------------------------------------------------------------------------------
constexpr int kNodeTile = 64;

// pseudo-dihedral of flat atom slot t (feature.py:134-154): stride-5 differences over
// the flattened [6N,3] atom list of structure b, zero padded residues included.
__device__ __forceinline__ float dihedral_angle(const float* __restrict__ Xb, int N, int t) {
    if (t < 3 || t >= 6 * N - 4) return 0.f;
    const int s = t - 3;
    V3 u[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) u[q] = unit_eps(ld3(Xb + (size_t)(s + q + 5) * 3) - ld3(Xb + (size_t)(s + q) * 3), 1e-12f);
    const V3 na = unit_eps(cross(u[0], u[1]), 1e-12f);
    const V3 nb = unit_eps(cross(u[1], u[2]), 1e-12f);
    const float lo = (float)(-1.0 + 1e-7), hi = (float)(1.0 - 1e-7);
    const float c = fminf(fmaxf(dot(na, nb), lo), hi);
    return sign_of(dot(u[0], nb)) * acosf(c);
}

// Four threads per node, each computing a quarter of the 101 features (same per-feature arithmetic as one thread per
// node, so the values are unchanged): with one thread per node the whole batch was 110 threads per SM of long serial
// work (80 expf, 6 acosf per thread) and the kernel took as long as a message-passing layer's node update.
constexpr int kNodeParts = 4;

__global__ void __launch_bounds__(kNodeTile * kNodeParts) node_raw_kernel(const float* __restrict__ X, int N,
                                                            const int32_t* __restrict__ node_src,
                                                            const int32_t* __restrict__ counts, int nv_cap,
                                                            RbfMu mu, const float* __restrict__ frames,
                                                            float* __restrict__ node_raw) {
    __shared__ float tile[kNodeTile * RMP_NODE_RAW];
    const int nv = min(counts[0], nv_cap);
    const int v0 = blockIdx.x * kNodeTile;
    if (v0 >= nv) return;
    const int t = threadIdx.x % kNodeTile, part = threadIdx.x / kNodeTile, v = v0 + t;
    if (v < nv) {
        float* f = tile + t * RMP_NODE_RAW;
        const int src = node_src[v];
        const int b = src / N, i = src - b * N;
        const float* Xb = X + (size_t)b * N * 18;
        const float* xi = Xb + (size_t)i * 18;
        if (part < 2) {
            // angle: cos | sin of the six slots (three per thread)
#pragma unroll
            for (int a3 = 0; a3 < 3; ++a3) {
                const int a = part * 3 + a3;
                const float ang = dihedral_angle(Xb, N, 6 * i + a);
                f[a] = cosf(ang);
                f[6 + a] = sinf(ang);
            }
        } else {
            // distance: atoms O5',C5',C4' (part 2) / C3',O3' (part 3) to own P
            const V3 p = ld3(xi);
            const int a_lo = part == 2 ? 1 : 4, a_hi = part == 2 ? 4 : 6;
            for (int a = a_lo; a < a_hi; ++a) {
                float* dst = f + 12 + (a - 1) * 16;
                rbf16(eps_dist(ld3(xi + a * 3), p), mu, [&](int r, float val) { dst[r] = val; });
            }
            if (part == 3) {
                // direction: frame_i * (P, C5', C4' - C3')
                const float* q = frames + (size_t)src * 9;
                const V3 c3 = ld3(xi + 12);
                const int atoms[3] = {0, 2, 3};
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const V3 d = unit_nan0(rotate(q, ld3(xi + atoms[a] * 3) - c3));
                    f[92 + a * 3 + 0] = d.x; f[92 + a * 3 + 1] = d.y; f[92 + a * 3 + 2] = d.z;
                }
            }
        }
    }
    __syncthreads();
    const int rows = min(kNodeTile, nv - v0);
    float* out = node_raw + (size_t)v0 * RMP_NODE_RAW;
    for (int o = threadIdx.x; o < rows * RMP_NODE_RAW; o += kNodeTile * kNodeParts) out[o] = tile[o];
}

// --------------------------------------------------------------------------------
// edge features: 128 edges per block (one thread each), staged in smem (58.9 KB),
// written as one contiguous coalesced span.
// --------------------------------------------------------------------------------
constexpr int kEdgeTile = 128;

template <int STRIDE>   // floats per output row: 115 (packed) or 128 (columns 115.. zero: the tensor-core arms' K)
__global__ void __launch_bounds__(kEdgeTile) edge_raw_kernel(const float* __restrict__ X, int N,
                                                            const int32_t* __restrict__ node_src,
                                                            const int32_t* __restrict__ ctr,
                                                            const int32_t* __restrict__ nbr,
                                                            const int32_t* __restrict__ counts, int ne_cap,
                                                            RbfMu mu, const float* __restrict__ frames,
                                                            float* __restrict__ edge_raw) {
    extern __shared__ float etile[];                       // [kEdgeTile][STRIDE]
    const int ne = min(counts[1], ne_cap);
    const int e0 = blockIdx.x * kEdgeTile;
    if (e0 >= ne) return;
    const int t = threadIdx.x, e = e0 + t;
    if (e < ne) {
        float* f = etile + t * STRIDE;
#pragma unroll
        for (int c = RMP_EDGE_RAW; c < STRIDE; ++c) f[c] = 0.f;
        const int si = node_src[ctr[e]], sj = node_src[nbr[e]];
        const float* xi = X + (size_t)si * 18;
        const float* xj = X + (size_t)sj * 18;
        const float* qi = frames + (size_t)si * 9;
        const float* qj = frames + (size_t)sj * 9;
        float Qi[9], Qj[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) { Qi[c] = qi[c]; Qj[c] = qj[c]; }
        // orientation quaternion of R = Qi^T Qj   (feature.py:70-88, :124)
        float quat[4];
        frame_quaternion(Qi, Qj, quat);
        f[0] = quat[0]; f[1] = quat[1]; f[2] = quat[2]; f[3] = quat[3];
        // distance: atom a of the CENTRE to P of the NEIGHBOUR   (feature.py:200-206, :62-63)
        const V3 pj = ld3(xj);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            float* dst = f + 4 + a * 16;
            rbf16(eps_dist(ld3(xi + a * 3), pj), mu, [&](int r, float val) { dst[r] = val; });
        }
        // direction: frame_i * (neighbour P,O5',C5',C4',O3' - centre C3')   (feature.py:116-122)
        const V3 c3 = ld3(xi + 12);
        const int atoms[5] = {0, 1, 2, 3, 5};
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            const V3 d = unit_nan0(rotate(Qi, ld3(xj + atoms[a] * 3) - c3));
            f[100 + a * 3 + 0] = d.x; f[100 + a * 3 + 1] = d.y; f[100 + a * 3 + 2] = d.z;
        }
    }
    __syncthreads();
    const int rows = min(kEdgeTile, ne - e0);
    float* out = edge_raw + (size_t)e0 * STRIDE;
    for (int o = t; o < rows * STRIDE; o += kEdgeTile) out[o] = etile[o];
}

}  // namespace rmp

using namespace rmp;

static int raw_features_impl(const float* X, int B, int N, const int32_t* node_src, const int32_t* ctr,
                             const int32_t* nbr, const int32_t* counts, int nv_cap, int ne_cap,
                             const float* rbf_mu_host, float* frames, float* node_raw, float* edge_raw, bool padded,
                             void* stream) {
    RMP_REQUIRE(X && node_src && ctr && nbr && counts && rbf_mu_host && frames, RMP_E_ARG,
                "rmp_rdesign_raw_features: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && nv_cap >= 0 && ne_cap >= 0, RMP_E_ARG, "rmp_rdesign_raw_features: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    RbfMu mu;
    memcpy(mu.mu, rbf_mu_host, sizeof(mu.mu));
    const size_t total = (size_t)B * N;
    frames_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, B, N, frames);
    if (node_raw && nv_cap > 0)
        node_raw_kernel<<<ceil_div(nv_cap, kNodeTile), kNodeTile * kNodeParts, 0, st>>>(X, N, node_src, counts, nv_cap, mu, frames,
                                                                           node_raw);
    if (edge_raw && ne_cap > 0) {
        const int smem = kEdgeTile * (padded ? 128 : RMP_EDGE_RAW) * (int)sizeof(float);
        auto kern = padded ? edge_raw_kernel<128> : edge_raw_kernel<RMP_EDGE_RAW>;
        // per device and not cached: the opt-in is a per-device function attribute
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        kern<<<ceil_div(ne_cap, kEdgeTile), kEdgeTile, smem, st>>>(X, N, node_src, ctr, nbr, counts, ne_cap, mu, frames, edge_raw);
    }
    return check_launch("rmp_rdesign_raw_features", 1 + (node_raw && nv_cap > 0) + (edge_raw && ne_cap > 0));
}

extern "C" int rmp_rdesign_raw_features(const float* X, int B, int N, const int32_t* node_src, const int32_t* ctr,
                                        const int32_t* nbr, const int32_t* counts, int nv_cap, int ne_cap,
                                        const float* rbf_mu_host, float* frames, float* node_raw, float* edge_raw,
                                        void* stream) {
    return raw_features_impl(X, B, N, node_src, ctr, nbr, counts, nv_cap, ne_cap, rbf_mu_host, frames, node_raw, edge_raw,
                             false, stream);
}

extern "C" int rmp_rdesign_raw_features_k128(const float* X, int B, int N, const int32_t* node_src, const int32_t* ctr,
                                             const int32_t* nbr, const int32_t* counts, int nv_cap, int ne_cap,
                                             const float* rbf_mu_host, float* frames, float* node_raw, float* edge_raw128,
                                             void* stream) {
    RMP_REQUIRE(!edge_raw128 || aligned16(edge_raw128), RMP_E_ALIGN, "rmp_rdesign_raw_features_k128: edge rows not 16-byte aligned");
    return raw_features_impl(X, B, N, node_src, ctr, nbr, counts, nv_cap, ne_cap, rbf_mu_host, frames, node_raw, edge_raw128,
                             true, stream);
}
