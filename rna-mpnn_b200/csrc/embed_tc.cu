// Fused edge featurizer of the bf16 arm: geometry -> raw edge features -> Linear(115->128) on tcgen05
// -> reference Normalize -> bf16 h_E, one kernel, nothing but coordinates in and h_E out.
//
// Replaces (rdesign/model/feature.py) :193-206 (_orientations_coarse, six edge _get_rbf calls),
// :215-220 (feature concat), :224 (edge_embedding + norm_edges).  The reference materialises six
// [B,N,N] distance matrices and a [Ne,115] feature tensor; here a tile of 128 edges computes its 115
// features in registers, writes them as the bf16 A operand (128-byte swizzle) and multiplies by the
// resident 32 KB weight tile.  K order inside the kernel is permuted for aligned 16-byte chunks:
//     [ rbf(atom 0..5) 96 | quaternion 4 | direction 15 | zero pad 13 ]
// (the host permutes the weight columns the same way; see ops.pack_edge_embedding).
//
// 320 threads, two CTAs per SM (64 KB smem, 128 TMEM columns each): warp 1 = MMA issuer,
// warps 2..9 = feature + epilogue threads (thread = one edge row, two feature / column halves).
#include "geom.cuh"
#include "tc_common.cuh"

namespace rmp {
namespace tc {

struct EmbedParams {
    const float* X;                  // [B*N, 6, 3]
    const float* frames;             // [B*N, 9]
    const int32_t* node_src;         // flat node -> dense residue
    const int32_t* ctr;
    const int32_t* nbr;
    const float* bias;               // [128]
    const float* gain;               // [128] Normalize gain
    const float* beta;               // [128] Normalize bias
    float mu[16];                    // RBF centres
    uint16_t* h_E;                   // [n_edges,128] bf16
    int n_edges;
};
struct alignas(8) EmbedBarriers {
    uint64_t w_full, acc_full, act_ready;
};

// exp(-((d - mu) / 1.25)^2) = 2^(-t^2), t = (d - mu) * (0.8 * sqrt(log2 e)): one FMA, one multiply and one MUFU.EX2
// per basis function.  The result is rounded to bf16 right away, so the fp32 arm's IEEE division (features.cu)
// is not reproduced here.
__device__ __forceinline__ void rbf16_bf16(float d, const float (&mu)[16], uint8_t* a_buf, int row, int chunk0) {
    constexpr float kC = 0.96089792f;          // 0.8 * sqrt(log2(e))
    const float dc = d * kC;
    float v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const float t = fmaf(-mu[r], kC, dc);
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(v[r]) : "f"(-(t * t)));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
        *reinterpret_cast<uint4*>(a_buf + ((chunk0 + j) >> 3) * kKBlockBytes + sw128_chunk(row, (chunk0 + j) & 7)) =
            make_uint4(pack_bf16(v[8 * j + 0], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                       pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
}

__global__ void __launch_bounds__(kThreads, 2)
edge_embed_tc_kernel(const __grid_constant__ CUtensorMap tm_w, const EmbedParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;                                  // [2][128][64] bf16 weight, resident
    uint8_t* sA = sW + kTileBytes;                       // [2][128][64] bf16 features of one tile
    EmbedBarriers* bars = reinterpret_cast<EmbedBarriers*>(sA + kTileBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    float* s_par = reinterpret_cast<float*>(s_tmem + 2);  // bias | gain | beta  [3][128]
    float* s_red = s_par + 3 * 128;                       // [2 stats][2 halves][128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.n_edges + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        mbar_init(&bars->w_full, 1);
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->act_ready, kEpiThreads);
        fence_barrier_init();
    }
    for (int t = threadIdx.x; t < 128; t += kThreads) {
        s_par[t] = p.bias[t]; s_par[128 + t] = p.gain[t]; s_par[256 + t] = p.beta[t];
    }
    if (warp == 1) tmem_alloc(s_tmem, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 1) {
        // ===================== weight load + MMA issuer =====================
        if (elect_one()) {
            mbar_expect_tx(&bars->w_full, kTileBytes);
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(sW + kb * kKBlockBytes, &tm_w, &bars->w_full, kb * 64, 0);
            mbar_wait(&bars->w_full, 0);
            constexpr uint32_t idesc = instr_desc_bf16(128, 128, 0);
            const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
            uint32_t ph = 0;
            for (int i = 0; i < my_tiles; ++i) {
                mbar_wait(&bars->act_ready, ph); ph ^= 1;
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t off = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                    umma_bf16(tmem_base, smem_desc_sw128(a_addr + off, 16, 1024), smem_desc_sw128(w_addr + off, 16, 1024),
                              idesc, k > 0);
                }
                umma_commit(&bars->acc_full);
            }
        }
        __syncwarp();
    } else if (warp >= 2) {
        const int half = (warp - 2) >> 2;
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        float* red_sum = s_red;
        float* red_sq = s_red + 2 * kTileM;
        uint32_t ph_acc = 0;
        // dense residue ids of this row's (centre, neighbour): a chain of two dependent global loads, fetched one
        // tile ahead so that only the coordinate loads sit on a tile's critical path
        auto load_ids = [&](int i, int& si, int& sj) {
            si = sj = 0;
            const int e = (blockIdx.x + i * gridDim.x) * kTileM + row;
            if (i < my_tiles && e < p.n_edges) { si = __ldg(p.node_src + __ldg(p.ctr + e)); sj = __ldg(p.node_src + __ldg(p.nbr + e)); }
        };
        int si_nx, sj_nx;
        load_ids(0, si_nx, sj_nx);
        for (int i = 0; i < my_tiles; ++i) {
            const int e = (blockIdx.x + i * gridDim.x) * kTileM + row;
            const bool valid = e < p.n_edges;
            const int si = si_nx, sj = sj_nx;
            load_ids(i + 1, si_nx, sj_nx);
            // ---- raw features of this row's half, straight into the swizzled A operand ----
            if (valid) {
                const float* xi = p.X + (size_t)si * 18;
                const float* xj = p.X + (size_t)sj * 18;
                const V3 pj = ld3(xj);
                // The two column halves of a row split the feature work evenly (the slower one gates the tile):
                //   half 0: RBF of atoms P, O5', C5', C4' (chunks 0..7) + the frame quaternion (first half of chunk 12)
                //   half 1: RBF of atoms C3', O3' (chunks 8..11) + the five local-frame directions (chunks 12..14)
                // Normalisations use MUFU.RSQ (v * rsqrt(|v|^2), 0 for a zero vector): the results are rounded to bf16
                // right away, so the fp32 arm's IEEE sqrt + divisions (geom.cuh) are not reproduced here.
                if (half == 0) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) rbf16_bf16(eps_dist(ld3(xi + a * 3), pj), p.mu, sA, row, 2 * a);
                    float Qi[9], Qj[9];
#pragma unroll
                    for (int c = 0; c < 9; ++c) { Qi[c] = __ldg(p.frames + (size_t)si * 9 + c); Qj[c] = __ldg(p.frames + (size_t)sj * 9 + c); }
                    float R[3][3];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) R[r][c] = fmaf(Qi[6 + r], Qj[6 + c], fmaf(Qi[3 + r], Qj[3 + c], Qi[r] * Qj[c]));
                    const float Rxx = R[0][0], Ryy = R[1][1], Rzz = R[2][2];
                    const float qx = sign_of(R[2][1] - R[1][2]) * (0.5f * sqrtf(fabsf(1.f + (Rxx - Ryy - Rzz))));
                    const float qy = sign_of(R[0][2] - R[2][0]) * (0.5f * sqrtf(fabsf(1.f + (-Rxx + Ryy - Rzz))));
                    const float qz = sign_of(R[1][0] - R[0][1]) * (0.5f * sqrtf(fabsf(1.f + (-Rxx - Ryy + Rzz))));
                    const float qw = 0.5f * sqrtf(fmaxf(1.f + (Rxx + Ryy + Rzz), 0.f));
                    const float n2 = fmaf(qw, qw, fmaf(qz, qz, fmaf(qy, qy, qx * qx)));
                    const float inv = rsqrtf(fmaxf(n2, 1e-24f));               // F.normalize: / max(|q|, 1e-12)
                    *reinterpret_cast<uint2*>(sA + kKBlockBytes + sw128_chunk(row, 4)) =
                        make_uint2(pack_bf16(qx * inv, qy * inv), pack_bf16(qz * inv, qw * inv));
                } else {
#pragma unroll
                    for (int a = 4; a < 6; ++a) rbf16_bf16(eps_dist(ld3(xi + a * 3), pj), p.mu, sA, row, 2 * a);
                    float Qi[9];
#pragma unroll
                    for (int c = 0; c < 9; ++c) Qi[c] = __ldg(p.frames + (size_t)si * 9 + c);
                    float f[20];
                    const V3 c3 = ld3(xi + 12);
                    const int atoms[5] = {0, 1, 2, 3, 5};
#pragma unroll
                    for (int a = 0; a < 5; ++a) {
                        const V3 d = rotate(Qi, ld3(xj + atoms[a] * 3) - c3);
                        const float n2 = fmaf(d.z, d.z, fmaf(d.y, d.y, d.x * d.x));
                        const float inv = n2 > 0.f ? rsqrtf(n2) : 0.f;            // 0/0 -> 0 (rdesign/utils/data.py:164-171)
                        f[a * 3 + 0] = d.x * inv; f[a * 3 + 1] = d.y * inv; f[a * 3 + 2] = d.z * inv;
                    }
#pragma unroll
                    for (int c = 15; c < 20; ++c) f[c] = 0.f;
                    // directions 0..3 complete chunk 12 behind the quaternion, 4..11 -> chunk 13, 12..14 + zeros -> chunk 14
                    *reinterpret_cast<uint2*>(sA + kKBlockBytes + sw128_chunk(row, 4) + 8) =
                        make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        *reinterpret_cast<uint4*>(sA + kKBlockBytes + sw128_chunk(row, 5 + j)) =
                            make_uint4(pack_bf16(f[4 + 8 * j + 0], f[4 + 8 * j + 1]), pack_bf16(f[4 + 8 * j + 2], f[4 + 8 * j + 3]),
                                       pack_bf16(f[4 + 8 * j + 4], f[4 + 8 * j + 5]), pack_bf16(f[4 + 8 * j + 6], f[4 + 8 * j + 7]));
                    *reinterpret_cast<uint4*>(sA + kKBlockBytes + sw128_chunk(row, 7)) = make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(sA + half * kKBlockBytes + sw128_chunk(row, j)) = make_uint4(0u, 0u, 0u, 0u);
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&bars->act_ready);

            // ---- epilogue: + bias, Normalize (unbiased variance, eps inside and outside the sqrt), bf16 store ----
            mbar_wait(&bars->acc_full, ph_acc); ph_acc ^= 1;
            tc_fence_after();
            const int c0 = half * 64;
            float x[64];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float v[32];
                tmem_ld32(t_row + (uint32_t)(c0 + ch * 32), v);
#pragma unroll
                for (int q = 0; q < 32; ++q) x[ch * 32 + q] = v[q] + s_par[c0 + ch * 32 + q];
            }
            float sm = 0.f;
#pragma unroll
            for (int q = 0; q < 64; ++q) sm += x[q];
            red_sum[half * kTileM + row] = sm;
            epi_bar_sync();
            const float mean = (sm + red_sum[(half ^ 1) * kTileM + row]) * (1.f / 128.f);
            float sq = 0.f;
#pragma unroll
            for (int q = 0; q < 64; ++q) { x[q] -= mean; sq = fmaf(x[q], x[q], sq); }
            red_sq[half * kTileM + row] = sq;
            epi_bar_sync();
            const float sigma = sqrtf((sq + red_sq[(half ^ 1) * kTileM + row]) * (1.f / 127.f) + 1e-6f);
            const float inv = 1.f / (sigma + 1e-6f);
            if (valid) {
                // 32-byte stores: every instruction writes full sectors (16-byte stores left each sector half written)
                uint16_t* dst = p.h_E + (size_t)e * 128 + c0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t pk[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int c = 16 * j + 2 * q;
                        pk[q] = pack_bf16(fmaf(s_par[128 + c0 + c], x[c] * inv, s_par[256 + c0 + c]),
                                          fmaf(s_par[128 + c0 + c + 1], x[c + 1] * inv, s_par[256 + c0 + c + 1]));
                    }
                    stg256u(dst + 16 * j, pk);
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 128);
}

}  // namespace tc
}  // namespace rmp

using namespace rmp;

extern "C" int rmp_edge_embed_bf16(const float* X, const float* frames, const int32_t* node_src, const int32_t* ctr,
                                   const int32_t* nbr, int n_edges, const float* rbf_mu_host, const void* W_perm_bf16,
                                   const float* bias, const float* gain, const float* beta, void* h_E_bf16,
                                   void* stream) {
    RMP_REQUIRE(X && frames && node_src && ctr && nbr && rbf_mu_host && W_perm_bf16 && bias && gain && beta && h_E_bf16,
                RMP_E_ARG, "rmp_edge_embed_bf16: null pointer");
    RMP_REQUIRE(aligned16(W_perm_bf16) && aligned32(h_E_bf16), RMP_E_ALIGN, "rmp_edge_embed_bf16: W must be 16-byte, h_E 32-byte aligned (256-bit stores)");
    if (n_edges <= 0) return RMP_OK;
    CUtensorMap tm_w;
    int rc = tc::make_map_bf16_rows128(&tm_w, W_perm_bf16, 128);
    if (rc) return rc;
    tc::EmbedParams p{};
    p.X = X; p.frames = frames; p.node_src = node_src; p.ctr = ctr; p.nbr = nbr; p.bias = bias; p.gain = gain;
    p.beta = beta; p.h_E = reinterpret_cast<uint16_t*>(h_E_bf16); p.n_edges = n_edges;
    memcpy(p.mu, rbf_mu_host, sizeof(p.mu));
    const size_t smem = 1024 + 2 * tc::kTileBytes + sizeof(tc::EmbedBarriers) + 8 + (3 * 128 + 4 * tc::kTileM) * sizeof(float);
    cudaFuncSetAttribute(tc::edge_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(n_edges, tc::kTileM);
    const int grid = n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs;
    tc::edge_embed_tc_kernel<<<grid, tc::kThreads, smem, (cudaStream_t)stream>>>(tm_w, p);
    return check_launch("rmp_edge_embed_bf16");
}
