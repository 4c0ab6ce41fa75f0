// Self-attention of RNABert on the tcgen05 tensor cores (bf16 arm; a-15):
//     out[q, h] = softmax_k(q_h . k_h / sqrt(hd)) v_h   over the VALID keys of the structure
// replaces nn.MultiheadAttention between its in/out projections (rnampnn/model/functional.py:161-168) like the SIMT
// kernel res_attention_kernel (rnampnn.cu), which stays the fp32 arm.  The reference pads every structure to 4500 rows
// and masks the padded keys; only the L_b valid keys carry weight, so only key tiles below N are visited and invalid
// keys get weight exactly 0.
//
// Input: qkv [B*N, 3E] bf16 (the in-projection's bf16 side copy: q | k | v, head h at columns h*hd), E = 128.
// One CTA per (128-query tile, structure).  All heads of the Q / K / V tiles sit in shared memory in the standard
// [128 rows][64 ch] x 2 swizzle-128B layout (TMA), so head h is a K-slice of the operands:
//     S_h  = Q[:, h] K[:, h]^T      one UMMA per 16 channels, A and B K-major at byte offset (h*hd % 64)*2 of block h*hd/64
//     O_h += P_h V[:, h]            A = P_h (bf16) in TENSOR MEMORY (TS form), B = V read MN-major, N = hd columns
// Two warpgroups (4 warps each, thread = query row) own the even / odd heads, so one group's softmax runs beside the
// other's MMAs and TMEM round trips.  Softmax in two passes over the key tiles (pass 1: row maxima, pass 2: weights and
// P.V; recomputing S costs one tiny MMA) so that the O accumulators are never rescaled.  TMEM (512 columns):
// S of group g at 128 g, P of group g at 256 + 64 g, O (all heads, 128 columns) at 384.
// The MMAs of a group are issued by its first thread: no separate issuer warp, no request queue.
#include "tc_common.cuh"

namespace rmp {
namespace tc {

constexpr int kAttnThreads = 320;        // warp 0 = TMA producer, warp 1 = TMEM owner, warps 2..5 / 6..9 = groups 0 / 1
constexpr int kAttnSlots = 2;            // key-tile ring: slot = K tile (32 KB) + V tile (32 KB)
constexpr uint32_t kAttnTmemCols = 512;
constexpr uint32_t kAttnPCol = 256, kAttnOCol = 384;
constexpr int kAttnMaxHeadsPerGroup = 4;

struct AttnBarriers {
    uint64_t q_full, kv_full[kAttnSlots], kv_empty[kAttnSlots], s_done[2], pv_done[2];
};

struct AttnParams {
    const float* mask;      // [B*N]
    float* out;             // [B*N, 128] fp32
    int N, heads, hd;
};

__device__ __forceinline__ void attn_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void attn_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void attn_group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ float attn_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
res_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;                                   // [2][128][64] bf16
    uint8_t* sKV = sQ + kTileBytes;                       // [slots][K | V][2][128][64] bf16
    AttnBarriers* bars = reinterpret_cast<AttnBarriers*>(sKV + (size_t)kAttnSlots * 2 * kTileBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, qt = blockIdx.x;
    const int N = p.N, T = (N + kTileM - 1) / kTileM;     // key tiles of the structure
    const int row0 = b * N;

    if (threadIdx.x == 0) {
        mbar_init(&bars->q_full, 1);
        for (int s = 0; s < kAttnSlots; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 2); }
        for (int g = 0; g < 2; ++g) { mbar_init(&bars->s_done[g], 1); mbar_init(&bars->pv_done[g], 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, kAttnTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ===================== TMA producer: Q once, then the key tiles of pass 1 (K) and pass 2 (K and V) =====
        if (elect_one()) {
            tma_prefetch_desc(&tm_qkv);
            mbar_expect_tx(&bars->q_full, kTileBytes);
            tma_load_2d(sQ, &tm_qkv, &bars->q_full, 0, row0 + qt * kTileM);
            tma_load_2d(sQ + kKBlockBytes, &tm_qkv, &bars->q_full, 64, row0 + qt * kTileM);
            for (int t = 0; t < 2 * T; ++t) {
                const int slot = t % kAttnSlots, pass = t / T, kt = t % T;
                uint8_t* sK = sKV + (size_t)slot * 2 * kTileBytes;
                mbar_wait_relaxed(&bars->kv_empty[slot], ((t / kAttnSlots) & 1) ^ 1, 100);
                mbar_expect_tx(&bars->kv_full[slot], pass == 0 ? kTileBytes : 2 * kTileBytes);
                const int krow = row0 + kt * kTileM;
                tma_load_2d(sK, &tm_qkv, &bars->kv_full[slot], 128, krow);
                tma_load_2d(sK + kKBlockBytes, &tm_qkv, &bars->kv_full[slot], 192, krow);
                if (pass == 1) {
                    tma_load_2d(sK + kTileBytes, &tm_qkv, &bars->kv_full[slot], 256, krow);
                    tma_load_2d(sK + kTileBytes + kKBlockBytes, &tm_qkv, &bars->kv_full[slot], 320, krow);
                }
            }
        }
        __syncwarp();
    } else if (warp >= 2) {
        // ===================== softmax groups: thread = query row; group g owns heads g, g + 2, ... =====
        const int g = (warp - 2) >> 2;
        const int lane_grp = warp & 3;                       // TMEM lanes this warp may touch
        const int row = lane_grp * 32 + lane;
        const bool issuer = ((warp - 2) & 3) == 0 && lane == 0;
        const uint32_t lane_off = (uint32_t)(lane_grp * 32) << 16;
        const uint32_t tS = tmem_base + (uint32_t)(g * 128), tP = tmem_base + kAttnPCol + (uint32_t)(g * 64);
        const uint32_t tO = tmem_base + kAttnOCol;
        const int hd = p.hd;
        const int nh = (p.heads - g + 1) / 2;                // heads of this group (the entry point requires nh >= 1)
        const int q_glob = qt * kTileM + row;
        const bool q_in = q_glob < N;
        const bool q_valid = q_in && __ldg(p.mask + row0 + q_glob) != 0.0f;
        const float cs = 1.4426950408889634f / sqrtf((float)hd);      // scores in the log2 domain
        const uint32_t idesc_s = instr_desc_bf16(128, 128, 0);
        const uint32_t idesc_pv = instr_desc_bf16(128, hd, 1);
        const uint32_t q_addr = smem_u32(sQ), kv_addr = smem_u32(sKV);
        const int J = 2 * T * nh;                            // work items (key tile of a pass, head)

        float m[kAttnMaxHeadsPerGroup], l[kAttnMaxHeadsPerGroup];
#pragma unroll
        for (int i = 0; i < kAttnMaxHeadsPerGroup; ++i) { m[i] = -INFINITY; l[i] = 0.f; }

        // S of item j: one UMMA per 16 channels of the head (A = Q slice, B = K slice, both K-major)
        auto issue_s = [&](int j) {
            const int t = j / nh, i = j - t * nh, h = g + 2 * i, slot = t % kAttnSlots;
            if (j == 0) mbar_wait(&bars->q_full, 0);
            if (i == 0) mbar_wait(&bars->kv_full[slot], (uint32_t)((t / kAttnSlots) & 1));
            tc_fence_after();
            const int c0 = h * hd;
            const uint32_t off = (uint32_t)((c0 >> 6) * kKBlockBytes + (c0 & 63) * 2);
            const uint32_t k_addr = kv_addr + (uint32_t)(slot * 2 * kTileBytes);
            for (int ks = 0; ks < hd / 16; ++ks)
                umma_bf16(tS, smem_desc_sw128(q_addr + off + ks * 32, 16, 1024), smem_desc_sw128(k_addr + off + ks * 32, 16, 1024),
                          idesc_s, (uint32_t)(ks > 0));
            umma_commit(&bars->s_done[g]);
        };
        if (issuer) issue_s(0);

        uint32_t kbits[4] = {0u, 0u, 0u, 0u};
        int n_pv = 0;
        for (int j = 0; j < J; ++j) {
            const int t = j / nh, i = j - t * nh, pass = t / T, kt = t - pass * T, slot = t % kAttnSlots, h = g + 2 * i;
            if (i == 0) {                                    // validity of the tile's 128 keys, one bit per key
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int kj = kt * kTileM + c * 32 + lane;
                    kbits[c] = __ballot_sync(0xffffffffu, kj < N && __ldg(p.mask + row0 + kj) != 0.0f);
                }
            }
            mbar_wait(&bars->s_done[g], (uint32_t)(j & 1));
            tc_fence_after();
            if (pass == 0) {
                float mx = m[i];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v[32];
                    tmem_ld32(tS + lane_off + (uint32_t)(c * 32), v);
                    const uint32_t bits = kbits[c];
                    if (bits == 0xffffffffu) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) mx = fmaxf(mx, v[q]);
                    } else if (bits != 0u) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) mx = fmaxf(mx, (bits >> q) & 1u ? v[q] : -INFINITY);
                    }
                }
                m[i] = mx;
                tc_fence_before();
                attn_group_bar(g);                            // everybody has read S
                if (issuer) {
                    tc_fence_after();
                    if (i == nh - 1) mbar_arrive(&bars->kv_empty[slot]);
                    if (j + 1 < J) issue_s(j + 1);
                }
            } else {
                if (n_pv > 0) {                               // the previous P.V has consumed P
                    mbar_wait(&bars->pv_done[g], (uint32_t)((n_pv - 1) & 1));
                    tc_fence_after();
                }
                const float mc = m[i] * cs;
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v[32];
                    tmem_ld32(tS + lane_off + (uint32_t)(c * 32), v);
                    const uint32_t bits = kbits[c];
                    uint32_t o[16];
                    if (bits == 0u) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) o[q] = 0u;
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            float e = attn_ex2(fmaf(v[q], cs, -mc));
                            if (bits != 0xffffffffu) e = (bits >> q) & 1u ? e : 0.f;
                            v[q] = e;
                        }
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            o[q] = pack_bf16(v[2 * q], v[2 * q + 1]);
                            // the row sum uses the bf16-rounded weights the tensor core will see
                            sum += __uint_as_float(o[q] << 16) + __uint_as_float(o[q] & 0xffff0000u);
                        }
                    }
                    attn_tmem_st16(tP + lane_off + (uint32_t)(c * 16), o);
                }
                l[i] += sum;
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                attn_group_bar(g);                            // P complete, S free
                if (issuer) {
                    tc_fence_after();
                    const int c0 = h * hd;
                    const uint32_t v_addr = kv_addr + (uint32_t)(slot * 2 * kTileBytes + kTileBytes) +
                                            (uint32_t)((c0 >> 6) * kKBlockBytes + (c0 & 63) * 2);
#pragma unroll
                    for (int k = 0; k < 8; ++k)               // 128 keys = 8 x UMMA_K(16); V read MN-major: 16 rows per step
                        attn_umma_ts(tO + (uint32_t)c0, tP + (uint32_t)(k * 8),
                                     smem_desc_sw128(v_addr + (uint32_t)(k * 2048), kKBlockBytes, 1024), idesc_pv,
                                     (uint32_t)((kt > 0) || (k > 0)));
                    umma_commit(&bars->pv_done[g]);
                    if (i == nh - 1) umma_commit(&bars->kv_empty[slot]);
                    if (j + 1 < J) issue_s(j + 1);
                }
                ++n_pv;
            }
        }
        if (n_pv > 0) {
            mbar_wait(&bars->pv_done[g], (uint32_t)((n_pv - 1) & 1));
            tc_fence_after();
        }
        // ---- O_h / l_h -> out (padded / masked query rows -> 0, like the reference's zeroed rows).  The TMEM loads are
        //      warp-collective: every lane takes part, rows beyond N only skip the store.
        float* dst = p.out + (size_t)(row0 + (q_in ? q_glob : 0)) * 128;
        for (int i = 0; i < nh; ++i) {
            const int c0 = (g + 2 * i) * hd;
            const float inv = (q_valid && l[i] > 0.f) ? 1.f / l[i] : 0.f;
            for (int cc = 0; cc < hd; cc += 16) {
                float v[16];
                tmem_ld16(tO + lane_off + (uint32_t)(c0 + cc), v);
                if (q_in) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        reinterpret_cast<float4*>(dst + c0 + cc)[q] =
                            q_valid ? make_float4(v[4 * q] * inv, v[4 * q + 1] * inv, v[4 * q + 2] * inv, v[4 * q + 3] * inv)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kAttnTmemCols);
}

}  // namespace tc
}  // namespace rmp

extern "C" int rmp_res_attention_tc(const void* qkv_bf16, const float* mask, int B, int N, int E, int heads, float* out,
                                    void* stream) {
    using namespace rmp;
    RMP_REQUIRE(qkv_bf16 && mask && out, RMP_E_ARG, "rmp_res_attention_tc: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && heads > 0 && E == 128 && E % heads == 0, RMP_E_SHAPE,
                "rmp_res_attention_tc: E must be 128 (got %d) and divisible by heads", E);
    const int hd = E / heads;
    RMP_REQUIRE((hd == 16 || hd == 32) && heads % 2 == 0, RMP_E_SHAPE,
                "rmp_res_attention_tc: head dim %d not in {16, 32} (use rmp_res_attention)", hd);
    RMP_REQUIRE(aligned16(qkv_bf16) && aligned16(out), RMP_E_ALIGN, "rmp_res_attention_tc: 16-byte alignment");
    RMP_REQUIRE(B <= 65535, RMP_E_SHAPE, "rmp_res_attention_tc: B=%d exceeds 65535 structures per call", B);
    CUtensorMap tm;
    int rc = tc::make_map_bf16(&tm, qkv_bf16, (uint64_t)B * N, (uint64_t)3 * E, 128);
    if (rc) return rc;
    tc::AttnParams p{mask, out, N, heads, hd};
    const size_t smem = 1024 + (size_t)(1 + 2 * tc::kAttnSlots) * tc::kTileBytes + sizeof(tc::AttnBarriers) + 16;
    cudaFuncSetAttribute(tc::res_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc::res_attention_tc_kernel<<<dim3(ceil_div(N, tc::kTileM), B), tc::kAttnThreads, smem, (cudaStream_t)stream>>>(tm, p);
    return check_launch("rmp_res_attention_tc");
}
