// Generic fused row MLP on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM): the bf16 arm of
// the secondary (rnampnn) path's Linear stacks -- ResMPNN message / edge MLPs (rnampnn/model/mpnn.py:154-265),
// the RNABert / RawFFN feed-forward stacks and attention projections (rnampnn/model/functional.py:105-127,
// :161-172, :179-202), the edge embedding and the readout hidden layer.
//
// One CTA owns 128-row tiles (persistent over tiles).  Per tile:
//   prologue : the 8 epilogue warps gather the input rows (up to 3 concatenated fp32 segments, each optionally
//              indexed), convert to bf16 and write them as the 128-byte-swizzled K-major A operand in smem;
//   stage s  : one elected thread issues tcgen05.mma 128 x 128 x 16 over K_s/64 k-blocks and N_s/128 column chunks,
//              the weight slabs [128 rows][64 k] arrive through a 6-slot TMA ring; accumulators [128][N_s] fp32 in TMEM;
//   epilogue : tcgen05.ld -> bias -> GELU (tanh form, packed bf16x2) -> next stage's A operand, in place in smem;
//              the last stage instead applies STORE (x row_scale) / SEGSUM (red.global of y * row_scale into
//              out[seg_id[r]]) / RESID_ADD (out = resid + y) straight from TMEM to HBM.
// Activations never leave the SM between stages.  Stage chain is serial per tile (one accumulator set).
#include "tc_common.cuh"

namespace rmp {
namespace tc {

constexpr int kRowMaxStages = RMP_MAX_STAGES;
constexpr int kRowRing = 6;                       // weight ring slots of 16 KB ([128 rows][64 k] bf16)
constexpr int kRowSlotBytes = 128 * 128;

// operand formats of the two instantiations: bf16 (range of fp32, 8 significant bits) or fp16 (11 significant bits,
// conversions saturate to +-65504 so that no inf / NaN is ever produced from finite inputs)
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) {
    if (!F16) return pack_bf16(lo, hi);
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool F16>
__device__ __forceinline__ void store_chunk16(uint8_t* a_buf, int row, int col0, const float (&v)[32], bool gelu) {
    if (!F16) {
        if (gelu) store_gelu_chunk(a_buf, row, col0, v);
        else store_act_chunk(a_buf, row, col0, v);
        return;
    }
    uint8_t* kb = a_buf + (col0 >> 6) * kKBlockBytes;
    const int j0 = (col0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float y[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) y[q] = gelu ? gelu_tanh_fast(v[8 * j + q]) : v[8 * j + q];
        *reinterpret_cast<uint4*>(kb + sw128_chunk(row, j0 + j)) =
            make_uint4(pack16<true>(y[0], y[1]), pack16<true>(y[2], y[3]), pack16<true>(y[4], y[5]), pack16<true>(y[6], y[7]));
    }
}

struct RowStage {
    int K, N, gelu;
    const float* bias;
};
struct RowParams {
    const float* seg_src[RMP_MAX_SEGS];
    const int32_t* seg_idx[RMP_MAX_SEGS];
    int seg_width[RMP_MAX_SEGS];
    int num_segs;
    RowStage st[kRowMaxStages];
    int S;
    int epilogue;
    const float* resid;
    const int32_t* seg_id;
    const float* row_scale;
    float alpha;
    float* out;
    int n_rows;
    int a_kblocks;                                // K-blocks of the activation buffer (max K over stages / 64)
    // the input as 128-wide pieces (a segment of width w contributes w/128 of them); used when there are <= 3
    long long* timeline;                          // optional clock64 trace of CTA 0's first epilogue thread (tuning aid)
    int a_pieces;
    const float* piece_src[3];
    const int32_t* piece_idx[3];
    int piece_stride[3];
    int all_bf16;                                 // every piece is a bf16 side copy: 256-bit gathers, no conversion
    uint16_t* out16;                              // optional bf16 copy of the STORE / RESID_ADD result
    uint32_t tmem_cols;
};
struct alignas(8) RowBarriers {
    uint64_t w_full[kRowRing], w_empty[kRowRing], acc_full, act_ready;
};

template <bool F16>
__global__ void __launch_bounds__(kThreads, 1)
row_mlp_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                  const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3,
                  const __grid_constant__ CUtensorMap tm4, const __grid_constant__ CUtensorMap tm5, const RowParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                                          // [a_kblocks][128][64] bf16 activation (in place)
    uint8_t* sW = sA + (size_t)p.a_kblocks * kKBlockBytes;       // weight ring
    RowBarriers* bars = reinterpret_cast<RowBarriers*>(sW + kRowRing * kRowSlotBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.n_rows + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.S;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRowRing; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->act_ready, kEpiThreads);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const CUtensorMap* maps[kRowMaxStages] = {&tm0, &tm1, &tm2, &tm3, &tm4, &tm5};

    if (warp == 0) {
        // ===================== TMA producer: weight slabs [128 rows][64 k] through the ring =====================
        if (elect_one()) {
            for (int s = 0; s < S; ++s) tma_prefetch_desc(maps[s]);
            uint32_t n = 0;
            for (int i = 0; i < my_tiles; ++i)
                for (int s = 0; s < S; ++s)
                    for (int kb = 0; kb < p.st[s].K / 64; ++kb)
                        for (int c = 0; c < p.st[s].N / 128; ++c, ++n) {
                            const uint32_t slot = n % kRowRing;
                            mbar_wait(&bars->w_empty[slot], ((n / kRowRing) & 1) ^ 1);
                            mbar_expect_tx(&bars->w_full[slot], (uint32_t)kRowSlotBytes);
                            tma_load_2d(sW + slot * kRowSlotBytes, maps[s], &bars->w_full[slot], kb * 64, c * 128);
                        }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            uint32_t n = 0, ph_act = 0;
            const uint32_t a_addr = smem_u32(sA);
            const uint32_t idesc = F16 ? instr_desc_f16(128, 128) : instr_desc_bf16(128, 128, 0);
            for (int i = 0; i < my_tiles; ++i)
                for (int s = 0; s < S; ++s) {
                    mbar_wait(&bars->act_ready, ph_act); ph_act ^= 1;
                    tc_fence_after();
                    for (int kb = 0; kb < p.st[s].K / 64; ++kb)
                        for (int c = 0; c < p.st[s].N / 128; ++c, ++n) {
                            const uint32_t slot = n % kRowRing;
                            mbar_wait(&bars->w_full[slot], (n / kRowRing) & 1);
                            tc_fence_after();
                            const uint32_t w_addr = smem_u32(sW + slot * kRowSlotBytes);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(tmem_base + (uint32_t)(c * 128),
                                          smem_desc_sw128(a_addr + kb * kKBlockBytes + k * 32, 16, 1024),
                                          smem_desc_sw128(w_addr + k * 32, 16, 1024), idesc, (kb | k) > 0);
                            umma_commit(&bars->w_empty[slot]);
                        }
                    umma_commit(&bars->acc_full);
                }
        }
        __syncwarp();
    } else {
        // ===================== prologue / epilogue warps =====================
        const int half = (warp - 2) >> 2;
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        const int ew = warp - 2;                                   // 8 warps x 16 rows in the prologue
        uint32_t ph_acc = 0;
        int tl_n = 0;
        const bool tl_on = p.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
        auto stamp = [&](int tag) {
            if (tl_on && tl_n < 200) { p.timeline[tl_n * 2] = tag; p.timeline[tl_n * 2 + 1] = clock64(); ++tl_n; }
        };
        stamp(1);

        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int r = tile * kTileM + row;
            const bool valid = r < p.n_rows;
            // ---- prologue: gather + concat + bf16 -> swizzled A operand.  One warp per row, a lane holds 4 consecutive
            //      channels of every 128-wide piece; the loads of 8 rows are issued before the first use ----
            if (p.all_bf16) {
                // bf16 side copies (h16 / e16): a 256-byte row is 8 lanes x 32 bytes, one LDG.256 instruction brings 4
                // rows; the 6 loads of 8 rows x 3 pieces are all in flight before the first use (two such batches)
                const int sub = lane >> 3, part = lane & 7;           // row inside the group of 4, 32-byte part of the row
                for (int g0 = 0; g0 < 4; g0 += 2) {
                uint32_t xr[3][2][8];
#pragma unroll
                for (int pc = 0; pc < 3; ++pc) {
                    if (pc < p.a_pieces) {
                        const uint16_t* src = reinterpret_cast<const uint16_t*>(p.piece_src[pc]);
                        const int32_t* idx = p.piece_idx[pc];
#pragma unroll
                        for (int gq = 0; gq < 2; ++gq) {
                            const int rg = tile * kTileM + ew * 16 + (g0 + gq) * 4 + sub;
#pragma unroll
                            for (int q = 0; q < 8; ++q) xr[pc][gq][q] = 0u;
                            if (rg < p.n_rows)
                                ldg256(src + (size_t)(idx ? __ldg(idx + rg) : rg) * 128 + part * 16, xr[pc][gq]);
                        }
                    }
                }
#pragma unroll
                for (int pc = 0; pc < 3; ++pc) {
                    if (pc < p.a_pieces) {
#pragma unroll
                        for (int gq = 0; gq < 2; ++gq) {
                            const int rl = ew * 16 + (g0 + gq) * 4 + sub;
                            uint8_t* kb = sA + (2 * pc + (part >> 2)) * kKBlockBytes;
                            const int j = (2 * part) & 7;
                            *reinterpret_cast<uint4*>(kb + sw128_chunk(rl, j)) =
                                make_uint4(xr[pc][gq][0], xr[pc][gq][1], xr[pc][gq][2], xr[pc][gq][3]);
                            *reinterpret_cast<uint4*>(kb + sw128_chunk(rl, j + 1)) =
                                make_uint4(xr[pc][gq][4], xr[pc][gq][5], xr[pc][gq][6], xr[pc][gq][7]);
                        }
                    }
                }
                }
            } else if (p.a_pieces <= 3) {
                // up to three 128-wide pieces (the gathered [h_i | h_j | e] input): all 12 row loads of a 4-row batch
                // are in flight before the first use, so a batch costs one L2 round trip instead of one per piece
                for (int r0 = 0; r0 < 16; r0 += 4) {
                    float4 xs[3][4];
#pragma unroll
                    for (int pc = 0; pc < 3; ++pc) {
                        if (pc < p.a_pieces) {
                            const float* src = p.piece_src[pc];
                            const int32_t* idx = p.piece_idx[pc];
                            const int wdt = p.piece_stride[pc];
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) {
                                const int rg = tile * kTileM + ew * 16 + r0 + rr;
                                xs[pc][rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (rg < p.n_rows)
                                    xs[pc][rr] = __ldg(reinterpret_cast<const float4*>(
                                                           src + (size_t)(idx ? __ldg(idx + rg) : rg) * wdt) + lane);
                            }
                        }
                    }
#pragma unroll
                    for (int pc = 0; pc < 3; ++pc) {
                        if (pc < p.a_pieces) {
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) {
                                const int rl = ew * 16 + r0 + rr;
                                *reinterpret_cast<uint2*>(sA + (2 * pc + (lane >> 4)) * kKBlockBytes +
                                                          sw128_chunk(rl, (lane & 15) >> 1) + (lane & 1) * 8) =
                                    make_uint2(pack16<F16>(xs[pc][rr].x, xs[pc][rr].y), pack16<F16>(xs[pc][rr].z, xs[pc][rr].w));
                            }
                        }
                    }
                }
            } else {
            for (int r0 = 0; r0 < 16; r0 += 8) {
                int kb0 = 0;
                for (int sg = 0; sg < p.num_segs; ++sg) {
                    const int wdt = p.seg_width[sg];
                    const float* src = p.seg_src[sg];
                    const int32_t* idx = p.seg_idx[sg];
                    const float* rp[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int rg = tile * kTileM + ew * 16 + r0 + rr;
                        rp[rr] = nullptr;
                        if (rg < p.n_rows) rp[rr] = src + (size_t)(idx ? __ldg(idx + rg) : rg) * wdt;
                    }
                    for (int pc = 0; pc < wdt; pc += 128) {
                        float4 xs[8];
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            xs[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (rp[rr]) xs[rr] = __ldg(reinterpret_cast<const float4*>(rp[rr] + pc) + lane);
                        }
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            const int rl = ew * 16 + r0 + rr;
                            *reinterpret_cast<uint2*>(sA + (kb0 + (pc >> 6) + (lane >> 4)) * kKBlockBytes +
                                                      sw128_chunk(rl, (lane & 15) >> 1) + (lane & 1) * 8) =
                                make_uint2(pack16<F16>(xs[rr].x, xs[rr].y), pack16<F16>(xs[rr].z, xs[rr].w));
                        }
                    }
                    kb0 += wdt >> 6;
                }
            }
            }
            stamp(2);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&bars->act_ready);

            // ---- stage epilogues ----
            for (int s = 0; s < S; ++s) {
                const RowStage stg = p.st[s];
                const int ncols = stg.N >> 1;                      // columns of this half
                const int cb = half * ncols;
                const bool last = s + 1 == S;
                float rs = 1.f;
                int dst_row = r;
                if (last && valid) {
                    if (p.row_scale) rs = __ldg(p.row_scale + r);
                    if (p.epilogue == RMP_EPI_SEGSUM) { dst_row = __ldg(p.seg_id + r); rs *= p.alpha; }
                }
                mbar_wait(&bars->acc_full, ph_acc); ph_acc ^= 1;
                tc_fence_after();
                stamp(10 + s);
                for (int ch = 0; ch < ncols / 32; ++ch) {
                    const int col0 = cb + ch * 32;
                    float v[32];
                    tmem_ld32(t_row + (uint32_t)col0, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(stg.bias + col0) + q);
                        v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
                    }
                    if (!last) {
                        store_chunk16<F16>(sA, row, col0, v, stg.gelu != 0);
                        continue;
                    }
                    if (stg.gelu) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = gelu_tanh_fast(v[q]);
                    }
                    if (!valid) continue;
                    if (p.epilogue == RMP_EPI_SEGSUM) {
                        if (rs != 0.f) {
                            float* dst = p.out + (size_t)dst_row * stg.N + col0;
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                red_add_v4(dst + 4 * q, v[4 * q] * rs, v[4 * q + 1] * rs, v[4 * q + 2] * rs, v[4 * q + 3] * rs);
                        }
                    } else {
                        float* dst = p.out + (size_t)r * stg.N + col0;
                        if (p.epilogue == RMP_EPI_RESID_ADD) {
                            const float* rsd = p.resid + (size_t)r * stg.N + col0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float t[8];
                                ldg256f(rsd + 8 * q, t);
#pragma unroll
                                for (int k = 0; k < 8; ++k) v[8 * q + k] += t[k];
                            }
                        } else if (p.row_scale) {                  // a zero scale writes exact zeros whatever the row held
#pragma unroll
                            for (int q = 0; q < 32; ++q) v[q] = rs == 0.f ? 0.f : v[q] * rs;
                        }
                        if (p.out) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) stg256f(dst + 8 * q, &v[8 * q]);
                        }
                        if (p.out16) {
                            uint32_t pk[16];
#pragma unroll
                            for (int q = 0; q < 16; ++q) pk[q] = pack_bf16(v[2 * q], v[2 * q + 1]);
                            stg256u(p.out16 + (size_t)r * stg.N + col0, &pk[0]);
                            stg256u(p.out16 + (size_t)r * stg.N + col0 + 16, &pk[8]);
                        }
                    }
                }
                tc_fence_before();
                stamp(20 + s);
                if (!last) {
                    fence_proxy_async();
                    mbar_arrive(&bars->act_ready);
                }
            }
        }
        stamp(3);
        if (tl_on) p.timeline[499] = tl_n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace tc
}  // namespace rmp

using namespace rmp;

static void* g_row_tc_timeline = nullptr;
extern "C" void rmp_set_row_tc_timeline(void* buf) { g_row_tc_timeline = buf; }

extern "C" int rmp_row_mlp_tc(const rmp_mlp_desc* dp, const void* const* W_bf16, int operand_format, void* stream) {
    RMP_REQUIRE(dp && W_bf16, RMP_E_ARG, "rmp_row_mlp_bf16: null descriptor");
    RMP_REQUIRE(operand_format == RMP_OPERAND_BF16 || operand_format == RMP_OPERAND_FP16, RMP_E_ARG,
                "rmp_row_mlp_tc: operand_format %d", operand_format);
    const rmp_mlp_desc& d = *dp;
    RMP_REQUIRE(d.num_segs >= 1 && d.num_segs <= RMP_MAX_SEGS, RMP_E_ARG, "rmp_row_mlp_bf16: num_segs=%d", d.num_segs);
    RMP_REQUIRE(d.num_stages >= 1 && d.num_stages <= RMP_MAX_STAGES, RMP_E_ARG, "rmp_row_mlp_bf16: num_stages=%d",
                d.num_stages);
    RMP_REQUIRE((d.out || (d.out_bf16 && d.epilogue == RMP_EPI_STORE)) && d.rows_cap >= 0 && !d.rows_dev, RMP_E_ARG,
                "rmp_row_mlp_bf16: bad output / rows (host row count only; out may be NULL only for STORE with out_bf16)");
    RMP_REQUIRE(d.epilogue == RMP_EPI_STORE || d.epilogue == RMP_EPI_SEGSUM || d.epilogue == RMP_EPI_RESID_ADD, RMP_E_SHAPE,
                "rmp_row_mlp_bf16: epilogue %d not supported on the tensor-core arm", d.epilogue);
    if (d.rows_cap == 0) return RMP_OK;
    tc::RowParams p;
    memset(&p, 0, sizeof(p));
    int k0 = 0;
    int n_bf16 = 0;
    for (int s = 0; s < d.num_segs; ++s) {
        n_bf16 += d.seg_bf16[s] != 0;
        RMP_REQUIRE(!d.seg_bf16[s] || d.seg_width[s] == 128, RMP_E_SHAPE, "rmp_row_mlp_bf16: bf16 segment %d must be 128 wide", s);
        RMP_REQUIRE(d.seg_src[s] && d.seg_width[s] > 0 && d.seg_width[s] % 128 == 0, RMP_E_SHAPE,
                    "rmp_row_mlp_bf16: segment %d width %d must be a multiple of 128", s, d.seg_width[s]);
        RMP_REQUIRE(d.seg_bf16[s] ? aligned32(d.seg_src[s]) : aligned16(d.seg_src[s]), RMP_E_ALIGN,
                    "rmp_row_mlp_bf16: segment %d unaligned (fp32 rows 16 bytes, bf16 rows 32 bytes: 256-bit gathers)", s);
        p.seg_src[s] = d.seg_src[s]; p.seg_idx[s] = d.seg_idx[s]; p.seg_width[s] = d.seg_width[s];
        k0 += d.seg_width[s];
    }
    p.num_segs = d.num_segs;
    p.a_pieces = k0 / 128;
    RMP_REQUIRE(n_bf16 == 0 || (n_bf16 == d.num_segs && p.a_pieces <= 3 && operand_format == RMP_OPERAND_BF16), RMP_E_SHAPE,
                "rmp_row_mlp_bf16: bf16 segments need all (<= 3) segments bf16 and bf16 operands");
    p.all_bf16 = n_bf16 > 0;
    p.out16 = reinterpret_cast<uint16_t*>(d.out_bf16);
    RMP_REQUIRE(!d.out_bf16 || (d.epilogue != RMP_EPI_SEGSUM && aligned32(d.out_bf16)), RMP_E_ARG,
                "rmp_row_mlp_bf16: out_bf16 goes with STORE / RESID_ADD and must be 16-byte aligned");
    if (p.a_pieces <= 3) {
        int pc = 0;
        for (int s = 0; s < d.num_segs; ++s)
            for (int o = 0; o < d.seg_width[s]; o += 128, ++pc) {
                p.piece_src[pc] = d.seg_src[s] + o; p.piece_idx[pc] = d.seg_idx[s]; p.piece_stride[pc] = d.seg_width[s];
            }
    }
    RMP_REQUIRE(k0 == d.k_in[0], RMP_E_SHAPE, "rmp_row_mlp_bf16: segments give %d inputs, stage 0 expects %d", k0, d.k_in[0]);
    int kmax = 0, nmax = 0;
    CUtensorMap maps[tc::kRowMaxStages];
    memset(maps, 0, sizeof(maps));
    for (int s = 0; s < d.num_stages; ++s) {
        const int K = d.k_in[s], N = d.n_out[s];
        RMP_REQUIRE(W_bf16[s] && aligned16(W_bf16[s]) && d.bias[s] && aligned16(d.bias[s]), RMP_E_ALIGN,
                    "rmp_row_mlp_bf16: stage %d weight / bias null or unaligned", s);
        RMP_REQUIRE(K % 64 == 0 && K >= 64 && K <= 512 && N % 128 == 0 && N >= 128 && N <= 512, RMP_E_SHAPE,
                    "rmp_row_mlp_bf16: stage %d is %d -> %d; K must be a multiple of 64 and N of 128, both <= 512", s, K, N);
        RMP_REQUIRE(s == 0 || K == d.n_out[s - 1], RMP_E_SHAPE, "rmp_row_mlp_bf16: stage %d input width mismatch", s);
        p.st[s].K = K; p.st[s].N = N; p.st[s].gelu = d.gelu[s]; p.st[s].bias = d.bias[s];
        kmax = K > kmax ? K : kmax;
        nmax = N > nmax ? N : nmax;
        const int rc = tc::make_map_bf16(&maps[s], W_bf16[s], (uint64_t)N, (uint64_t)K, 128);
        if (rc != RMP_OK) return rc;
    }
    for (int s = d.num_stages; s < tc::kRowMaxStages; ++s) maps[s] = maps[0];
    // every intermediate activation is written back into the A buffer: it has to hold the widest of them too
    for (int s = 0; s + 1 < d.num_stages; ++s) kmax = d.n_out[s] > kmax ? d.n_out[s] : kmax;
    p.S = d.num_stages;
    p.epilogue = d.epilogue;
    if (d.epilogue == RMP_EPI_RESID_ADD) RMP_REQUIRE(d.resid && aligned32(d.resid), RMP_E_ARG, "rmp_row_mlp_bf16: resid missing or not 32-byte aligned");
    if (d.epilogue == RMP_EPI_SEGSUM) RMP_REQUIRE(d.seg_id, RMP_E_ARG, "rmp_row_mlp_bf16: seg_id missing");
    RMP_REQUIRE(aligned32(d.out), RMP_E_ALIGN, "rmp_row_mlp_bf16: out not 32-byte aligned (256-bit stores)");
    p.resid = d.resid; p.seg_id = d.seg_id; p.row_scale = d.row_scale; p.alpha = d.alpha; p.out = d.out;
    p.n_rows = d.rows_cap;
    p.timeline = reinterpret_cast<long long*>(g_row_tc_timeline);
    p.a_kblocks = kmax / 64;
    p.tmem_cols = nmax <= 128 ? 128u : (nmax <= 256 ? 256u : 512u);
    const size_t smem = 1024 + (size_t)p.a_kblocks * tc::kKBlockBytes + (size_t)tc::kRowRing * tc::kRowSlotBytes +
                        sizeof(tc::RowBarriers) + 64;
    RMP_REQUIRE(smem <= 227 * 1024, RMP_E_SHAPE, "rmp_row_mlp_bf16: needs %zu B of shared memory", smem);
    const int f = operand_format == RMP_OPERAND_FP16 ? 1 : 0;
    auto kern = f ? tc::row_mlp_tc_kernel<true> : tc::row_mlp_tc_kernel<false>;
    // per device and not cached: the opt-in is a per-device function attribute (a process may drive several GPUs)
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(d.rows_cap, tc::kTileM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    kern<<<grid, tc::kThreads, smem, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
    return check_launch("rmp_row_mlp_tc");
}

extern "C" int rmp_row_mlp_bf16(const rmp_mlp_desc* dp, const void* const* W_bf16, void* stream) {
    return rmp_row_mlp_tc(dp, W_bf16, RMP_OPERAND_BF16, stream);
}
