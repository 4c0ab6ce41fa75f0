// fp32-parity row MLP on tcgen05 tensor cores: every fp32 operand is carried as TWO bf16 terms, x = x_hi + x_lo with
// x_hi = rn_bf16(x), x_lo = rn_bf16(x - x_hi)  (|x - x_hi - x_lo| <= 2^-18 |x|), and every product as three MMAs
//   x w ~= x_hi w_hi + x_lo w_hi + x_hi w_lo                  (dropped term x_lo w_lo <= 2^-18 |x w|)
// accumulated in fp32 in TMEM.  Bias, the exact erf GELU (nn.GELU() default, rdesign/model/mpnn.py:17) and every
// epilogue stay fp32 on the CUDA cores, so the arm keeps the fp32 path's 1e-4 logit parity with the reference's CPU
// path (measured 1.5e-5 on the C2 shapes) at 3 bf16 MMAs per fp32 one.  Same descriptor as rmp_row_mlp.
//
// One CTA owns 128-row tiles (persistent over tiles), two CTAs per SM when the activation fits (K <= 128), so that one
// CTA's erf epilogue overlaps the other's MMAs and gathers.  Per tile:
//   prologue : the 8 epilogue warps gather the fp32 input rows, split them and write the two swizzled A operands;
//   stage s  : weight slabs [128 n][64 k] (hi, then lo) arrive through a TMA ring; one elected thread issues
//              A_hi W_hi, A_lo W_hi, A_hi W_lo per slab pair; accumulator [128][N_s] fp32 in TMEM;
//              optional gathered fp32 "add" rows (the per-node projections that replace the [h_V[ctr] | h_V[nbr]]
//              columns of the message MLP's first Linear) are summed and written into the accumulator with
//              tcgen05.st, and stage 0's MMAs accumulate onto them;
//   epilogue : tcgen05.ld -> + bias -> erf GELU -> split -> next A operand; the last stage applies STORE / SEGSUM /
//              RESID_ADD straight from TMEM to HBM.
#include "tc_common.cuh"

namespace rmp {
namespace tc {

constexpr int kX3MaxStages = RMP_MAX_STAGES;
constexpr int kX3MaxRing = 6;
constexpr int kX3SlotBytes = 128 * 128;           // [128 n][64 k] bf16

struct X3Stage {
    int K, N, gelu;
    const float* bias;
};
struct X3Params {
    const float* seg_src[RMP_MAX_SEGS];
    const int32_t* seg_idx[RMP_MAX_SEGS];
    int seg_width[RMP_MAX_SEGS];
    int num_segs;
    X3Stage st[kX3MaxStages];
    int S;
    int epilogue;
    const float* resid;
    const int32_t* seg_id;
    const float* row_scale;
    const float* gain;
    const float* beta;
    float alpha;
    float* out;
    int n_rows;
    int a_kblocks;
    int ring;
    const float* add_src[2];                      // fp32 [*, N_0] rows added to stage 0's pre-activation
    const int32_t* add_idx[2];
    int add_stride;                               // floats between consecutive add rows
    uint32_t tmem_cols;
    long long* timeline;                          // optional clock64 trace of CTA 0's first epilogue thread (tuning aid)
};
struct alignas(8) X3Barriers {
    uint64_t w_full[kX3MaxRing], w_empty[kX3MaxRing], acc_full, act_ready;
};

// two fp32 values -> (hi, lo) bf16x2 words; element a in the low half like pack_bf16
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16(a, b);
    lo = pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

// erf-form GELU without the branch of erff: erfc(t) = 2^(t q(t)) on t = |x| / sqrt(2) in [0, 3.95] (q of degree 7 fitted
// to log2 erfc with the GELU's own weight, one MUFU.EX2), then gelu = x/2 * (2 - erfc) for x >= 0 and x/2 * erfc below.
// |gelu - exact| <= 4.5e-7 over [-9, 9], the same bound 0.5 x (1 + erff(x / sqrt 2)) has in fp32 (the final rounding).
__device__ __forceinline__ float gelu_erf_poly(float x) {
    const float t = fminf(fabsf(x) * 0.70710678118654752440f, 3.95f);
    float q = -3.043286051e-05f;
    q = fmaf(q, t, 3.174262238e-04f);
    q = fmaf(q, t, -1.051778789e-03f);
    q = fmaf(q, t, -1.539106597e-03f);
    q = fmaf(q, t, 2.898170054e-02f);
    q = fmaf(q, t, -1.488533765e-01f);
    q = fmaf(q, t, -9.183242917e-01f);
    q = fmaf(q, t, -1.627916813e+00f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q * t));
    return (0.5f * x) * (x >= 0.f ? 2.f - e : e);
}

// 32 lanes x 32 consecutive fp32 columns, thread i of the warp writes row (lane base + i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void ldg_row32(const float* src, float4 (&a)[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = src ? __ldg(reinterpret_cast<const float4*>(src) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(kThreads, 2)
row_mlp_x3_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                  const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3,
                  const __grid_constant__ CUtensorMap tm4, const __grid_constant__ CUtensorMap tm5, const X3Params p) {
    // the 128-byte swizzle needs a 1024-aligned base; with no static shared memory the dynamic window starts aligned, and
    // carrying no slack lets two CTAs of the K = 128 shape hold a 3-slot weight ring each (checked, never assumed)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (smem_u32(smem_raw) & 1023u) __trap();
    uint8_t* smem = smem_raw;
    uint8_t* sAhi = smem;                                        // [a_kblocks][128][64] bf16, high terms
    uint8_t* sAlo = sAhi + (size_t)p.a_kblocks * kKBlockBytes;   // low terms
    uint8_t* sW = sAlo + (size_t)p.a_kblocks * kKBlockBytes;     // weight ring
    X3Barriers* bars = reinterpret_cast<X3Barriers*>(sW + (size_t)p.ring * kX3SlotBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    int* s_seg = reinterpret_cast<int*>(s_tmem + 4);             // [128] destination row of every tile row (SEGSUM)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.n_rows + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.S;
    const uint32_t ring = (uint32_t)p.ring;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kX3MaxRing; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->act_ready, kEpiThreads);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const CUtensorMap* maps[kX3MaxStages] = {&tm0, &tm1, &tm2, &tm3, &tm4, &tm5};

    if (warp == 0) {
        // ===================== TMA producer: per (k-block, column chunk) the hi slab, then the lo slab =====================
        if (elect_one()) {
            for (int s = 0; s < S; ++s) tma_prefetch_desc(maps[s]);
            uint32_t n = 0;
            for (int i = 0; i < my_tiles; ++i)
                for (int s = 0; s < S; ++s)
                    for (int kb = 0; kb < p.st[s].K / 64; ++kb)
                        for (int c = 0; c < p.st[s].N / 128; ++c)
                            for (int t = 0; t < 2; ++t, ++n) {
                                const uint32_t slot = n % ring;
                                mbar_wait(&bars->w_empty[slot], ((n / ring) & 1) ^ 1);
                                mbar_expect_tx(&bars->w_full[slot], (uint32_t)kX3SlotBytes);
                                tma_load_2d(sW + slot * kX3SlotBytes, maps[s], &bars->w_full[slot], kb * 64,
                                            t * p.st[s].N + c * 128);
                            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            uint32_t n = 0, ph_act = 0;
            const uint32_t hi_addr = smem_u32(sAhi), lo_addr = smem_u32(sAlo);
            const uint32_t idesc = instr_desc_bf16(128, 128, 0);
            const bool preloaded = p.add_src[0] != nullptr;      // stage 0 accumulates onto the add rows held in TMEM
            for (int i = 0; i < my_tiles; ++i)
                for (int s = 0; s < S; ++s) {
                    mbar_wait(&bars->act_ready, ph_act); ph_act ^= 1;
                    tc_fence_after();
                    for (int kb = 0; kb < p.st[s].K / 64; ++kb)
                        for (int c = 0; c < p.st[s].N / 128; ++c) {
                            const uint32_t d = tmem_base + (uint32_t)(c * 128);
                            uint32_t slot = n % ring;
                            mbar_wait(&bars->w_full[slot], (n / ring) & 1);
                            tc_fence_after();
                            uint32_t w_addr = smem_u32(sW + slot * kX3SlotBytes);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t wd = smem_desc_sw128(w_addr + k * 32, 16, 1024);
                                umma_bf16(d, smem_desc_sw128(lo_addr + kb * kKBlockBytes + k * 32, 16, 1024), wd, idesc,
                                          (kb | k) > 0 || (s == 0 && preloaded));
                                umma_bf16(d, smem_desc_sw128(hi_addr + kb * kKBlockBytes + k * 32, 16, 1024), wd, idesc, 1);
                            }
                            umma_commit(&bars->w_empty[slot]);
                            ++n;
                            slot = n % ring;
                            mbar_wait(&bars->w_full[slot], (n / ring) & 1);
                            tc_fence_after();
                            w_addr = smem_u32(sW + slot * kX3SlotBytes);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d, smem_desc_sw128(hi_addr + kb * kKBlockBytes + k * 32, 16, 1024),
                                          smem_desc_sw128(w_addr + k * 32, 16, 1024), idesc, 1);
                            umma_commit(&bars->w_empty[slot]);
                            ++n;
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
            if (tl_on && tl_n < 240) { p.timeline[tl_n * 2] = tag; p.timeline[tl_n * 2 + 1] = clock64(); ++tl_n; }
        };
        stamp(1);

        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int r = tile * kTileM + row;
            const bool valid = r < p.n_rows;
            // ---- gathered add rows (the per-node projections, n_out[0] = 128): summed in fp32 and written into the
            //      accumulator the stage-0 MMAs then accumulate onto.  Each thread stores the columns of its row that it
            //      reads back in the epilogues, so program order covers the reuse of the accumulator across tiles ----
            if (p.add_src[0]) {
                // coalesced: one warp per row, lane = 16-byte chunk of the 512-byte row, 8 rows x 2 sources in flight;
                // staged as fp32 [128][128] (chunk j of row r at j ^ (r & 31): conflict-free both ways) in the A buffers,
                // which nothing uses between the last MMA of the previous tile and this tile's gather
                for (int r0 = 0; r0 < 16; r0 += 8) {
                    int i0[8], i1[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int rg = tile * kTileM + ew * 16 + r0 + rr;
                        const bool ok = rg < p.n_rows;
                        i0[rr] = ok ? (p.add_idx[0] ? __ldg(p.add_idx[0] + rg) : rg) : -1;
                        i1[rr] = ok && p.add_src[1] ? (p.add_idx[1] ? __ldg(p.add_idx[1] + rg) : rg) : -1;
                    }
                    float4 x0[8], x1[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        x0[rr] = x1[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (i0[rr] >= 0) x0[rr] = __ldg(reinterpret_cast<const float4*>(p.add_src[0] + (size_t)i0[rr] * p.add_stride) + lane);
                        if (i1[rr] >= 0) x1[rr] = __ldg(reinterpret_cast<const float4*>(p.add_src[1] + (size_t)i1[rr] * p.add_stride) + lane);
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int rl = ew * 16 + r0 + rr;
                        *reinterpret_cast<float4*>(sAhi + rl * 512 + ((lane ^ (rl & 31)) << 4)) =
                            make_float4(x0[rr].x + x1[rr].x, x0[rr].y + x1[rr].y, x0[rr].z + x1[rr].z, x0[rr].w + x1[rr].w);
                    }
                }
                epi_bar_sync();
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    float v[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int j = half * 16 + (c0 >> 2) + q;
                        const float4 t = *reinterpret_cast<const float4*>(sAhi + row * 512 + ((j ^ (row & 31)) << 4));
                        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                    }
                    tmem_st32(t_row + (uint32_t)(half * 64 + c0), v);
                }
                epi_bar_sync();                                    // the gather below overwrites the staging rows
            }
            stamp(4);
            // ---- prologue: gather + concat + split -> the two swizzled A operands.  One warp per row, a lane holds 4
            //      consecutive channels of every 128-wide piece; the loads of all 16 rows are issued before the first use ----
            {
                int kb0 = 0;
                for (int sg = 0; sg < p.num_segs; ++sg) {
                    const int wdt = p.seg_width[sg];
                    const float* src = p.seg_src[sg];
                    const int32_t* idx = p.seg_idx[sg];
                    for (int pc = 0; pc < wdt; pc += 128) {
                        // all row indices first, then all row loads: a row load waits for its own index only
                        int ri[16];
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            const int rg = tile * kTileM + ew * 16 + rr;
                            ri[rr] = rg < p.n_rows ? (idx ? __ldg(idx + rg) : rg) : -1;
                        }
                        float4 xs[16];
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            xs[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ri[rr] >= 0) xs[rr] = __ldg(reinterpret_cast<const float4*>(src + (size_t)ri[rr] * wdt + pc) + lane);
                        }
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            const int rl = ew * 16 + rr;
                            const uint32_t off = (uint32_t)(kb0 + (pc >> 6) + (lane >> 4)) * kKBlockBytes +
                                                 sw128_chunk(rl, (lane & 15) >> 1) + (lane & 1) * 8;
                            uint32_t h0, l0, h1, l1;
                            split2(xs[rr].x, xs[rr].y, h0, l0);
                            split2(xs[rr].z, xs[rr].w, h1, l1);
                            *reinterpret_cast<uint2*>(sAhi + off) = make_uint2(h0, h1);
                            *reinterpret_cast<uint2*>(sAlo + off) = make_uint2(l0, l1);
                        }
                    }
                    kb0 += wdt >> 6;
                }
            }
            stamp(2);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&bars->act_ready);

            // ---- stage epilogues ----
            for (int s = 0; s < S; ++s) {
                const X3Stage stg = p.st[s];
                const int ncols = stg.N >> 1;                      // columns of this half
                const int cb = half * ncols;
                const bool last = s + 1 == S;
                const bool seg_staged = last && p.epilogue == RMP_EPI_SEGSUM && stg.N == 128;
                const bool norm_staged = last && (p.epilogue == RMP_EPI_NORMALIZE || p.epilogue == RMP_EPI_RESID_LN);
                float rs = 1.f;
                int dst_row = r;
                if (last && valid) {
                    if (p.row_scale) rs = __ldg(p.row_scale + r);
                    if (p.epilogue == RMP_EPI_SEGSUM) { dst_row = __ldg(p.seg_id + r); rs *= p.alpha; }
                }
                mbar_wait(&bars->acc_full, ph_acc); ph_acc ^= 1;
                tc_fence_after();
                stamp(10 + s);
                const int nch = ncols / 32;
                for (int ch = 0; ch < nch; ++ch) {
                    const int col0 = cb + ch * 32;
                    float v[32];
                    tmem_ld32(t_row + (uint32_t)col0, v);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(stg.bias + col0) + q);
                        v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
                    }
                    if (stg.gelu) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = gelu_erf_poly(v[q]);
                    }
                    if (!last) {
                        const uint32_t kb_off = (uint32_t)(col0 >> 6) * kKBlockBytes;
                        const int j0 = (col0 & 63) >> 3;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t h[4], l[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) split2(v[8 * j + 2 * q], v[8 * j + 2 * q + 1], h[q], l[q]);
                            const uint32_t off = kb_off + sw128_chunk(row, j0 + j);
                            *reinterpret_cast<uint4*>(sAhi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                            *reinterpret_cast<uint4*>(sAlo + off) = make_uint4(l[0], l[1], l[2], l[3]);
                        }
                        continue;
                    }
                    if (norm_staged) {                             // whole rows are needed: staged like the segment sums
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int j = (col0 >> 2) + q;
                            *reinterpret_cast<float4*>(sAhi + row * 512 + ((j ^ (row & 31)) << 4)) =
                                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        }
                        continue;
                    }
                    if (seg_staged) {
                        // rows of one centre node are consecutive: the tile is staged in the (now idle) A buffers and
                        // reduced per column below, so a segment costs one coalesced red per column instead of one
                        // 16-byte red per row and 4 columns
                        const bool keep = valid && rs != 0.f;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int j = (col0 >> 2) + q;
                            *reinterpret_cast<float4*>(sAhi + row * 512 + ((j ^ (row & 31)) << 4)) =
                                keep ? make_float4(v[4 * q] * rs, v[4 * q + 1] * rs, v[4 * q + 2] * rs, v[4 * q + 3] * rs)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        continue;
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
#pragma unroll
                        for (int q = 0; q < 4; ++q) stg256f(dst + 8 * q, &v[8 * q]);
                    }
                }
                if (seg_staged) {
                    if (half == 0) s_seg[row] = valid ? dst_row : -1;
                    epi_bar_sync();
                    const int te = threadIdx.x - 64, col = te & 127, rb = (te >> 7) * 64;
                    int cur = s_seg[rb];
                    float acc = 0.f;
                    for (int r8 = 0; r8 < 64; r8 += 8) {
                        int sg[8];
                        float x[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int rw = rb + r8 + k;
                            sg[k] = s_seg[rw];
                            x[k] = *reinterpret_cast<const float*>(sAhi + rw * 512 + (((col >> 2) ^ (rw & 31)) << 4) + (col & 3) * 4);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if (sg[k] != cur) {
                                if (cur >= 0) atomicAdd(p.out + (size_t)cur * 128 + col, acc);
                                cur = sg[k];
                                acc = 0.f;
                            }
                            acc += x[k];
                        }
                    }
                    if (cur >= 0) atomicAdd(p.out + (size_t)cur * 128 + col, acc);
                    epi_bar_sync();                                // the next tile's prologue overwrites the staging rows
                }
                if (norm_staged) {
                    // reference Normalize (unbiased variance, eps inside and outside the root; rdesign/model/functional.py:79-96)
                    // or nn.LayerNorm of resid + y (rdesign/model/mpnn.py:38), one warp per row, coalesced 512-byte rows
                    epi_bar_sync();
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gain) + lane);
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.beta) + lane);
                    for (int rr = 0; rr < 16; ++rr) {
                        const int rl = ew * 16 + rr;
                        const int rg = tile * kTileM + rl;
                        if (rg >= p.n_rows) break;
                        float4 x = *reinterpret_cast<const float4*>(sAhi + rl * 512 + ((lane ^ (rl & 31)) << 4));
                        if (p.epilogue == RMP_EPI_RESID_LN) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)rg * 128) + lane);
                            x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
                        }
                        const float mu = warp_sum((x.x + x.y) + (x.z + x.w)) * (1.f / 128.f);
                        x.x -= mu; x.y -= mu; x.z -= mu; x.w -= mu;
                        const float ss = warp_sum((x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w));
                        const float inv = p.epilogue == RMP_EPI_NORMALIZE ? 1.f / (sqrtf(ss * (1.f / 127.f) + 1e-6f) + 1e-6f)
                                                                          : 1.f / sqrtf(ss * (1.f / 128.f) + 1e-5f);
                        reinterpret_cast<float4*>(p.out + (size_t)rg * 128)[lane] =
                            make_float4(g4.x * (x.x * inv) + b4.x, g4.y * (x.y * inv) + b4.y, g4.z * (x.z * inv) + b4.z,
                                        g4.w * (x.w * inv) + b4.w);
                    }
                    epi_bar_sync();                                // the next tile's prologue overwrites the staging rows
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

static void* g_row_x3_timeline = nullptr;
extern "C" void rmp_set_row_x3_timeline(void* buf) { g_row_x3_timeline = buf; }

extern "C" int rmp_row_mlp_x3(const rmp_mlp_desc* dp, const void* const* W_split, const float* add0, const int32_t* add0_idx,
                              const float* add1, const int32_t* add1_idx, int add_stride, void* stream) {
    RMP_REQUIRE(dp && W_split, RMP_E_ARG, "rmp_row_mlp_x3: null descriptor");
    const rmp_mlp_desc& d = *dp;
    RMP_REQUIRE(d.num_segs >= 1 && d.num_segs <= RMP_MAX_SEGS, RMP_E_ARG, "rmp_row_mlp_x3: num_segs=%d", d.num_segs);
    RMP_REQUIRE(d.num_stages >= 1 && d.num_stages <= RMP_MAX_STAGES, RMP_E_ARG, "rmp_row_mlp_x3: num_stages=%d", d.num_stages);
    RMP_REQUIRE(d.out && d.rows_cap >= 0 && !d.rows_dev && !d.out_bf16, RMP_E_ARG,
                "rmp_row_mlp_x3: fp32 output and a host row count only");
    RMP_REQUIRE(d.epilogue >= RMP_EPI_STORE && d.epilogue <= RMP_EPI_RESID_ADD, RMP_E_SHAPE, "rmp_row_mlp_x3: epilogue %d",
                d.epilogue);
    if (d.epilogue == RMP_EPI_NORMALIZE || d.epilogue == RMP_EPI_RESID_LN)
        RMP_REQUIRE(d.gain && d.beta && aligned16(d.gain) && aligned16(d.beta) && d.n_out[d.num_stages - 1] == 128 && !d.row_scale,
                    RMP_E_ARG, "rmp_row_mlp_x3: Normalize / LayerNorm epilogues need gain, beta and 128 output channels");
    RMP_REQUIRE(add0 || !add1, RMP_E_ARG, "rmp_row_mlp_x3: add1 without add0");
    RMP_REQUIRE(!add0 || (add_stride >= d.n_out[0] && add_stride % 4 == 0 && d.n_out[0] == 128), RMP_E_ARG,
                "rmp_row_mlp_x3: add rows need n_out[0] = 128 and a stride that is a multiple of 4 (got %d)", add_stride);
    if (d.rows_cap == 0) return RMP_OK;
    tc::X3Params p;
    memset(&p, 0, sizeof(p));
    int k0 = 0;
    for (int s = 0; s < d.num_segs; ++s) {
        RMP_REQUIRE(!d.seg_bf16[s], RMP_E_SHAPE, "rmp_row_mlp_x3: segment %d must be fp32", s);
        RMP_REQUIRE(d.seg_src[s] && d.seg_width[s] > 0 && d.seg_width[s] % 128 == 0, RMP_E_SHAPE,
                    "rmp_row_mlp_x3: segment %d width %d must be a multiple of 128", s, d.seg_width[s]);
        RMP_REQUIRE(aligned16(d.seg_src[s]), RMP_E_ALIGN, "rmp_row_mlp_x3: segment %d not 16-byte aligned", s);
        p.seg_src[s] = d.seg_src[s]; p.seg_idx[s] = d.seg_idx[s]; p.seg_width[s] = d.seg_width[s];
        k0 += d.seg_width[s];
    }
    p.num_segs = d.num_segs;
    RMP_REQUIRE(k0 == d.k_in[0], RMP_E_SHAPE, "rmp_row_mlp_x3: segments give %d inputs, stage 0 expects %d", k0, d.k_in[0]);
    int kmax = 0, nmax = 0;
    CUtensorMap maps[tc::kX3MaxStages];
    memset(maps, 0, sizeof(maps));
    for (int s = 0; s < d.num_stages; ++s) {
        const int K = d.k_in[s], N = d.n_out[s];
        RMP_REQUIRE(W_split[s] && aligned16(W_split[s]) && d.bias[s] && aligned16(d.bias[s]), RMP_E_ALIGN,
                    "rmp_row_mlp_x3: stage %d weight / bias null or unaligned", s);
        RMP_REQUIRE(K % 64 == 0 && K >= 64 && K <= 384 && N % 128 == 0 && N >= 128 && N <= 256, RMP_E_SHAPE,
                    "rmp_row_mlp_x3: stage %d is %d -> %d; K a multiple of 64 up to 384, N 128 or 256", s, K, N);
        RMP_REQUIRE(s == 0 || K == d.n_out[s - 1], RMP_E_SHAPE, "rmp_row_mlp_x3: stage %d input width mismatch", s);
        p.st[s].K = K; p.st[s].N = N; p.st[s].gelu = d.gelu[s]; p.st[s].bias = d.bias[s];
        kmax = K > kmax ? K : kmax;
        nmax = N > nmax ? N : nmax;
        // [2][N][K]: rows [0, N) the high terms, rows [N, 2N) the low terms
        const int rc = tc::make_map_bf16(&maps[s], W_split[s], (uint64_t)(2 * N), (uint64_t)K, 128);
        if (rc != RMP_OK) return rc;
    }
    for (int s = d.num_stages; s < tc::kX3MaxStages; ++s) maps[s] = maps[0];
    for (int s = 0; s + 1 < d.num_stages; ++s) kmax = d.n_out[s] > kmax ? d.n_out[s] : kmax;
    p.S = d.num_stages;
    p.epilogue = d.epilogue;
    if (d.epilogue == RMP_EPI_RESID_ADD || d.epilogue == RMP_EPI_RESID_LN)
        RMP_REQUIRE(d.resid && aligned32(d.resid), RMP_E_ARG, "rmp_row_mlp_x3: resid missing or not 32-byte aligned");
    if (d.epilogue == RMP_EPI_SEGSUM) RMP_REQUIRE(d.seg_id, RMP_E_ARG, "rmp_row_mlp_x3: seg_id missing");
    RMP_REQUIRE(aligned32(d.out), RMP_E_ALIGN, "rmp_row_mlp_x3: out not 32-byte aligned (256-bit stores)");
    RMP_REQUIRE((!add0 || aligned16(add0)) && (!add1 || aligned16(add1)), RMP_E_ALIGN, "rmp_row_mlp_x3: add rows not 16-byte aligned");
    p.resid = d.resid; p.seg_id = d.seg_id; p.row_scale = d.row_scale; p.alpha = d.alpha; p.out = d.out;
    p.gain = d.gain; p.beta = d.beta;
    p.n_rows = d.rows_cap;
    p.timeline = reinterpret_cast<long long*>(g_row_x3_timeline);
    p.add_src[0] = add0; p.add_idx[0] = add0_idx; p.add_src[1] = add1; p.add_idx[1] = add1_idx; p.add_stride = add_stride;
    p.a_kblocks = kmax / 64;
    p.tmem_cols = nmax <= 128 ? 128u : 256u;
    // two CTAs per SM (228 KB, 1 KB reserved per CTA) when the operands leave room for a 2-slot ring; else the deepest
    // ring one CTA can hold
    const size_t fixed = 2 * (size_t)p.a_kblocks * tc::kKBlockBytes + sizeof(tc::X3Barriers) + 16 + 512;
    int ctas = 1;
    if (2 * (fixed + 3 * tc::kX3SlotBytes + 1024) <= 228 * 1024) { ctas = 2; p.ring = 3; }
    else if (2 * (fixed + 2 * tc::kX3SlotBytes + 1024) <= 228 * 1024) { ctas = 2; p.ring = 2; }
    else {
        p.ring = (int)((227 * 1024 - fixed) / tc::kX3SlotBytes);
        if (p.ring > tc::kX3MaxRing) p.ring = tc::kX3MaxRing;
    }
    RMP_REQUIRE(p.ring >= 2, RMP_E_SHAPE, "rmp_row_mlp_x3: %d-wide activations do not fit shared memory", kmax);
    const size_t smem = fixed + (size_t)p.ring * tc::kX3SlotBytes;
    cudaFuncSetAttribute(tc::row_mlp_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(d.rows_cap, tc::kTileM);
    const int grid = n_tiles < kNumSMs * ctas ? n_tiles : kNumSMs * ctas;
    tc::row_mlp_x3_kernel<<<grid, tc::kThreads, smem, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4],
                                                                              maps[5], p);
    return check_launch("rmp_row_mlp_x3");
}
