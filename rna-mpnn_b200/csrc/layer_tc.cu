// bf16 tensor-core arm of the message-passing layer (RMP_PREC_BF16): tcgen05.mma with
// TMEM accumulators, TMA-fed 128-byte-swizzled tiles, warp-specialised persistent CTAs.
//
// Edge kernel (this file): for a tile of 128 consecutive edges
//     pre0 = h_E[e] . W1e^T            (tcgen05, K=128)   + Pc[centre(e)] + Pn[neighbour(e)]
//     m1 = GELU(pre0);  m2 = GELU(m1 . W2^T + b2);  m3 = GELU(m2 . W3^T + b3)
//     agg[node] += (1/scale) * sum_{e in node} m3[e]       (tcgen05: selector matrix . m3)
// where Pc = h_V . W1c^T + b1 and Pn = h_V . W1n^T are per-node projections (W1 = [W1e|W1c|W1n]
// column blocks, rdesign/model/rdesign.py:85), so the [Ne,3H] concatenation and the three
// [Ne,H] activations of the reference (rdesign/model/mpnn.py:33) never reach HBM: per edge the
// kernel reads 256 B of bf16 h_E and the L2-resident projections, nothing else.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..9 = epilogue (thread = one edge row, two column halves).
#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

namespace rmp {
namespace tc {

constexpr int kTileM = 128;                 // edges per tile (UMMA M)
constexpr int kKBlockBytes = kTileM * 128;  // one [128 rows][64 bf16] swizzle-128B block
constexpr int kTileBytes = 2 * kKBlockBytes;
constexpr int kMaxStages = 4;
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr uint32_t kTmemCols = 256;         // [0,128) stage accumulator, [128,256) neighbour sums

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, bf16 operands, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, 128-byte swizzle, Blackwell version bit (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor kind::f16: D fp32, A/B bf16, A K-major, B K-major or MN-major, M x N tile.
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk `j` (8 bf16) of row `r` inside one [128][64] swizzle-128B block
__device__ __forceinline__ uint32_t sw128_chunk(int r, int j) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

struct EdgeParams {
    const float* Pc;                 // [Nv, ld_p] centre projection incl. stage-0 bias
    const float* Pn;                 // [Nv, ld_p] neighbour projection
    int ld_p;
    const int32_t* ctr;
    const int32_t* nbr;
    const float* bias[kMaxStages];   // stages >= 1
    int num_stages;
    float alpha;
    float* agg;                      // [Nv,128] fp32, pre-zeroed
    int n_edges;
    float* debug;                    // optional [num_stages][n_edges][128] post-activation dump
};

struct alignas(8) Barriers {
    uint64_t w_full, a_full[2], a_empty[2], acc_full, act_ready, agg_full, agg_empty;
};

__global__ void __launch_bounds__(kThreads, 1)
edge_mlp_tc_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_w,
                   const EdgeParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;                                         // [S][2][128][64] bf16
    uint8_t* sA = sW + (size_t)p.num_stages * kTileBytes;       // [2 buffers][2][128][64] bf16
    uint8_t* sSel = sA + 2 * kTileBytes;                        // [2][128][64] bf16 selector
    Barriers* bars = reinterpret_cast<Barriers*>(sSel + kTileBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    int* s_ctr = reinterpret_cast<int*>(s_tmem + 4);            // [128]
    int* s_start = s_ctr + kTileM;                              // [128]
    int* s_end = s_start + kTileM;                              // [128]
    int* s_misc = s_end + kTileM;                               // [0] = nslots

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.n_edges + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.num_stages;

    if (threadIdx.x == 0) {
        mbar_init(&bars->w_full, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&bars->a_full[b], 1); mbar_init(&bars->a_empty[b], 1); }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->act_ready, kEpiThreads);
        mbar_init(&bars->agg_full, 1);
        mbar_init(&bars->agg_empty, kEpiThreads);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&tm_e);
            tma_prefetch_desc(&tm_w);
            mbar_expect_tx(&bars->w_full, (uint32_t)(S * kTileBytes));
            for (int s = 0; s < S; ++s)
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d(sW + (size_t)s * kTileBytes + kb * kKBlockBytes, &tm_w, &bars->w_full, kb * 64, s * 128);
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1;
                const int tile = blockIdx.x + i * gridDim.x;
                mbar_wait(&bars->a_empty[b], ((i >> 1) & 1) ^ 1);
                mbar_expect_tx(&bars->a_full[b], kTileBytes);
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d(sA + (size_t)b * kTileBytes + kb * kKBlockBytes, &tm_e, &bars->a_full[b], kb * 64,
                                tile * kTileM);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_kk = instr_desc_bf16(128, 128, 0);
            constexpr uint32_t idesc_kmn = instr_desc_bf16(128, 128, 1);
            const uint32_t acc = tmem_base, agg = tmem_base + 128;
            uint32_t ph_act = 0, ph_ae = 0;
            mbar_wait(&bars->w_full, 0);
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1;
                const uint32_t a_addr = smem_u32(sA + (size_t)b * kTileBytes);
                mbar_wait(&bars->a_full[b], (i >> 1) & 1);
                for (int s = 0; s < S; ++s) {
                    if (s > 0) { mbar_wait(&bars->act_ready, ph_act); ph_act ^= 1; }
                    tc_fence_after();
                    const uint32_t w_addr = smem_u32(sW + (size_t)s * kTileBytes);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {               // K = 128 = 8 x UMMA_K(16)
                        const uint32_t off = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                        umma_bf16(acc, smem_desc_sw128(a_addr + off, 16, 1024), smem_desc_sw128(w_addr + off, 16, 1024),
                                  idesc_kk, k > 0);
                    }
                    umma_commit(&bars->acc_full);
                }
                // neighbour sums: agg[slot][ch] = sum_r Sel[slot][r] * m3[r][ch]; m3 is read MN-major in place
                mbar_wait(&bars->act_ready, ph_act); ph_act ^= 1;
                mbar_wait(&bars->agg_empty, ph_ae ^ 1); ph_ae ^= 1;
                tc_fence_after();
                const uint32_t sel_addr = smem_u32(sSel);
#pragma unroll
                for (int k = 0; k < 8; ++k) {                   // K = 128 edge rows = 8 x 16
                    const uint32_t a_off = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                    const uint32_t b_off = (uint32_t)(k * 2048);     // 16 rows of 128 B
                    umma_bf16(agg, smem_desc_sw128(sel_addr + a_off, 16, 1024),
                              smem_desc_sw128(a_addr + b_off, kKBlockBytes, 1024), idesc_kmn, k > 0);
                }
                umma_commit(&bars->agg_full);
                umma_commit(&bars->a_empty[b]);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps =====================
        const int half = (warp - 2) >> 2;            // output column half == K-block of the next stage
        const int lane_grp = warp & 3;               // TMEM lanes this warp may touch: 32*(warp%4)..+31
        const int row = lane_grp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        uint32_t ph_acc = 0, ph_agg = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1;
            const int tile = blockIdx.x + i * gridDim.x;
            const int e = tile * kTileM + row;
            const bool valid = e < p.n_edges;
            const int c = valid ? __ldg(p.ctr + e) : -1;
            const int nb = valid ? __ldg(p.nbr + e) : 0;
            uint8_t* a_buf = sA + (size_t)b * kTileBytes;

            // ---- segment bookkeeping + selector matrix (edges of a node are contiguous) ----
            if (half == 0) s_ctr[row] = c;
            epi_bar_sync();
            const int first_node = s_ctr[0];
            {
                const int prev = row > 0 ? s_ctr[row - 1] : -2;
                const int next = row < kTileM - 1 ? s_ctr[row + 1] : -2;
                if (half == 0 && valid) {
                    if (c != prev) s_start[c - first_node] = row;
                    if (c != next) {
                        s_end[c - first_node] = row + 1;
                        if (next < 0) s_misc[0] = c - first_node + 1;
                    }
                }
            }
            epi_bar_sync();
            const int nslots = s_misc[0];
            if (row < nslots) {
                const int st = s_start[row], en = s_end[row];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r0 = half * 64 + j * 8 + q * 2;
                        const uint32_t lo = (r0 >= st && r0 < en) ? 0x3F80u : 0u;
                        const uint32_t hi = (r0 + 1 >= st && r0 + 1 < en) ? 0x3F80u : 0u;
                        w[q] = lo | (hi << 16);
                    }
                    *reinterpret_cast<uint4*>(sSel + half * kKBlockBytes + sw128_chunk(row, j)) =
                        make_uint4(w[0], w[1], w[2], w[3]);
                }
            }

            // ---- chained stage epilogues: TMEM -> (+gather / +bias) -> GELU -> bf16 -> smem A operand ----
            for (int s = 0; s < S; ++s) {
                mbar_wait(&bars->acc_full, ph_acc); ph_acc ^= 1;
                tc_fence_after();
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const int col0 = half * 64 + ch * 32;
                    float v[32];
                    tmem_ld32(t_row + (uint32_t)col0, v);
                    if (s == 0) {
                        if (valid) {
                            const float4* pc = reinterpret_cast<const float4*>(p.Pc + (size_t)c * p.ld_p + col0);
                            const float4* pn = reinterpret_cast<const float4*>(p.Pn + (size_t)nb * p.ld_p + col0);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 x = __ldg(pc + q), y = __ldg(pn + q);
                                v[4 * q + 0] += x.x + y.x; v[4 * q + 1] += x.y + y.y;
                                v[4 * q + 2] += x.z + y.z; v[4 * q + 3] += x.w + y.w;
                            }
                        }
                    } else {
                        const float4* bs = reinterpret_cast<const float4*>(p.bias[s] + col0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 x = __ldg(bs + q);
                            v[4 * q + 0] += x.x; v[4 * q + 1] += x.y; v[4 * q + 2] += x.z; v[4 * q + 3] += x.w;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = gelu_erf(v[q]);
                    if (p.debug && valid) {
                        float* dbg = p.debug + ((size_t)s * p.n_edges + e) * 128 + col0;
#pragma unroll
                        for (int q = 0; q < 32; ++q) dbg[q] = v[q];
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 pk = make_uint4(pack_bf16(v[8 * j + 0], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                    pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
                        *reinterpret_cast<uint4*>(a_buf + half * kKBlockBytes + sw128_chunk(row, ch * 4 + j)) = pk;
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&bars->act_ready);
            }

            // ---- neighbour sums out of TMEM: one row per node slot ----
            mbar_wait(&bars->agg_full, ph_agg); ph_agg ^= 1;
            tc_fence_after();
            if (lane_grp * 32 < nslots) {                   // warp-uniform
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const int col0 = half * 64 + ch * 32;
                    float v[32];
                    tmem_ld32(t_row + (uint32_t)(128 + col0), v);
                    if (row < nslots) {
                        float* dst = p.agg + (size_t)(first_node + row) * 128 + col0;
#pragma unroll
                        for (int q = 0; q < 32; ++q) atomicAdd(dst + q, v[q] * p.alpha);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->agg_empty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// bf16 row-major [rows][128] tensor, box = 64 columns x 128 rows, 128-byte swizzle
static int make_map_bf16_rows128(CUtensorMap* map, const void* base, uint64_t rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not found"); return RMP_E_ARCH; }
    const cuuint64_t dims[2] = {128, rows};
    const cuuint64_t strides[1] = {256};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return RMP_E_ARG; }
    return RMP_OK;
}

}  // namespace tc
}  // namespace rmp

using namespace rmp;

extern "C" int rmp_edge_mlp_bf16(const void* h_E_bf16, const void* W_bf16, int num_stages, const float* const* bias,
                                 const float* Pc, const float* Pn, int ld_p, const int32_t* ctr, const int32_t* nbr,
                                 int n_edges, float alpha, float* agg, float* debug, void* stream) {
    RMP_REQUIRE(h_E_bf16 && W_bf16 && Pc && Pn && ctr && nbr && agg && bias, RMP_E_ARG, "rmp_edge_mlp_bf16: null pointer");
    RMP_REQUIRE(num_stages >= 1 && num_stages <= tc::kMaxStages, RMP_E_SHAPE, "rmp_edge_mlp_bf16: 1..4 message layers");
    RMP_REQUIRE(n_edges >= 0 && ld_p >= 128 && (ld_p & 3) == 0, RMP_E_ARG, "rmp_edge_mlp_bf16: bad sizes");
    RMP_REQUIRE(aligned16(h_E_bf16) && aligned16(W_bf16) && aligned16(Pc) && aligned16(Pn), RMP_E_ALIGN,
                "rmp_edge_mlp_bf16: pointers must be 16-byte aligned");
    if (n_edges == 0) return RMP_OK;
    CUtensorMap tm_e, tm_w;
    int rc = tc::make_map_bf16_rows128(&tm_e, h_E_bf16, (uint64_t)n_edges);
    if (rc) return rc;
    rc = tc::make_map_bf16_rows128(&tm_w, W_bf16, (uint64_t)num_stages * 128);
    if (rc) return rc;
    tc::EdgeParams p{};
    p.Pc = Pc; p.Pn = Pn; p.ld_p = ld_p; p.ctr = ctr; p.nbr = nbr;
    for (int s = 1; s < num_stages; ++s) {
        RMP_REQUIRE(bias[s] && aligned16(bias[s]), RMP_E_ALIGN, "rmp_edge_mlp_bf16: bias[%d] null or unaligned", s);
        p.bias[s] = bias[s];
    }
    p.num_stages = num_stages; p.alpha = alpha; p.agg = agg; p.n_edges = n_edges; p.debug = debug;
    const size_t smem = 1024 + (size_t)(num_stages + 3) * tc::kTileBytes + sizeof(tc::Barriers) + 16 +
                        (3 * tc::kTileM + 4) * sizeof(int);
    RMP_REQUIRE(smem <= 227 * 1024, RMP_E_SHAPE, "rmp_edge_mlp_bf16: %zu B of shared memory needed", smem);
    cudaFuncSetAttribute(tc::edge_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(n_edges, tc::kTileM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    tc::edge_mlp_tc_kernel<<<grid, tc::kThreads, smem, (cudaStream_t)stream>>>(tm_e, tm_w, p);
    return check_launch("rmp_edge_mlp_bf16");
}
