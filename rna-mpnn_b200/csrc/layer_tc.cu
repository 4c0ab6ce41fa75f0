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
// Roles (768 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 4..7 = gather /
// drain (accumulator pre-initialisation, selector matrix, neighbour sums), warps 8..23 = epilogue.
#include "tc_common.cuh"
#include "dropout.cuh"

#include <cstdio>
#include <cstdlib>

namespace rmp {
namespace tc {

struct EdgeParams {
    const float* Pc;                 // [Nv, ld_pc] fp32 centre projection incl. stage-0 bias (L1-resident rows)
    const uint16_t* Pn;              // [Nv, ld_pn] bf16 neighbour projection (random rows: gathered from L2)
    int ld_pc, ld_pn;
    const int32_t* ctr;
    const int32_t* nbr;
    const int32_t* row_ptr;          // [Nv+1] CSR edge offsets (edges of a node are contiguous)
    const float* bias[kMaxStages];   // stages >= 1
    int num_stages;
    float alpha;
    float* agg;                      // [Nv,128] fp32, pre-zeroed
    int n_edges;
    int flags;                       // 1 = first edge tile requested before W2.., 2 = run-ahead GEMMs behind hand-offs
    // training forward (template variant bit 2; rmp_edge_mlp_train_fwd_bf16): Dropout after every GELU with the training
    // path's regenerable mask, and what the backward needs written out: z_s = bf16 pre-activations of every stage, a_s =
    // the activations of the non-final stages IN OPERAND FORM 2 * GELU(z) * mask (the caller scales by keep_scale / 2)
    uint16_t* zsave[kMaxStages]; uint16_t* asave[kMaxStages];
    uint32_t thresh8; float keep_scale; uint64_t seed; const uint64_t* seed_dev; uint32_t stream0;
    float* debug;                    // optional [num_stages][n_edges][128] post-activation dump
    long long* timeline;             // optional int64[32768] clock64 trace of CTA 0 (per epilogue warp + MMA thread; tuning)
};

// Pipeline.  A tile owns slot = i % 3: one 32 KB smem buffer (h_E from TMA; at the end the last activation
// m3), one 128-column TMEM accumulator X[slot] and one 4 KB selector matrix.
//   gather warps : X[slot] <- Pc[centre] + Pn[neighbour] (tcgen05.st, a tile or more ahead), selector matrix of
//                  the tile, and -- once the selector MMA of the slot's previous tile is done -- its neighbour
//                  sums TMEM -> red.global.  They own every global load/store of the kernel.
//   MMA thread   : stage 0  X += h_E . W1e^T          (A, B from smem)
//                  stage s  X  = m_s . W_s^T          (A = the activations in TMEM, Y[tile & 1]; B from smem)
//                  last     X[:, 0:16] = m3^T . Sel^T (A = m3 read MN-major from smem)
//   epilogue     : X -> (+bias) -> GELU -> bf16 -> Y in TMEM (tcgen05.st; no shared-memory round trip, no proxy
//                  fence), the last stage -> smem.  Its only inputs are TMEM and shared memory.
// Schedule of the 16 epilogue warps (thread = row x 32 columns) over the two tiles in flight: stage s of the
// even-tile stream, then a stage of the odd-tile stream, so the GEMM of one tile is hidden behind the epilogue
// of the other.
// Registers: the roles are rebalanced with setmaxnreg (warpgroup 0 = TMA/MMA 40, warpgroup 1 = gather 120; the CTA total stays 768 x 80,
// warpgroups 2..5 = epilogue 80).
constexpr int kSlots = 3;
constexpr int kGatherThreads = 128;
constexpr int kEdgeEpiThreads = 512;
constexpr int kEdgeThreads = 128 + kGatherThreads + kEdgeEpiThreads;        // 768
constexpr int kSelSlots = 16;                                  // node slots per tile served by the selector MMA
constexpr int kSelBytes = 2 * kSelSlots * 128;                 // [2 K-blocks][16 slots][64 edges] bf16
constexpr int kPcStride = 132;                                 // floats per row of the per-tile centre table
constexpr uint32_t kEdgeTmemCols = 512;                        // X[s] at 128*s; Y[t] (bf16 A operand) at 384 + 64*t
constexpr uint32_t kYCol = 384;

struct alignas(8) Barriers {
    uint64_t w_full[kMaxStages], a_full[kSlots], a_empty[kSlots];
    uint64_t init_done[kSlots], acc_full[kSlots], act_ready[kSlots], agg_full[kSlots];
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
        "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]),
        "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]),
        "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// split TMEM load: request 32 columns, and later wait for them (the registers are tied to the wait so that no use
// can be scheduled before it)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 2 * GELU(x), packed (the epilogue's form): y = x + x * tanh(..) saves the 0.5 x multiply of gelu_bf16x2 on the co-limiting
// FMA pipe.  The factor 2 is exact in bf16 and is taken back where the next GEMM's result is read (one packed
// fma(acc, 0.5, bias) instead of the bias add; the neighbour sums with alpha / 2): bit-identical results.
__device__ __forceinline__ uint32_t gelu2x_bf16x2(float a, float b) {
    const __nv_bfloat162 x = __floats2bfloat162_rn(a, b);
    const __nv_bfloat162 c1 = __float2bfloat162_rn(0.0356774081f), c0 = __float2bfloat162_rn(0.7978845608f);
    const __nv_bfloat162 inner = __hmul2(x, __hfma2(__hmul2(x, x), c1, c0));
    uint32_t t;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&inner)));
    const __nv_bfloat162 y = __hfma2(x, *reinterpret_cast<const __nv_bfloat162*>(&t), x);
    return *reinterpret_cast<const uint32_t*>(&y);
}

// D[tmem] (+)= A[tmem] . B[smem]^T: the A operand (128 rows x 16 bf16 = 8 columns) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
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

// 256-bit gather load that does not allocate in L1 (the neighbour rows are used once per tile)
__device__ __forceinline__ void ldg256_stream(const void* ptr, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(ptr));
}
// (a, b) += (x, y): one packed fp32x2 add
__device__ __forceinline__ void add2(float& a, float& b, float x, float y) {
    asm("{\n\t"
        ".reg .b64 va, vx;\n\t"
        "mov.b64 va, {%0, %1};\n\t"
        "mov.b64 vx, {%2, %3};\n\t"
        "add.rn.f32x2 va, va, vx;\n\t"
        "mov.b64 {%0, %1}, va;\n\t"
        "}"
        : "+f"(a), "+f"(b)
        : "f"(x), "f"(y));
}
// 2 * GELU of a packed bf16x2 pre-activation (the training forward also stores that bf16 z)
__device__ __forceinline__ uint32_t gelu2x_of_bf16x2(uint32_t xz) {
    const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&xz);
    const __nv_bfloat162 c1 = __float2bfloat162_rn(0.0356774081f), c0 = __float2bfloat162_rn(0.7978845608f);
    const __nv_bfloat162 inner = __hmul2(x, __hfma2(__hmul2(x, x), c1, c0));
    uint32_t t;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&inner)));
    const __nv_bfloat162 y = __hfma2(x, *reinterpret_cast<const __nv_bfloat162*>(&t), x);
    return *reinterpret_cast<const uint32_t*>(&y);
}
// (a, b) = (a, b) * s + (x, y) with the same scale s in both lanes
__device__ __forceinline__ void fma2_scale(float& a, float& b, float s, float x, float y) {
    asm("{\n\t"
        ".reg .b64 va, vx, vs;\n\t"
        "mov.b64 va, {%0, %1};\n\t"
        "mov.b64 vx, {%2, %3};\n\t"
        "mov.b64 vs, {%4, %4};\n\t"
        "fma.rn.f32x2 va, va, vs, vx;\n\t"
        "mov.b64 {%0, %1}, va;\n\t"
        "}"
        : "+f"(a), "+f"(b)
        : "f"(x), "f"(y), "f"(s));
}
// (a, b) = (a, b) * 0.5 + (x, y): one packed fp32x2 fma
__device__ __forceinline__ void fma2_half(float& a, float& b, float x, float y) {
    asm("{\n\t"
        ".reg .b64 va, vx, vh;\n\t"
        "mov.b64 va, {%0, %1};\n\t"
        "mov.b64 vx, {%2, %3};\n\t"
        "mov.b64 vh, {0f3F000000, 0f3F000000};\n\t"
        "fma.rn.f32x2 va, va, vh, vx;\n\t"
        "mov.b64 {%0, %1}, va;\n\t"
        "}"
        : "+f"(a), "+f"(b)
        : "f"(x), "f"(y));
}
template <int N>
__device__ __forceinline__ void set_max_regs_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void set_max_regs_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void gather_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(kGatherThreads) : "memory"); }

// 16 packed bf16x2 GELU values of 32 fp32 pre-activations
__device__ __forceinline__ void gelu_pack32(const float (&v)[32], uint32_t (&o)[16]) {
#pragma unroll
    for (int q = 0; q < 16; ++q) o[q] = gelu_bf16x2(v[2 * q], v[2 * q + 1]);
}

template <bool TRACE, bool DECOUPLED, int V = 0>
__global__ void __launch_bounds__(kEdgeThreads, 1)
edge_mlp_tc_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_w,
                   const EdgeParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;                                         // [S][2][128][64] bf16, resident
    uint8_t* sA = sW + (size_t)p.num_stages * kTileBytes;       // [3 slots][2][128][64] bf16
    uint8_t* sSel = sA + kSlots * kTileBytes;                   // [3 slots][2][16][64] bf16 selector
    float* s_pc = reinterpret_cast<float*>(sSel + kSlots * kSelBytes);                // [16][kPcStride] centre rows of a tile
    Barriers* bars = reinterpret_cast<Barriers*>(s_pc + kSelSlots * kPcStride);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    // per (generation = (i / 3) & 1, slot): first node, node count of the tile.  Two generations, because the
    // gather warps read the previous tile's entry (drain) while one of them already writes the next tile's
    int* s_meta = reinterpret_cast<int*>(s_tmem + 4);
    float* s_bias = reinterpret_cast<float*>(s_meta + 4 * kSlots + 4);                // [kMaxStages][128]
    s_bias += (4 - ((smem_u32(s_bias) >> 2) & 3)) & 3;                                // 16-byte aligned for LDS.128

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.n_edges + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.num_stages;
    constexpr bool kFold = (V & 1) != 0;            // activations held as 2 * GELU (gelu2x_bf16x2)
    constexpr bool kTrain = (V & 4) != 0;           // training forward: dropout + saved z / a (needs kFold)
    static_assert(!kTrain || (kFold && !DECOUPLED), "the training forward builds on the folded, staggered variant");

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) mbar_init(&bars->w_full[s], 1);
        for (int b = 0; b < kSlots; ++b) {
            mbar_init(&bars->a_full[b], 1); mbar_init(&bars->a_empty[b], 1);
            mbar_init(&bars->init_done[b], kGatherThreads);
            mbar_init(&bars->acc_full[b], 1);
            mbar_init(&bars->act_ready[b], DECOUPLED ? kEdgeEpiThreads / 2 : kEdgeEpiThreads);
            mbar_init(&bars->agg_full[b], 1);
        }
        fence_barrier_init();
    }
    // stage biases (stages >= 1) live in shared memory: broadcast LDS instead of global loads per chunk
    for (int t = threadIdx.x; t < (S - 1) * 128; t += kEdgeThreads) {
        const int st = 1 + t / 128;
        const float* bsp = st == 1 ? p.bias[1] : (st == 2 ? p.bias[2] : p.bias[3]);
        s_bias[st * 128 + (t & 127)] = __ldg(bsp + (t & 127));
    }
    if (warp == 1) tmem_alloc(s_tmem, kEdgeTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp < 4) {
        set_max_regs_dec<40>();
        if (warp == 0) {
            // ===================== TMA producer =====================
            if (elect_one()) {
                tma_prefetch_desc(&tm_e);
                tma_prefetch_desc(&tm_w);
                // one barrier per stage's weights; with flag 1 the first edge tile is requested right behind W1e, so
                // that the first GEMM does not queue behind the 64 KB of later-stage weights of all 148 CTAs
                auto load_w = [&](int s) {
                    mbar_expect_tx(&bars->w_full[s], (uint32_t)kTileBytes);
                    for (int kb = 0; kb < 2; ++kb)
                        tma_load_2d(sW + (size_t)s * kTileBytes + kb * kKBlockBytes, &tm_w, &bars->w_full[s], kb * 64, s * 128);
                };
                const bool w_late = (p.flags & 1) != 0;
                for (int s = 0; s < (w_late ? 1 : S); ++s) load_w(s);
                for (int i = 0; i < my_tiles; ++i) {
                    const int b = i % kSlots;
                    const int tile = blockIdx.x + i * gridDim.x;
                    mbar_wait_relaxed(&bars->a_empty[b], ((i / kSlots) & 1) ^ 1, 200);
                    mbar_expect_tx(&bars->a_full[b], kTileBytes);
                    for (int kb = 0; kb < 2; ++kb)
                        tma_load_2d(sA + (size_t)b * kTileBytes + kb * kKBlockBytes, &tm_e, &bars->a_full[b], kb * 64,
                                    tile * kTileM);
                    if (w_late && i == 0)
                        for (int s = 1; s < S; ++s) load_w(s);
                }
                if (w_late && my_tiles == 0)
                    for (int s = 1; s < S; ++s) load_w(s);
            }
            __syncwarp();
        } else if (warp == 1) {
            // ===================== MMA issuer: serves the tiles in flight as their operands become ready ====
            if (elect_one()) {
                constexpr uint32_t idesc_stage = instr_desc_bf16(128, 128, 0);
                constexpr uint32_t idesc_sel = instr_desc_bf16_amn(128, kSelSlots);
                int ti[kSlots], st[kSlots];
                bool sel_path[kSlots];
#pragma unroll
                for (int b = 0; b < kSlots; ++b) { ti[b] = b; st[b] = 0; sel_path[b] = true; }
                const uint64_t d_a0 = smem_desc_sw128(smem_u32(sA), 16, 1024);              // h_E tile, K-major
                const uint64_t d_am0 = smem_desc_sw128(smem_u32(sA), kKBlockBytes, 1024);   // m3 read MN-major
                const uint64_t d_w0 = smem_desc_sw128(smem_u32(sW), 16, 1024);
                const uint64_t d_sel0 = smem_desc_sw128(smem_u32(sSel), 16, 1024);
                mbar_wait(&bars->w_full[0], 0);
                uint32_t w_seen = 1u;                   // stages whose weights have been waited for
                const bool defer_ahead = (p.flags & 2) != 0;
                int remaining = my_tiles;
                int tlm = 0;
                const bool tlm_on = TRACE && p.timeline != nullptr && blockIdx.x == 0;
                auto mstamp = [&](int tag) {        // MMA-thread trace lives in the upper part of the buffer
                    if constexpr (TRACE) {
                        if (tlm_on && tlm < 1990) { p.timeline[25600 + tlm * 2] = tag; p.timeline[25601 + tlm * 2] = clock64(); ++tlm; }
                    }
                };
                while (remaining > 0) {
                    bool progressed = false;
                    // pass 0: later stages of tiles in flight (the epilogue is waiting for these);
                    // pass 1: stage-0 GEMMs of tiles ahead (nobody waits yet) -- only when pass 0 found nothing
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        if (!defer_ahead) {
                            if (pass == 1 && progressed) break;
                        } else if (pass == 1 && !progressed) {
                            // flag 2: a run-ahead stage-0 GEMM is issued right behind a GEMM somebody waits for (the next
                            // such hand-off is most of an epilogue phase away), or when fewer than two tiles are in their
                            // later stages (start, tail) -- never at a random time in front of a hand-off
                            int late = 0;
#pragma unroll
                            for (int b = 0; b < kSlots; ++b) late += (ti[b] < my_tiles && st[b] > 0) ? 1 : 0;
                            if (late >= 2) break;
                        }
#pragma unroll
                        for (int b = 0; b < kSlots; ++b) {
                            const int i = ti[b], s = st[b];
                            if (i >= my_tiles || (pass == 0) != (s > 0)) continue;
                            const uint32_t par_tile = (uint32_t)((i / kSlots) & 1);
                            bool ready;
                            if (s == 0) ready = mbar_test(&bars->a_full[b], par_tile) && mbar_test(&bars->init_done[b], par_tile);
                            else ready = mbar_test(&bars->act_ready[b], (uint32_t)((S * (i / kSlots) + s - 1) & 1));
                            if (!ready) continue;
                            progressed = true;
                            tc_fence_after();
                            mstamp(1000 * (i + 1) + s);
                            // descriptors are derived from four base descriptors by adding (byte offset >> 4) to the
                            // 14-bit start-address field (keeps the single issuing thread inside its register budget)
                            // (`zero` is opaque to the compiler so that it does not hoist ~50 precomputed descriptors
                            // out of the polling loop and spill them)
                            uint32_t zero;
                            asm volatile("mov.u32 %0, 0;" : "=r"(zero));
                            const uint64_t d_a = d_a0 + (uint64_t)(zero + b * (kTileBytes >> 4));
                            const uint32_t acc = tmem_base + (uint32_t)(b * 128);
                            if (s == 0) {
#pragma unroll
                                for (int k = 0; k < 8; ++k) {               // K = 128 = 8 x UMMA_K(16)
                                    const uint64_t off = (uint64_t)(((k >> 2) * kKBlockBytes + (k & 3) * 32) >> 4);
                                    // accumulates onto the gathered projections the gather warps stored
                                    umma_bf16(acc, d_a + off, d_w0 + zero + off, idesc_stage, 1u);
                                }
                                umma_commit(&bars->acc_full[b]);
                                mstamp(-(1000 * (i + 1) + s));
                                const volatile int* meta = s_meta + 2 * (kSlots * (int)par_tile + b);
                                sel_path[b] = meta[1] <= kSelSlots;
                                st[b] = 1;
                            } else if (s < S) {
                                if (!(w_seen & (1u << s))) { mbar_wait(&bars->w_full[s], 0); w_seen |= 1u << s; }
                                const uint64_t d_w = d_w0 + (uint64_t)(zero + s * (kTileBytes >> 4));
                                const uint32_t a_tmem = tmem_base + kYCol + (uint32_t)((i & 1) * 64);
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const uint64_t off = (uint64_t)(((k >> 2) * kKBlockBytes + (k & 3) * 32) >> 4);
                                    umma_bf16_ts(acc, a_tmem + (uint32_t)(k * 8), d_w + off, idesc_stage, (uint32_t)(k > 0));
                                }
                                umma_commit(&bars->acc_full[b]);
                                mstamp(-(1000 * (i + 1) + s));
                                st[b] = s + 1;
                            } else {
                                // neighbour sums, transposed: aggT[ch][slot] = sum_r m3[r][ch] * Sel[slot][r], written
                                // over the (dead) first 16 accumulator columns.  A = m3 read MN-major from smem,
                                // B = 16-slot selector (K-major).
                                if (sel_path[b]) {
                                    const uint64_t d_am = d_am0 + (uint64_t)(zero + b * (kTileBytes >> 4));
                                    const uint64_t d_sel = d_sel0 + (uint64_t)(zero + b * (kSelBytes >> 4));
#pragma unroll
                                    for (int k = 0; k < 8; ++k) {           // K = 128 edge rows = 8 x 16
                                        const uint64_t a_off = (uint64_t)((k * 2048) >> 4);       // 16 rows of 128 B
                                        const uint64_t b_off = (uint64_t)(((k >> 2) * (kSelSlots * 128) + (k & 3) * 32) >> 4);
                                        umma_bf16(acc, d_am + a_off, d_sel + b_off, idesc_sel, k > 0);
                                    }
                                }
                                umma_commit(&bars->agg_full[b]);
                                umma_commit(&bars->a_empty[b]);
                                mstamp(-(1000 * (i + 1) + s));
                                st[b] = 0;
                                ti[b] = i + kSlots;
                                --remaining;
                            }
                        }
                    }
                    if (!progressed) __nanosleep(20);      // do not steal issue slots from the epilogue warps
                }
                if (tlm_on) p.timeline[32016] = tlm;
            }
            __syncwarp();
        }
    } else if (warp < 8) {
        // ===================== gather / drain warps (thread = edge row, and = channel when draining) =====
        set_max_regs_inc<120>();
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const int gt = threadIdx.x - 128;                          // 0..127
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        // neighbour sums of the tile that used slot b before: TMEM -> red.global (lane = channel, column = node slot)
        auto drain = [&](int b, int j_prev) {
            mbar_wait(&bars->agg_full[b], (uint32_t)((j_prev / kSlots) & 1));
            tc_fence_after();
            const volatile int* meta = s_meta + 2 * (kSlots * ((j_prev / kSlots) & 1) + b);
            const int fn = meta[0], nslots = meta[1];
            // m3 is held as 2 * GELU (training: 2 * GELU * mask, survivors scaled here)
            const float alpha_sum = (kFold ? 0.5f : 1.f) * p.alpha * (kTrain ? p.keep_scale : 1.f);
            if (nslots <= kSelSlots) {
                float v[16];
                tmem_ld16(t_row + (uint32_t)(b * 128), v);
                float* dst = p.agg + (size_t)fn * 128 + row;
#pragma unroll
                for (int n = 0; n < 16; ++n)
                    if (n < nslots)
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + n * 128), "f"(v[n] * alpha_sum) : "memory");
            }
            tc_fence_before();
        };
        auto load_idx = [&](int i, int& c, int& nb, int& c_first, int& c_last) {
            c = -1; nb = 0; c_first = 0; c_last = 0;
            if (i < my_tiles) {
                const int e_base = (blockIdx.x + i * gridDim.x) * kTileM;
                const int e = e_base + row;
                if (e < p.n_edges) { c = __ldg(p.ctr + e); nb = __ldg(p.nbr + e); }
                c_first = __ldg(p.ctr + e_base);
                c_last = __ldg(p.ctr + min(e_base + kTileM, p.n_edges) - 1);
            }
        };
        int c_nx, nb_nx, cf_nx, cl_nx;
        load_idx(0, c_nx, nb_nx, cf_nx, cl_nx);
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i % kSlots;
            const int c = c_nx, nb = nb_nx, fn = cf_nx, nslots = cl_nx - cf_nx + 1;
            const bool table = nslots <= kSelSlots;            // else: tiny structures, > 16 nodes in the tile
            // every load of the tile is requested up front: the whole neighbour row (random, L2) ...
            const uint16_t* pn = p.Pn + (size_t)nb * p.ld_pn;
            uint32_t w[8][8];
#pragma unroll
            for (int q = 0; q < 8; ++q) ldg256_stream(pn + 16 * q, w[q]);
            // ... and this thread's 16 floats of the tile's <= 16 distinct centre rows (coalesced)
            uint32_t pcf[2][8];
            const bool fill = table && (gt >> 3) < nslots;
            if (fill) {
                const float* src = p.Pc + (size_t)(fn + (gt >> 3)) * p.ld_pc + (gt & 7) * 16;
                ldg256(src, pcf[0]); ldg256(src + 8, pcf[1]);
            }
            load_idx(i + 1, c_nx, nb_nx, cf_nx, cl_nx);        // streaming index loads of the next tile
            if (i >= kSlots) drain(b, i - kSlots);             // also: X[b], sel[b], s_meta[gen][b] are free from here on
            gather_bar_sync();                                 // everybody is done with the previous tile's centre table
            if (fill) {
                float4* dst = reinterpret_cast<float4*>(s_pc + (gt >> 3) * kPcStride + (gt & 7) * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    dst[q] = make_float4(__uint_as_float(pcf[q >> 1][(q & 1) * 4 + 0]), __uint_as_float(pcf[q >> 1][(q & 1) * 4 + 1]),
                                         __uint_as_float(pcf[q >> 1][(q & 1) * 4 + 2]), __uint_as_float(pcf[q >> 1][(q & 1) * 4 + 3]));
            }
            if (row == 0) {
                int* meta = s_meta + 2 * (kSlots * ((i / kSlots) & 1) + b);
                meta[0] = fn; meta[1] = nslots;
            }
            const int my_slot = c >= 0 ? c - fn : -1;
            if (table) {
                // column `row` of the selector matrix Sel[slot][edge] (K-major in the edge index, 128B swizzle)
                uint8_t* sel = sSel + b * kSelBytes + (row >> 6) * (kSelSlots * 128) + (row & 7) * 2;
#pragma unroll
                for (int sl = 0; sl < kSelSlots; ++sl)
                    *reinterpret_cast<uint16_t*>(sel + sw128_chunk(sl, (row & 63) >> 3)) = sl == my_slot ? 0x3F80 : 0;
            }
            gather_bar_sync();                                 // centre table complete
            const float* pc_row = s_pc + max(my_slot, 0) * kPcStride;
            const float* pc_glob = p.Pc + (size_t)max(c, 0) * p.ld_pc;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {                   // 32 columns at a time
                float v[32];
                if (table) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 x = reinterpret_cast<const float4*>(pc_row + q4 * 32)[q];
                        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = __ldg(pc_glob + q4 * 32 + q);
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {                 // bf16 pair q of this 32-column chunk
                    const uint32_t ww = w[q4 * 2 + (q >> 3)][q & 7];
                    v[2 * q] += __uint_as_float(ww << 16);
                    v[2 * q + 1] += __uint_as_float(ww & 0xffff0000u);
                }
                if (c < 0) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = 0.f;
                }
                tmem_st32(t_row + (uint32_t)(b * 128 + q4 * 32), v);
            }
            tmem_st_wait();
            fence_proxy_async();                               // selector bytes -> visible to the tensor core
            tc_fence_before();
            mbar_arrive(&bars->init_done[b]);
        }
        for (int i = my_tiles; i < my_tiles + kSlots; ++i)     // neighbour sums of the last tiles
            if (i >= kSlots) drain(i % kSlots, i - kSlots);
    } else {
        // ===================== epilogue warps: thread = one edge row x 32 columns ==========
        // Two tile streams (even / odd tiles), one stage of each alternately, the odd stream lagging by half a
        // tile: a stream's next GEMM runs while the other stream's epilogue is on the CUDA cores, tile starts are
        // evenly spaced, and a finished tile's slot has four phases to be drained, re-initialised and refilled by
        // TMA before its next owner asks for it.
        const int ew = warp - 8;
        const int lane_grp = warp & 3;                   // TMEM lanes this warp may touch: 32*(warp%4)..+31
        const int row = lane_grp * 32 + lane;
        constexpr int kChunks = DECOUPLED ? 2 : 1;       // 32-column chunks per thread and phase
        const int colb = DECOUPLED ? ((ew >> 2) & 1) * 64 : (ew >> 2) * 32;     // this thread's first column
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        int tl_n = 0;
        const bool tl_on = TRACE && p.timeline != nullptr && blockIdx.x == 0 && lane == 0;
        long long* tl_buf = p.timeline + ew * 1600;
        auto stamp = [&](int tag) {
            if constexpr (TRACE) {
                if (tl_on && tl_n < 799) { tl_buf[tl_n * 2] = tag; tl_buf[tl_n * 2 + 1] = clock64(); ++tl_n; }
            }
        };

        [[maybe_unused]] const float half_ks = kTrain ? 0.5f * p.keep_scale : 0.5f;
        [[maybe_unused]] const uint64_t seed_tr = kTrain ? p.seed + (p.seed_dev ? __ldg(p.seed_dev) : 0ull) : 0ull;
        // phase sequence: the even-tile stream at (pair k, stage sa), then the odd-tile stream S/2 stages behind it
        const int off = S >> 1;
        int seq_k = 0, seq_s = 0, seq_half = 0;
        auto advance = [&](int& i, int& st) -> bool {
            if constexpr (DECOUPLED) {                       // group g = ew / 8 owns the tiles i % 2 == g, stage by stage
                if (seq_half == 0) { seq_half = 1; seq_k = ew >> 3; seq_s = 0; }
                else if (++seq_s == S) { seq_s = 0; seq_k += 2; }
                i = seq_k; st = seq_s;
                return seq_k < my_tiles;
            }
            while (2 * seq_k < my_tiles + 2) {               // one extra pair index lets the lagging odd stream finish
                if (seq_half == 0) {
                    seq_half = 1;
                    if (2 * seq_k < my_tiles) { i = 2 * seq_k; st = seq_s; return true; }
                } else {
                    seq_half = 0;
                    const int kb = seq_s >= off ? seq_k : seq_k - 1;
                    const int sb = seq_s >= off ? seq_s - off : seq_s - off + S;
                    if (++seq_s == S) { seq_s = 0; ++seq_k; }
                    if (kb >= 0 && 2 * kb + 1 < my_tiles) { i = 2 * kb + 1; st = sb; return true; }
                }
            }
            return false;
        };
        int i = 0, st = 0;
        while (advance(i, st)) {
            const int b = i % kSlots;
            const int e_base = (blockIdx.x + i * gridDim.x) * kTileM;
            const bool valid = e_base + row < p.n_edges;
            const bool last = st == S - 1;
            stamp(100 + st * 10 + (i & 1));                    // about to wait for the accumulator
            mbar_wait(&bars->acc_full[b], (uint32_t)((S * (i / kSlots) + st) & 1));
            tc_fence_after();
            stamp(200 + st * 10 + (i & 1));                    // accumulator ready
            const bool bypass = last && reinterpret_cast<const volatile int*>(s_meta)[2 * (kSlots * ((i / kSlots) & 1) + b) + 1] > kSelSlots;
            [[maybe_unused]] uint32_t o_keep[16], xz_keep[16];       // training forward: this phase's z and activation, stored below
#pragma unroll
            for (int ch = 0; ch < kChunks; ++ch) {
                const int col0 = colb + ch * 32;
                float v[32];
                tmem_ld32(t_row + (uint32_t)(b * 128 + col0), v);
                stamp(300 + st * 10 + (i & 1));
                if (st > 0) {
                    const float4* bs = reinterpret_cast<const float4*>(s_bias + st * 128 + col0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 x = bs[q];
                        if constexpr (kTrain) {                // operand was 2 * GELU * mask: v = acc * keep_scale / 2 + bias
                            fma2_scale(v[4 * q + 0], v[4 * q + 1], half_ks, x.x, x.y);
                            fma2_scale(v[4 * q + 2], v[4 * q + 3], half_ks, x.z, x.w);
                        } else if constexpr (kFold) {          // the previous stage wrote 2 * GELU: v = acc / 2 + bias
                            fma2_half(v[4 * q + 0], v[4 * q + 1], x.x, x.y);
                            fma2_half(v[4 * q + 2], v[4 * q + 3], x.z, x.w);
                        } else {
                            add2(v[4 * q + 0], v[4 * q + 1], x.x, x.y);
                            add2(v[4 * q + 2], v[4 * q + 3], x.z, x.w);
                        }
                    }
                }
                uint32_t o[16];
                [[maybe_unused]] uint32_t xz[16];
                if constexpr (kTrain) {
                    // z (bf16) is what the backward sees; the activation is evaluated at exactly that value
#pragma unroll
                    for (int q = 0; q < 16; ++q) { xz[q] = pack_bf16(v[2 * q], v[2 * q + 1]); o[q] = gelu2x_of_bf16x2(xz[q]); }
                    if (p.thresh8 != 0) {
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            uint32_t m[4];
                            keep_masks_bf16x2(seed_tr, (uint64_t)(e_base + row) * 16 + (uint64_t)(col0 / 8 + g8),
                                              p.stream0 + (uint32_t)st, p.thresh8, m);
#pragma unroll
                            for (int j = 0; j < 4; ++j) o[4 * g8 + j] &= m[j];
                        }
                    }
                    if (bypass && valid) {                     // > 16 nodes in the tile: vector reductions instead of the selector
                        float* dst = p.agg + (size_t)__ldg(p.ctr + e_base + row) * 128 + col0;
                        const float sc = 0.5f * p.alpha * p.keep_scale;
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            red_add_v4(dst + 4 * q, __uint_as_float(o[2 * q] << 16) * sc, __uint_as_float(o[2 * q] & 0xffff0000u) * sc,
                                       __uint_as_float(o[2 * q + 1] << 16) * sc, __uint_as_float(o[2 * q + 1] & 0xffff0000u) * sc);
                    }
                } else if (p.debug != nullptr || bypass) {
                    // fp32 GELU: debug dump (tests) and the tiny-degree bypass that sums with vector reductions
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = gelu_tanh_fast(v[q]);
                    if (p.debug && valid) {
                        float* dbg = p.debug + ((size_t)st * p.n_edges + e_base + row) * 128 + col0;
#pragma unroll
                        for (int q = 0; q < 32; ++q) dbg[q] = v[q];
                    }
                    if (bypass && valid) {
                        float* dst = p.agg + (size_t)__ldg(p.ctr + e_base + row) * 128 + col0;
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            red_add_v4(dst + 4 * q, v[4 * q] * p.alpha, v[4 * q + 1] * p.alpha,
                                       v[4 * q + 2] * p.alpha, v[4 * q + 3] * p.alpha);
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        o[q] = kFold ? pack_bf16(2.f * v[2 * q], 2.f * v[2 * q + 1]) : pack_bf16(v[2 * q], v[2 * q + 1]);
                } else {
                    if constexpr (kFold) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) o[q] = gelu2x_bf16x2(v[2 * q], v[2 * q + 1]);
                    } else {
                        gelu_pack32(v, o);                         // packed bf16x2 GELU
                    }
                }
                stamp(500 + st * 10 + (i & 1));
                if (last) {
                    if (!bypass) {                                 // m3 -> smem: A operand (MN-major) of the selector MMA
                        uint8_t* kb = sA + (size_t)b * kTileBytes + (col0 >> 6) * kKBlockBytes;
                        const int j0 = (col0 & 63) >> 3;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<uint4*>(kb + sw128_chunk(row, j0 + j)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    }
                } else {
                    tmem_st16(t_row + kYCol + (uint32_t)((i & 1) * 64 + col0 / 2), o);     // next stage's A operand, in TMEM
                }
                if constexpr (kTrain) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) { o_keep[q] = o[q]; xz_keep[q] = xz[q]; }
                }
            }
            if (last) fence_proxy_async();
            else tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->act_ready[b]);
            stamp(700 + st * 10 + (i & 1));
            if constexpr (kTrain) {
                // saved tensors, written AFTER the hand-off (an mbarrier arrive has release semantics and would wait for
                // these stores): this row's 32 columns = 64 contiguous bytes per tensor, two full sectors
                if (valid) {
                    const size_t at = (size_t)(e_base + row) * 128 + colb;
                    uint16_t* zs = st == 0 ? p.zsave[0] : (st == 1 ? p.zsave[1] : (st == 2 ? p.zsave[2] : p.zsave[3]));
                    stg256u(zs + at, &xz_keep[0]); stg256u(zs + at + 16, &xz_keep[8]);
                    if (!last) {
                        uint16_t* as = st == 0 ? p.asave[0] : (st == 1 ? p.asave[1] : p.asave[2]);
                        stg256u(as + at, &o_keep[0]); stg256u(as + at + 16, &o_keep[8]);
                    }
                }
            }
        }
        if (tl_on) p.timeline[32000 + ew] = tl_n;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kEdgeTmemCols);
}

// =================================================================================================
// Node kernel: per tile of 128 nodes
//     x1 = LayerNorm1(h_V + agg)                       (prologue, also re-zeroes agg for the next layer)
//     d  = W_last(GELU(.. GELU(W_a x1 + b_a) ..)) + b  (tcgen05 chain, weights streamed by TMA)
//     h' = LayerNorm2(x1 + d)                          -> h_out (fp32)
//     P  = [W0c | W0n] h' + [b0 | 0]                   -> next layer's per-node projections
// (rdesign/model/mpnn.py:35-37 and the h_V[E_idx] operand of rdesign/model/rdesign.py:85 for the
// next layer).  Stage kinds: 0 = bias+GELU -> bf16 A operand, 1 = bias+residual+LN -> h_out (+A),
// 2 = bias -> fp32 store (projection).  Hidden width <= 256 columns of TMEM.
// =================================================================================================
constexpr int kNodeMaxStages = 5;
constexpr int kRing = 4;                          // weight ring slots of 32 KB ([256 rows][64 k])
constexpr int kSlotBytes = 256 * 128;
// 16 prologue / epilogue warps: thread = row x quarter of the stage's columns.  With 8 warps (two per SM
// sub-partition) the tcgen05.ld -> GELU -> st.shared chain of a 256-wide stage ran 128 columns deep per thread and its
// latency was exposed (4.3-7.3 k cycles per stage, profiles/r1_node_timeline.txt); four warps per sub-partition halve
// the per-thread chain and cover each other's TMEM / L2 latencies.
constexpr int kNodeParts = 4;
constexpr int kNodeEpiThreads = 128 * kNodeParts;            // 512
constexpr int kNodeThreads = 64 + kNodeEpiThreads;           // 576
__device__ __forceinline__ void node_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kNodeEpiThreads) : "memory"); }

struct NodeStage {
    int K, N, kind;
    const float* bias;
};
struct NodeParams {
    const float* h_in;          // [n_rows,128]
    float* agg;                 // [n_rows,128] neighbour sums (already scaled); zeroed after reading; may be NULL
    int prologue_ln;            // 1: x1 = LN1(h_in + agg), 0: x1 = h_in
    const float* ln1_g; const float* ln1_b;
    const float* ln2_g; const float* ln2_b;
    float* x1;                  // [n_rows,128] fp32 scratch (residual), needed when a kind-1 stage exists
    float* h_out;               // [n_rows,128]
    float* p_out;               // kind-2 stage: columns [0,128) fp32 [n_rows,128] (centre projection)
    uint16_t* pn_out;           //               columns [128,256) bf16 [n_rows,128] (neighbour projection)
    NodeStage st[kNodeMaxStages];
    int S;
    int n_rows;
    long long* timeline;        // optional clock64 trace of CTA 0's first epilogue thread
};
struct alignas(8) NodeBarriers {
    uint64_t w_full[kRing], w_empty[kRing], acc_full, act_ready;
};

__global__ void __launch_bounds__(kNodeThreads, 1)
node_mlp_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                   const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3,
                   const __grid_constant__ CUtensorMap tm4, const NodeParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                                  // [4 K-blocks][128][64] bf16 activation (in place)
    uint8_t* sW = sA + 4 * kKBlockBytes;                 // weight ring
    NodeBarriers* bars = reinterpret_cast<NodeBarriers*>(sW + kRing * kSlotBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    float* s_red = reinterpret_cast<float*>(s_tmem + 4);  // [2 stats][4 column parts][128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.n_rows + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.S;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRing; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->act_ready, kNodeEpiThreads);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const CUtensorMap* maps[kNodeMaxStages] = {&tm0, &tm1, &tm2, &tm3, &tm4};

    if (warp == 0) {
        // ===================== TMA producer: weight slabs [N rows][64 k] through the ring =====================
        if (elect_one()) {
            for (int s = 0; s < S; ++s) tma_prefetch_desc(maps[s]);
            uint32_t n = 0;
            for (int i = 0; i < my_tiles; ++i)
                for (int s = 0; s < S; ++s)
                    for (int kb = 0; kb < p.st[s].K / 64; ++kb, ++n) {
                        const uint32_t slot = n % kRing;
                        mbar_wait(&bars->w_empty[slot], ((n / kRing) & 1) ^ 1);
                        mbar_expect_tx(&bars->w_full[slot], (uint32_t)(p.st[s].N * 128));
                        tma_load_2d(sW + slot * kSlotBytes, maps[s], &bars->w_full[slot], kb * 64, 0);
                    }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            uint32_t n = 0, ph_act = 0;
            const uint32_t a_addr = smem_u32(sA);
            for (int i = 0; i < my_tiles; ++i)
                for (int s = 0; s < S; ++s) {
                    const uint32_t idesc = instr_desc_bf16(128, p.st[s].N, 0);
                    mbar_wait(&bars->act_ready, ph_act); ph_act ^= 1;
                    tc_fence_after();
                    for (int kb = 0; kb < p.st[s].K / 64; ++kb, ++n) {
                        const uint32_t slot = n % kRing;
                        mbar_wait(&bars->w_full[slot], (n / kRing) & 1);
                        tc_fence_after();
                        const uint32_t w_addr = smem_u32(sW + slot * kSlotBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base, smem_desc_sw128(a_addr + kb * kKBlockBytes + k * 32, 16, 1024),
                                      smem_desc_sw128(w_addr + k * 32, 16, 1024), idesc, (kb | k) > 0);
                        umma_commit(&bars->w_empty[slot]);
                    }
                    umma_commit(&bars->acc_full);
                }
        }
        __syncwarp();
    } else {
        // ===================== prologue / epilogue warps =====================
        const int part = (warp - 2) >> 2;           // quarter of the stage's columns
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        float* red_sum = s_red;                            // [4][128]
        float* red_sq = s_red + kNodeParts * kTileM;       // [4][128]
        uint32_t ph_acc = 0;
        int tl_n = 0;
        const bool tl_on = p.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
        auto stamp = [&](int tag) {
            if (tl_on && tl_n < 400) { p.timeline[tl_n * 2] = tag; p.timeline[tl_n * 2 + 1] = clock64(); ++tl_n; }
        };
        stamp(1);

        // LayerNorm over the 128 channels of a row held as 4 x 32 by the four column parts
        auto layer_norm32 = [&](float (&x)[32], const float* gamma, const float* beta, int c0) {
            float sm = 0.f;
#pragma unroll
            for (int q = 0; q < 32; ++q) sm += x[q];
            red_sum[part * kTileM + row] = sm;
            node_bar_sync();
            const float mu = ((red_sum[row] + red_sum[kTileM + row]) + (red_sum[2 * kTileM + row] + red_sum[3 * kTileM + row])) *
                             (1.f / 128.f);
            float sq = 0.f;
#pragma unroll
            for (int q = 0; q < 32; ++q) { x[q] -= mu; sq = fmaf(x[q], x[q], sq); }
            red_sq[part * kTileM + row] = sq;
            node_bar_sync();
            const float var = ((red_sq[row] + red_sq[kTileM + row]) + (red_sq[2 * kTileM + row] + red_sq[3 * kTileM + row])) *
                              (1.f / 128.f);
            const float inv = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0) + q);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c0) + q);
                x[4 * q + 0] = fmaf(x[4 * q + 0] * inv, g.x, b.x); x[4 * q + 1] = fmaf(x[4 * q + 1] * inv, g.y, b.y);
                x[4 * q + 2] = fmaf(x[4 * q + 2] * inv, g.z, b.z); x[4 * q + 3] = fmaf(x[4 * q + 3] * inv, g.w, b.w);
            }
        };

        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int r = tile * kTileM + row;
            const bool valid = r < p.n_rows;
            // ---- prologue: x1 = LN1(h + agg) (or h).  One warp per row, lanes across the 128 channels: fully
            //      coalesced 512-byte row accesses, LayerNorm by warp shuffles, bf16 copy -> swizzled A operand.
            //      16 warps x 8 rows: every load of the tile is in flight at once ----
            {
                const int ew = warp - 2;
                float4 g1 = make_float4(1.f, 1.f, 1.f, 1.f), b1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.prologue_ln) {
                    g1 = __ldg(reinterpret_cast<const float4*>(p.ln1_g) + lane);
                    b1 = __ldg(reinterpret_cast<const float4*>(p.ln1_b) + lane);
                }
                float4 xs[8], as[8];
                const bool with_agg = p.prologue_ln && p.agg;
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    const int rg = tile * kTileM + ew * 8 + rr;
                    xs[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                    as[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rg < p.n_rows) {
                        xs[rr] = __ldg(reinterpret_cast<const float4*>(p.h_in + (size_t)rg * 128) + lane);
                        if (with_agg) as[rr] = reinterpret_cast<const float4*>(p.agg + (size_t)rg * 128)[lane];
                    }
                }
                stamp(100);
                if (with_agg) {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int rg = tile * kTileM + ew * 8 + rr;
                        xs[rr].x += as[rr].x; xs[rr].y += as[rr].y; xs[rr].z += as[rr].z; xs[rr].w += as[rr].w;
                        if (rg < p.n_rows)                         // re-arm the neighbour-sum buffer
                            reinterpret_cast<float4*>(p.agg + (size_t)rg * 128)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                stamp(110);
                if (p.prologue_ln) {
                    // LayerNorm of the 8 rows at once: the 8 butterfly reductions of a step are independent
                    // shuffles, so their latencies overlap instead of adding up row after row
                    float sm[8], sq[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) sm[rr] = (xs[rr].x + xs[rr].y) + (xs[rr].z + xs[rr].w);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) sm[rr] += __shfl_xor_sync(0xffffffffu, sm[rr], o);
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const float mu = sm[rr] * (1.f / 128.f);
                        xs[rr].x -= mu; xs[rr].y -= mu; xs[rr].z -= mu; xs[rr].w -= mu;
                        sq[rr] = (xs[rr].x * xs[rr].x + xs[rr].y * xs[rr].y) + (xs[rr].z * xs[rr].z + xs[rr].w * xs[rr].w);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) sq[rr] += __shfl_xor_sync(0xffffffffu, sq[rr], o);
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int rg = tile * kTileM + ew * 8 + rr;
                        const float inv = 1.f / sqrtf(sq[rr] * (1.f / 128.f) + 1e-5f);
                        float4& x = xs[rr];
                        x.x = fmaf(x.x * inv, g1.x, b1.x); x.y = fmaf(x.y * inv, g1.y, b1.y);
                        x.z = fmaf(x.z * inv, g1.z, b1.z); x.w = fmaf(x.w * inv, g1.w, b1.w);
                        if (rg < p.n_rows && p.x1) reinterpret_cast<float4*>(p.x1 + (size_t)rg * 128)[lane] = x;
                    }
                }
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    const int rl = ew * 8 + rr;                    // row inside the tile
                    const float4 x = xs[rr];
                    // columns 4*lane .. 4*lane+3 -> K-block lane/16, 16-byte chunk (lane%16)/2, 8-byte half lane%2
                    *reinterpret_cast<uint2*>(sA + (lane >> 4) * kKBlockBytes + sw128_chunk(rl, (lane & 15) >> 1) + (lane & 1) * 8) =
                        make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
                }
                stamp(3);
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&bars->act_ready);
                node_bar_sync();                                       // x1 rows written above are re-read by other warps
                stamp(4);
            }
            // ---- stage epilogues ----
            for (int s = 0; s < S; ++s) {
                const NodeStage stg = p.st[s];
                const int ncols = stg.N >> 2;                 // columns of this part: 32 (N = 128) or 64 (N = 256)
                const int cb = part * ncols;
                if (stg.kind == 1) {                          // N == 128: bias + residual + LN2 -> h_out (+ operand)
                    float x[32];
                    // residual x1 (written by this CTA's prologue, L2): requested before the accumulator wait
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float rsd[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (valid) ldg256f(p.x1 + (size_t)r * 128 + cb + 8 * q, rsd);
#pragma unroll
                        for (int k = 0; k < 8; ++k) x[8 * q + k] = rsd[k];
                    }
                    mbar_wait(&bars->acc_full, ph_acc); ph_acc ^= 1;
                    tc_fence_after();
                    stamp(10 + s);
                    {
                        float v[32];
                        tmem_ld32(t_row + (uint32_t)cb, v);
#pragma unroll
                        for (int q = 0; q < 32; ++q) x[q] += v[q];
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 ba = __ldg(reinterpret_cast<const float4*>(stg.bias + cb) + q);
                        x[4 * q + 0] += ba.x; x[4 * q + 1] += ba.y; x[4 * q + 2] += ba.z; x[4 * q + 3] += ba.w;
                    }
                    layer_norm32(x, p.ln2_g, p.ln2_b, cb);
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) stg256f(p.h_out + (size_t)r * 128 + cb + 8 * q, &x[8 * q]);
                    }
                    if (s + 1 < S) store_act_chunk(sA, row, cb, x);
                } else {
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
                        if (stg.kind == 0) {
                            store_gelu_chunk(sA, row, col0, v);
                        } else if (valid) {                   // kind 2: projections, centre fp32 | neighbour bf16
                            if (col0 < 128) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) stg256f(p.p_out + (size_t)r * 128 + col0 + 8 * q, &v[8 * q]);
                            } else {
                                uint32_t pk[16];
#pragma unroll
                                for (int q = 0; q < 16; ++q) pk[q] = pack_bf16(v[2 * q], v[2 * q + 1]);
                                stg256u(p.pn_out + (size_t)r * 128 + (col0 - 128), &pk[0]);
                                stg256u(p.pn_out + (size_t)r * 128 + (col0 - 128) + 16, &pk[8]);
                            }
                        }
                    }
                }
                tc_fence_before();
                stamp(20 + s);
                if (s + 1 < S) {
                    fence_proxy_async();
                    mbar_arrive(&bars->act_ready);
                } else {
                    node_bar_sync();         // every TMEM read of this tile is done before the next prologue arrives
                }
            }
        }
        if (tl_on) p.timeline[799] = tl_n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace rmp

using namespace rmp;

extern "C" int rmp_edge_mlp_bf16_traced(const void* h_E_bf16, const void* W_bf16, int num_stages,
                                        const float* const* bias, const float* Pc, const void* Pn_bf16,
                                        const int32_t* ctr, const int32_t* nbr, const int32_t* row_ptr, int n_edges,
                                        float alpha, float* agg, float* debug, long long* timeline, void* stream);

// tuning hooks (scripts/edge_variants.py): epilogue variant (template) and producer / issuer flags of the edge kernel
static int g_edge_variant = 1, g_edge_flags = 3;
static bool g_edge_tuning_set = false;
extern "C" void rmp_set_edge_tuning(int variant, int flags) {
    g_edge_variant = variant; g_edge_flags = flags; g_edge_tuning_set = true;
}

extern "C" int rmp_edge_mlp_bf16(const void* h_E_bf16, const void* W_bf16, int num_stages, const float* const* bias,
                                 const float* Pc, const void* Pn_bf16, const int32_t* ctr, const int32_t* nbr,
                                 const int32_t* row_ptr, int n_edges, float alpha, float* agg, float* debug,
                                 void* stream) {
    return rmp_edge_mlp_bf16_traced(h_E_bf16, W_bf16, num_stages, bias, Pc, Pn_bf16, ctr, nbr, row_ptr, n_edges, alpha,
                                    agg, debug, nullptr, stream);
}

struct EdgeTrainArgs {
    void* const* z_save; void* const* a_save; float p_drop; uint64_t seed; const uint64_t* seed_dev; uint32_t stream0;
};
static int edge_mlp_launch(const void* h_E_bf16, const void* W_bf16, int num_stages, const float* const* bias, const float* Pc,
                           const void* Pn_bf16, const int32_t* ctr, const int32_t* nbr, const int32_t* row_ptr, int n_edges,
                           float alpha, float* agg, float* debug, long long* timeline, const EdgeTrainArgs* train, void* stream);

extern "C" int rmp_edge_mlp_bf16_traced(const void* h_E_bf16, const void* W_bf16, int num_stages,
                                        const float* const* bias, const float* Pc, const void* Pn_bf16,
                                        const int32_t* ctr, const int32_t* nbr, const int32_t* row_ptr, int n_edges,
                                        float alpha, float* agg, float* debug, long long* timeline, void* stream) {
    return edge_mlp_launch(h_E_bf16, W_bf16, num_stages, bias, Pc, Pn_bf16, ctr, nbr, row_ptr, n_edges, alpha, agg, debug, timeline,
                           nullptr, stream);
}

extern "C" int rmp_edge_mlp_train_fwd_bf16(const void* h_E_bf16, const void* W_bf16, int num_stages, const float* const* bias,
                                           const float* Pc, const void* Pn_bf16, const int32_t* ctr, const int32_t* nbr,
                                           const int32_t* row_ptr, int n_edges, float alpha, float* agg, void* const* z_save,
                                           void* const* a_save, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                                           uint32_t stream_id0, void* stream) {
    RMP_REQUIRE(z_save && a_save, RMP_E_ARG, "rmp_edge_mlp_train_fwd_bf16: z_save / a_save are null");
    RMP_REQUIRE(p_drop >= 0.f && p_drop < 1.f, RMP_E_ARG, "rmp_edge_mlp_train_fwd_bf16: dropout probability in [0,1)");
    for (int s = 0; s < num_stages && s < tc::kMaxStages; ++s) {
        RMP_REQUIRE(z_save[s] && aligned32(z_save[s]), RMP_E_ALIGN, "rmp_edge_mlp_train_fwd_bf16: z_save[%d] null or not 32-byte aligned", s);
        RMP_REQUIRE(s == num_stages - 1 || (a_save[s] && aligned32(a_save[s])), RMP_E_ALIGN,
                    "rmp_edge_mlp_train_fwd_bf16: a_save[%d] null or not 32-byte aligned", s);
    }
    const EdgeTrainArgs t{z_save, a_save, p_drop, seed, seed_dev, stream_id0};
    return edge_mlp_launch(h_E_bf16, W_bf16, num_stages, bias, Pc, Pn_bf16, ctr, nbr, row_ptr, n_edges, alpha, agg, nullptr, nullptr,
                           &t, stream);
}

static int edge_mlp_launch(const void* h_E_bf16, const void* W_bf16, int num_stages, const float* const* bias, const float* Pc,
                           const void* Pn_bf16, const int32_t* ctr, const int32_t* nbr, const int32_t* row_ptr, int n_edges,
                           float alpha, float* agg, float* debug, long long* timeline, const EdgeTrainArgs* train, void* stream) {
    const void* Pn = Pn_bf16;
    RMP_REQUIRE(row_ptr, RMP_E_ARG, "rmp_edge_mlp_bf16: row_ptr is null");
    RMP_REQUIRE(h_E_bf16 && W_bf16 && Pc && Pn && ctr && nbr && agg && bias, RMP_E_ARG, "rmp_edge_mlp_bf16: null pointer");
    RMP_REQUIRE(num_stages >= 1 && num_stages <= tc::kMaxStages, RMP_E_SHAPE, "rmp_edge_mlp_bf16: 1..4 message layers");
    RMP_REQUIRE(n_edges >= 0, RMP_E_ARG, "rmp_edge_mlp_bf16: bad sizes");
    RMP_REQUIRE(aligned16(h_E_bf16) && aligned16(W_bf16) && aligned32(Pc) && aligned32(Pn), RMP_E_ALIGN,
                "rmp_edge_mlp_bf16: h_E / W must be 16-byte aligned, Pc / Pn 32-byte aligned (256-bit gathers)");
    if (n_edges == 0) return RMP_OK;
    CUtensorMap tm_e, tm_w;
    int rc = tc::make_map_bf16_rows128(&tm_e, h_E_bf16, (uint64_t)n_edges);
    if (rc) return rc;
    rc = tc::make_map_bf16_rows128(&tm_w, W_bf16, (uint64_t)num_stages * 128);
    if (rc) return rc;
    tc::EdgeParams p{};
    p.Pc = Pc; p.Pn = reinterpret_cast<const uint16_t*>(Pn); p.ld_pc = 128; p.ld_pn = 128; p.ctr = ctr; p.nbr = nbr;
    p.row_ptr = row_ptr;
    for (int s = 1; s < num_stages; ++s) {
        RMP_REQUIRE(bias[s] && aligned16(bias[s]), RMP_E_ALIGN, "rmp_edge_mlp_bf16: bias[%d] null or unaligned", s);
        p.bias[s] = bias[s];
    }
    p.num_stages = num_stages; p.alpha = alpha; p.agg = agg; p.n_edges = n_edges; p.debug = debug;
    p.timeline = timeline;
    const size_t smem = 1024 + (size_t)(num_stages + tc::kSlots) * tc::kTileBytes + tc::kSlots * tc::kSelBytes +
                        tc::kSelSlots * tc::kPcStride * sizeof(float) + sizeof(tc::Barriers) + 16 +
                        (4 * tc::kSlots + 4) * sizeof(int) + tc::kMaxStages * 128 * sizeof(float) + 16;
    RMP_REQUIRE(smem <= 227 * 1024, RMP_E_SHAPE, "rmp_edge_mlp_bf16: %zu B of shared memory needed", smem);
    const int n_tiles = ceil_div(n_edges, tc::kTileM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    static const bool decoupled = []() { const char* e = getenv("RMP_EDGE_SCHEDULE"); return e && e[0] == 'd'; }();
    auto launch = [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kernel<<<grid, tc::kEdgeThreads, smem, (cudaStream_t)stream>>>(tm_e, tm_w, p);
    };
    if (!g_edge_tuning_set) {                        // RMP_EDGE_TUNING="<variant>,<flags>" (tuning aid)
        g_edge_tuning_set = true;
        if (const char* e = getenv("RMP_EDGE_TUNING")) sscanf(e, "%d,%d", &g_edge_variant, &g_edge_flags);
    }
    p.flags = g_edge_flags;
    const int v = g_edge_variant & 1;
    if (train) {
        // same 8-bit quantisation of the drop probability as the stand-alone activation kernels (train.cu)
        const uint32_t thresh = train->p_drop <= 0.f ? 0u : (uint32_t)(train->p_drop * 256.0f + 0.5f);
        p.thresh8 = thresh; p.keep_scale = 256.0f / (256.0f - (float)thresh);
        p.seed = train->seed; p.seed_dev = train->seed_dev; p.stream0 = train->stream0;
        for (int s = 0; s < num_stages; ++s) {
            p.zsave[s] = reinterpret_cast<uint16_t*>(train->z_save[s]);
            p.asave[s] = s < num_stages - 1 ? reinterpret_cast<uint16_t*>(train->a_save[s]) : nullptr;
        }
        launch(tc::edge_mlp_tc_kernel<false, false, 5>);
    } else if (timeline) {
        if (decoupled) launch(tc::edge_mlp_tc_kernel<true, true>);
        else if (v == 1) launch(tc::edge_mlp_tc_kernel<true, false, 1>);
        else launch(tc::edge_mlp_tc_kernel<true, false>);
    } else {
        if (decoupled) launch(tc::edge_mlp_tc_kernel<false, true>);
        else if (v == 1) launch(tc::edge_mlp_tc_kernel<false, false, 1>);
        else launch(tc::edge_mlp_tc_kernel<false, false>);
    }
    return check_launch("rmp_edge_mlp_bf16");
}

static void* g_node_timeline = nullptr;
// tuning aid: device buffer (int64 [800]) receiving the node kernel's clock64 trace when RMP_NODE_TIMELINE is set
extern "C" void rmp_set_node_timeline(void* buf) { g_node_timeline = buf; }

extern "C" int rmp_node_mlp_bf16(const float* h_in, float* agg, int prologue_ln, const float* ln1_g, const float* ln1_b,
                                 int num_stages, const void* const* W_bf16, const int32_t* K, const int32_t* N,
                                 const int32_t* kind, const float* const* bias, const float* ln2_g, const float* ln2_b,
                                 float* x1_scratch, float* h_out, float* p_out, void* pn_out_bf16, int n_rows,
                                 void* stream) {
    RMP_REQUIRE(h_in && W_bf16 && K && N && kind && bias, RMP_E_ARG, "rmp_node_mlp_bf16: null pointer");
    RMP_REQUIRE(num_stages >= 1 && num_stages <= tc::kNodeMaxStages, RMP_E_SHAPE, "rmp_node_mlp_bf16: 1..5 stages");
    RMP_REQUIRE(!prologue_ln || (ln1_g && ln1_b), RMP_E_ARG, "rmp_node_mlp_bf16: LayerNorm-1 parameters missing");
    if (n_rows <= 0) return RMP_OK;
    RMP_REQUIRE(aligned16(h_in) && (!agg || aligned16(agg)) && (!x1_scratch || aligned32(x1_scratch)) &&
                (!h_out || aligned32(h_out)) && (!p_out || aligned32(p_out)) && (!pn_out_bf16 || aligned32(pn_out_bf16)),
                RMP_E_ALIGN, "rmp_node_mlp_bf16: h_in / agg must be 16-byte aligned, x1 / h_out / p_out / pn_out 32-byte aligned");
    tc::NodeParams p{};
    CUtensorMap maps[tc::kNodeMaxStages];
    memset(maps, 0, sizeof(maps));
    int width = 128;
    for (int s = 0; s < num_stages; ++s) {
        RMP_REQUIRE(W_bf16[s] && bias[s] && aligned16(W_bf16[s]) && aligned16(bias[s]), RMP_E_ALIGN,
                    "rmp_node_mlp_bf16: stage %d weights/bias null or unaligned", s);
        RMP_REQUIRE(K[s] == width && (N[s] == 128 || N[s] == 256), RMP_E_SHAPE,
                    "rmp_node_mlp_bf16: stage %d is %d->%d, expected %d->{128,256}", s, K[s], N[s], width);
        RMP_REQUIRE(kind[s] >= 0 && kind[s] <= 2, RMP_E_ARG, "rmp_node_mlp_bf16: stage kind");
        if (kind[s] == 1) RMP_REQUIRE(N[s] == 128 && ln2_g && ln2_b && x1_scratch && h_out && prologue_ln, RMP_E_ARG,
                                      "rmp_node_mlp_bf16: residual+LayerNorm stage needs 128 outputs, LN params, scratch");
        if (kind[s] == 2) RMP_REQUIRE(p_out && pn_out_bf16 && s == num_stages - 1 && N[s] == 256, RMP_E_ARG,
                                      "rmp_node_mlp_bf16: projection stage must be last, 256 wide, with p_out and pn_out");
        int rc = tc::make_map_bf16(&maps[s], W_bf16[s], (uint64_t)N[s], (uint64_t)K[s], (uint32_t)N[s]);
        if (rc) return rc;
        p.st[s].K = K[s]; p.st[s].N = N[s]; p.st[s].kind = kind[s]; p.st[s].bias = bias[s];
        width = N[s];
    }
    p.h_in = h_in; p.agg = agg; p.prologue_ln = prologue_ln; p.ln1_g = ln1_g; p.ln1_b = ln1_b;
    p.ln2_g = ln2_g; p.ln2_b = ln2_b; p.x1 = x1_scratch; p.h_out = h_out; p.p_out = p_out;
    p.pn_out = reinterpret_cast<uint16_t*>(pn_out_bf16);
    p.S = num_stages; p.n_rows = n_rows;
    p.timeline = reinterpret_cast<long long*>(getenv("RMP_NODE_TIMELINE") ? g_node_timeline : nullptr);
    const size_t smem = 1024 + 4 * tc::kKBlockBytes + tc::kRing * tc::kSlotBytes + sizeof(tc::NodeBarriers) + 16 +
                        2 * tc::kNodeParts * tc::kTileM * sizeof(float);
    cudaFuncSetAttribute(tc::node_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(n_rows, tc::kTileM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    tc::node_mlp_tc_kernel<<<grid, tc::kNodeThreads, smem, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3],
                                                                                  maps[4], p);
    return check_launch("rmp_node_mlp_bf16");
}
