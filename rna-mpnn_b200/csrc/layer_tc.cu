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
// Roles (704 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = gather
// (accumulator pre-initialisation with the gathered projections), warps 6..21 = two epilogue groups.
#include "tc_common.cuh"

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
    float* debug;                    // optional [num_stages][n_edges][128] post-activation dump
    long long* timeline;             // optional int64[32768] clock64 trace of CTA 0 (per epilogue warp + MMA thread; tuning)
};

// Pipeline.  A tile owns slot = i % 3: one 32 KB A buffer (h_E, then the activations in place) and one
// 128-column TMEM accumulator.
//   gather warps : acc[slot] <- Pc[centre] + Pn[neighbour]          (tcgen05.st, a tile or more ahead)
//   MMA thread   : acc[slot] += h_E . W1e^T as soon as the TMA tile and the pre-initialised accumulator are
//                  there, so the stage-0 pre-activation of tile i+2 is complete before an epilogue group asks
//   epilogue     : group g (8 warps) owns the tiles with i % 2 == g and runs their stage epilogues
//                  (bias + GELU -> bf16 A operand in place); the two groups are decoupled, so while one is on
//                  the CUDA cores / MUFU the other group's GEMM is on the tensor pipe.
// The epilogue never touches global memory on its critical path: its only inputs are TMEM and shared memory.
constexpr int kGroups = 2;
constexpr int kSlots = 3;
constexpr int kGatherThreads = 128;
constexpr int kEdgeThreads = 64 + kGatherThreads + kGroups * kEpiThreads;   // 704
constexpr int kSelSlots = 16;                                  // node slots per tile served by the selector MMA
constexpr int kSelBytes = 2 * kSelSlots * 128;                 // [2 K-blocks][16 slots][64 edges] bf16
constexpr uint32_t kEdgeTmemCols = 512;                        // acc slot s at 128*s, neighbour sums g at 384 + 32*g
constexpr uint32_t kAggCol = 384;

struct alignas(8) Barriers {
    uint64_t w_full, a_full[kSlots], a_empty[kSlots];
    uint64_t init_done[kSlots], acc_free[kSlots], acc_full[kSlots], act_ready[kSlots];
    uint64_t agg_full[kGroups], agg_empty[kGroups];
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kEdgeThreads, 1)
edge_mlp_tc_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_w,
                   const EdgeParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;                                         // [S][2][128][64] bf16, resident
    uint8_t* sA = sW + (size_t)p.num_stages * kTileBytes;       // [3 slots][2][128][64] bf16
    uint8_t* sSel = sA + kSlots * kTileBytes;                   // [2 groups][2][16][64] bf16 selector
    Barriers* bars = reinterpret_cast<Barriers*>(sSel + kGroups * kSelBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    int* s_meta = reinterpret_cast<int*>(s_tmem + 4);           // per slot: first node, last node of the tile
    float* s_bias = reinterpret_cast<float*>(s_meta + 2 * kSlots + 2);                // [kMaxStages][128]
    s_bias += (4 - ((smem_u32(s_bias) >> 2) & 3)) & 3;                                // 16-byte aligned for LDS.128

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.n_edges + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.num_stages;

    if (threadIdx.x == 0) {
        mbar_init(&bars->w_full, 1);
        for (int b = 0; b < kSlots; ++b) {
            mbar_init(&bars->a_full[b], 1); mbar_init(&bars->a_empty[b], 1);
            mbar_init(&bars->init_done[b], kGatherThreads);
            mbar_init(&bars->acc_free[b], kEpiThreads);
            mbar_init(&bars->acc_full[b], 1);
            mbar_init(&bars->act_ready[b], kEpiThreads);
        }
        for (int g = 0; g < kGroups; ++g) { mbar_init(&bars->agg_full[g], 1); mbar_init(&bars->agg_empty[g], kEpiThreads); }
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
                const int b = i % kSlots;
                const int tile = blockIdx.x + i * gridDim.x;
                mbar_wait_relaxed(&bars->a_empty[b], ((i / kSlots) & 1) ^ 1, 200);
                mbar_expect_tx(&bars->a_full[b], kTileBytes);
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d(sA + (size_t)b * kTileBytes + kb * kKBlockBytes, &tm_e, &bars->a_full[b], kb * 64,
                                tile * kTileM);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer: serves the tiles in flight as their operands become ready ====
        if (lane == 0) {
            constexpr uint32_t idesc_stage = instr_desc_bf16(128, 128, 0);
            constexpr uint32_t idesc_sel = instr_desc_bf16_amn(128, kSelSlots);
            int ti[kSlots], st[kSlots];
            bool sel_path[kSlots];             // latched at stage-0 issue: s_meta[slot] is rewritten once acc_free fires
#pragma unroll
            for (int b = 0; b < kSlots; ++b) { ti[b] = b; st[b] = 0; sel_path[b] = true; }
            mbar_wait(&bars->w_full, 0);
            int remaining = my_tiles;
            int tlm = 0;
            const bool tlm_on = p.timeline != nullptr && blockIdx.x == 0;
            auto mstamp = [&](int tag) {        // MMA-thread trace lives in the upper half of the buffer
                if (tlm_on && tlm < 1990) { p.timeline[25600 + tlm * 2] = tag; p.timeline[25601 + tlm * 2] = clock64(); ++tlm; }
            };
            while (remaining > 0) {
                bool progressed = false;
                // pass 0: later stages of tiles in flight (an epilogue group is waiting for these);
                // pass 1: stage-0 GEMMs of tiles ahead (nobody waits yet) -- only when pass 0 found nothing
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    if (pass == 1 && progressed) break;
#pragma unroll
                    for (int b = 0; b < kSlots; ++b) {
                        const int i = ti[b], s = st[b];
                        if (i >= my_tiles || (pass == 0) != (s > 0)) continue;
                        const uint32_t par_tile = (uint32_t)((i / kSlots) & 1);
                        const int g = i & 1, k_g = i >> 1;           // group and per-group tile counter
                        bool ready;
                        if (s == 0) ready = mbar_test(&bars->a_full[b], par_tile) && mbar_test(&bars->init_done[b], par_tile);
                        else {
                            ready = mbar_test(&bars->act_ready[b], (uint32_t)((S * (i / kSlots) + s - 1) & 1));
                            if (s == S && k_g > 0) ready = ready && mbar_test(&bars->agg_empty[g], (uint32_t)((k_g - 1) & 1));
                        }
                        if (!ready) continue;
                        progressed = true;
                        tc_fence_after();
                        mstamp(1000 * (i + 1) + s);
                        const uint32_t a_addr = smem_u32(sA + (size_t)b * kTileBytes);
                        if (s < S) {
                            const uint32_t w_addr = smem_u32(sW + (size_t)s * kTileBytes);
                            const uint32_t acc = tmem_base + (uint32_t)(b * 128);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {               // K = 128 = 8 x UMMA_K(16)
                                const uint32_t off = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                                // stage 0 accumulates onto the gathered projections the gather warps stored
                                umma_bf16(acc, smem_desc_sw128(a_addr + off, 16, 1024),
                                          smem_desc_sw128(w_addr + off, 16, 1024), idesc_stage, (s == 0) || (k > 0));
                            }
                            umma_commit(&bars->acc_full[b]);
                            if (s == 0) {
                                const volatile int* meta = s_meta + 2 * b;
                                sel_path[b] = meta[1] - meta[0] + 1 <= kSelSlots;
                            }
                            st[b] = s + 1;
                        } else {
                            // neighbour sums, transposed: aggT[ch][slot] = sum_r m3[r][ch] * Sel[slot][r].
                            // A = m3 read MN-major in place, B = 16-slot selector (K-major).
                            if (sel_path[b]) {                              // selector path (<= 16 node slots)
                                const uint32_t sel_addr = smem_u32(sSel + g * kSelBytes);
                                const uint32_t agg = tmem_base + kAggCol + (uint32_t)(g * 32);
#pragma unroll
                                for (int k = 0; k < 8; ++k) {           // K = 128 edge rows = 8 x 16
                                    const uint32_t a_off = (uint32_t)(k * 2048);            // 16 rows of 128 B
                                    const uint32_t b_off = (uint32_t)((k >> 2) * (kSelSlots * 128) + (k & 3) * 32);
                                    umma_bf16(agg, smem_desc_sw128(a_addr + a_off, kKBlockBytes, 1024),
                                              smem_desc_sw128(sel_addr + b_off, 16, 1024), idesc_sel, k > 0);
                                }
                            }
                            umma_commit(&bars->agg_full[g]);
                            umma_commit(&bars->a_empty[b]);
                            st[b] = 0;
                            ti[b] = i + kSlots;
                            --remaining;
                        }
                    }
                }
                if (!progressed) __nanosleep(32);      // do not steal issue slots from the epilogue warps
            }
            if (tlm_on) p.timeline[32016] = tlm;
        }
        __syncwarp();
    } else if (warp < 2 + kGatherThreads / 32) {
        // ===================== gather warps: acc[slot] <- Pc[centre] + Pn[neighbour] (thread = edge row) =====
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        auto load_idx = [&](int i, int& c, int& nb) {
            c = -1; nb = 0;
            if (i < my_tiles) {
                const int e = (blockIdx.x + i * gridDim.x) * kTileM + row;
                if (e < p.n_edges) { c = __ldg(p.ctr + e); nb = __ldg(p.nbr + e); }
            }
        };
        int c_nx, nb_nx;
        load_idx(0, c_nx, nb_nx);
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i % kSlots;
            const int c = c_nx, nb = nb_nx;
            const int e_base = (blockIdx.x + i * gridDim.x) * kTileM;
            const int last_row = min(kTileM, p.n_edges - e_base) - 1;
            const float* pc = p.Pc + (size_t)max(c, 0) * p.ld_pc;
            const uint16_t* pn = p.Pn + (size_t)nb * p.ld_pn;
            // first half of the neighbour row is requested before the (possibly blocking) wait for the accumulator
            uint32_t w[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q) ldg256(pn + 16 * q, w[q]);
            load_idx(i + 1, c_nx, nb_nx);                      // streaming index loads of the next tile
            mbar_wait(&bars->acc_free[b], (uint32_t)(((i / kSlots) & 1) ^ 1));
            tc_fence_after();
            if (row == 0) s_meta[2 * b] = c;
            if (row == last_row) s_meta[2 * b + 1] = c;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) ldg256(pn + 64 + 16 * q, w[q]);
                }
#pragma unroll
                for (int q2 = 0; q2 < 2; ++q2) {               // 32 columns: h*64 + q2*32
                    const int col0 = h * 64 + q2 * 32;
                    uint32_t f[4][8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) ldg256(pc + col0 + 8 * q, f[q]);
                    float v[32];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {             // bf16 pair q of this 32-column chunk
                        const uint32_t ww = w[q2 * 2 + (q >> 3)][q & 7];
                        v[2 * q] = __uint_as_float(f[q >> 2][(2 * q) & 7]) + __uint_as_float(ww << 16);
                        v[2 * q + 1] = __uint_as_float(f[q >> 2][(2 * q + 1) & 7]) + __uint_as_float(ww & 0xffff0000u);
                    }
                    if (c < 0) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = 0.f;
                    }
                    tmem_st32(t_row + (uint32_t)(b * 128 + col0), v);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->init_done[b]);
        }
    } else {
        // ===================== epilogue: two decoupled groups of 8 warps, group g owns tiles i % 2 == g ==========
        // warp -> (TMEM lane group = warp % 4, column half): thread = one edge row x 64 columns (two chunks of 32)
        const int ew = warp - (2 + kGatherThreads / 32);
        const int g = ew >> 3;
        const int half = (ew >> 2) & 1;
        const int lane_grp = warp & 3;                   // TMEM lanes this warp may touch: 32*(warp%4)..+31
        const int row = lane_grp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        int tl_n = 0;
        const bool tl_on = p.timeline != nullptr && blockIdx.x == 0 && lane == 0;
        long long* tl_buf = p.timeline + ew * 1600;
        auto stamp = [&](int tag) {
            if (tl_on && tl_n < 799) { tl_buf[tl_n * 2] = tag; tl_buf[tl_n * 2 + 1] = clock64(); ++tl_n; }
        };

        for (int i = g; i < my_tiles; i += kGroups) {
            const int b = i % kSlots;
            const int pb = S * (i / kSlots);
            const int e_base = (blockIdx.x + i * gridDim.x) * kTileM;
            const bool valid = e_base + row < p.n_edges;
            uint8_t* a_buf = sA + (size_t)b * kTileBytes;
            const uint32_t t_acc = t_row + (uint32_t)(b * 128 + half * 64);
            int fn = 0, nslots = 1;
            bool bypass = false;
            stamp(90);
            for (int s = 0; s < S; ++s) {
                stamp(100 + s * 10);                           // about to wait for the accumulator
                mbar_wait(&bars->acc_full[b], (uint32_t)((pb + s) & 1));
                tc_fence_after();
                stamp(200 + s * 10);                           // accumulator ready
                if (s == 0) {
                    fn = reinterpret_cast<const volatile int*>(s_meta)[2 * b];
                    nslots = reinterpret_cast<const volatile int*>(s_meta)[2 * b + 1] - fn + 1;
                    bypass = nslots > kSelSlots;
                }
                const bool slow = p.debug != nullptr || (s == S - 1 && bypass);
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const int col0 = half * 64 + ch * 32;
                    float v[32];
                    tmem_ld32(t_acc + (uint32_t)(ch * 32), v);
                    if (s == S - 1 && ch == 1) {               // this thread is done with the accumulator
                        tc_fence_before();
                        mbar_arrive(&bars->acc_free[b]);
                    }
                    if (s > 0) {
                        const float4* bs = reinterpret_cast<const float4*>(s_bias + s * 128 + col0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 x = bs[q];
                            v[4 * q + 0] += x.x; v[4 * q + 1] += x.y; v[4 * q + 2] += x.z; v[4 * q + 3] += x.w;
                        }
                    }
                    if (slow) {
                        // fp32 GELU: debug dump (tests) and the tiny-degree bypass that sums with vector reductions
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = gelu_tanh_fast(v[q]);
                        if (p.debug && valid) {
                            float* dbg = p.debug + ((size_t)s * p.n_edges + e_base + row) * 128 + col0;
#pragma unroll
                            for (int q = 0; q < 32; ++q) dbg[q] = v[q];
                        }
                        if (s == S - 1 && bypass) {
                            if (valid) {
                                float* dst = p.agg + (size_t)__ldg(p.ctr + e_base + row) * 128 + col0;
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    red_add_v4(dst + 4 * q, v[4 * q] * p.alpha, v[4 * q + 1] * p.alpha,
                                               v[4 * q + 2] * p.alpha, v[4 * q + 3] * p.alpha);
                            }
                        } else {
                            store_act_chunk(a_buf, row, col0, v);
                        }
                    } else {
                        store_gelu_chunk(a_buf, row, col0, v);     // packed bf16x2 GELU straight into the A operand
                    }
                }
                stamp(600 + s * 10);                           // operand stored
                if (s == 0 && !bypass && row < kSelSlots) {
                    // selector matrix of this tile (rows = node slots, K = the tile's 128 edge rows; this thread
                    // fills K-block `half` of slot `row`); the previous tile of this group is done with it
                    int st = 0, en = 0;
                    if (row < nslots) {
                        st = max(__ldg(p.row_ptr + fn + row) - e_base, 0);
                        en = min(__ldg(p.row_ptr + fn + row + 1) - e_base, kTileM);
                    }
                    uint8_t* sel = sSel + g * kSelBytes + half * (kSelSlots * 128);
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
                        *reinterpret_cast<uint4*>(sel + sw128_chunk(row, j)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&bars->act_ready[b]);
                stamp(700 + s * 10);                           // fences + arrive done
            }

            // ---- neighbour sums out of TMEM: lane = channel, columns = node slots, 8 slots per column half ----
            stamp(800);
            mbar_wait(&bars->agg_full[g], (uint32_t)((i >> 1) & 1));
            tc_fence_after();
            stamp(810);
            if (!bypass && half * 8 < nslots) {                          // warp-uniform
                float v[8];
                tmem_ld8(t_row + kAggCol + (uint32_t)(g * 32 + half * 8), v);
                float* dst = p.agg + (size_t)(fn + half * 8) * 128 + row;          // row == channel here
#pragma unroll
                for (int n = 0; n < 8; ++n)
                    if (half * 8 + n < nslots)
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + n * 128), "f"(v[n] * p.alpha) : "memory");
            }
            tc_fence_before();
            mbar_arrive(&bars->agg_empty[g]);
            stamp(820);
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

__global__ void __launch_bounds__(kThreads, 1)
node_mlp_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                   const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3,
                   const __grid_constant__ CUtensorMap tm4, const NodeParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                                  // [4 K-blocks][128][64] bf16 activation (in place)
    uint8_t* sW = sA + 4 * kKBlockBytes;                 // weight ring
    NodeBarriers* bars = reinterpret_cast<NodeBarriers*>(sW + kRing * kSlotBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    float* s_red = reinterpret_cast<float*>(s_tmem + 4);  // [2 stats][2 halves][128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.n_rows + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.S;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRing; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->act_ready, kEpiThreads);
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
        if (lane == 0) {
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
        if (lane == 0) {
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
        const int half = (warp - 2) >> 2;
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        float* red_sum = s_red;                    // [2][128]
        float* red_sq = s_red + 2 * kTileM;        // [2][128]
        uint32_t ph_acc = 0;
        int tl_n = 0;
        const bool tl_on = p.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
        auto stamp = [&](int tag) {
            if (tl_on && tl_n < 400) { p.timeline[tl_n * 2] = tag; p.timeline[tl_n * 2 + 1] = clock64(); ++tl_n; }
        };
        stamp(1);

        // LayerNorm over the 128 channels of a row held as 64 + 64 by the two column halves
        auto layer_norm64 = [&](float (&x)[64], const float* gamma, const float* beta, int c0) {
            float sm = 0.f;
#pragma unroll
            for (int q = 0; q < 64; ++q) sm += x[q];
            red_sum[half * kTileM + row] = sm;
            epi_bar_sync();
            const float mu = (sm + red_sum[(half ^ 1) * kTileM + row]) * (1.f / 128.f);
            float sq = 0.f;
#pragma unroll
            for (int q = 0; q < 64; ++q) { x[q] -= mu; sq = fmaf(x[q], x[q], sq); }
            red_sq[half * kTileM + row] = sq;
            epi_bar_sync();
            const float inv = 1.f / sqrtf((sq + red_sq[(half ^ 1) * kTileM + row]) * (1.f / 128.f) + 1e-5f);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0) + q);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c0) + q);
                x[4 * q + 0] = fmaf(x[4 * q + 0] * inv, g.x, b.x); x[4 * q + 1] = fmaf(x[4 * q + 1] * inv, g.y, b.y);
                x[4 * q + 2] = fmaf(x[4 * q + 2] * inv, g.z, b.z); x[4 * q + 3] = fmaf(x[4 * q + 3] * inv, g.w, b.w);
            }
        };
        auto store_row64 = [&](float* dst, const float (&x)[64]) {      // 8 full 32-byte sectors per thread
#pragma unroll
            for (int q = 0; q < 8; ++q) stg256f(dst + 8 * q, &x[8 * q]);
        };
        auto to_operand64 = [&](const float (&x)[64], int c0) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float v[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = x[ch * 32 + q];
                store_act_chunk(sA, row, c0 + ch * 32, v);
            }
        };

        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int r = tile * kTileM + row;
            const bool valid = r < p.n_rows;
            const int c0 = half * 64;
            // ---- prologue: x1 = LN1(h + agg) (or h).  One warp per row, lanes across the 128 channels: fully
            //      coalesced 512-byte row accesses, LayerNorm by warp shuffles, bf16 copy -> swizzled A operand ----
            {
                const int ew = warp - 2;                               // 8 warps x 16 rows
                float4 g1 = make_float4(1.f, 1.f, 1.f, 1.f), b1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.prologue_ln) {
                    g1 = __ldg(reinterpret_cast<const float4*>(p.ln1_g) + lane);
                    b1 = __ldg(reinterpret_cast<const float4*>(p.ln1_b) + lane);
                }
                for (int r0 = 0; r0 < 16; r0 += 8) {                  // 8 rows per batch: 16 independent loads in flight
                    float4 xs[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int rg = tile * kTileM + ew * 16 + r0 + rr;
                        xs[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (rg < p.n_rows) xs[rr] = __ldg(reinterpret_cast<const float4*>(p.h_in + (size_t)rg * 128) + lane);
                    }
                    if (p.prologue_ln && p.agg) {
                        float4 as[8];
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            const int rg = tile * kTileM + ew * 16 + r0 + rr;
                            as[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (rg < p.n_rows) as[rr] = reinterpret_cast<const float4*>(p.agg + (size_t)rg * 128)[lane];
                        }
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            const int rg = tile * kTileM + ew * 16 + r0 + rr;
                            xs[rr].x += as[rr].x; xs[rr].y += as[rr].y; xs[rr].z += as[rr].z; xs[rr].w += as[rr].w;
                            if (rg < p.n_rows)                         // re-arm the neighbour-sum buffer
                                reinterpret_cast<float4*>(p.agg + (size_t)rg * 128)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int rl = ew * 16 + r0 + rr;              // row inside the tile
                        const int rg = tile * kTileM + rl;
                        float4 x = xs[rr];
                        if (p.prologue_ln) {
                            const float mu = warp_sum(x.x + x.y + x.z + x.w) * (1.f / 128.f);
                            x.x -= mu; x.y -= mu; x.z -= mu; x.w -= mu;
                            const float var = warp_sum(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w) * (1.f / 128.f);
                            const float inv = 1.f / sqrtf(var + 1e-5f);
                            x.x = fmaf(x.x * inv, g1.x, b1.x); x.y = fmaf(x.y * inv, g1.y, b1.y);
                            x.z = fmaf(x.z * inv, g1.z, b1.z); x.w = fmaf(x.w * inv, g1.w, b1.w);
                            if (rg < p.n_rows && p.x1) reinterpret_cast<float4*>(p.x1 + (size_t)rg * 128)[lane] = x;
                        }
                        // columns 4*lane .. 4*lane+3 -> K-block lane/16, 16-byte chunk (lane%16)/2, 8-byte half lane%2
                        *reinterpret_cast<uint2*>(sA + (lane >> 4) * kKBlockBytes + sw128_chunk(rl, (lane & 15) >> 1) + (lane & 1) * 8) =
                            make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
                    }
                }
                stamp(3);
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&bars->act_ready);
                epi_bar_sync();                                        // x1 rows written above are re-read by other warps
                stamp(4);
            }
            // ---- stage epilogues ----
            for (int s = 0; s < S; ++s) {
                const NodeStage stg = p.st[s];
                mbar_wait(&bars->acc_full, ph_acc); ph_acc ^= 1;
                tc_fence_after();
                stamp(10 + s);
                const int ncols = stg.N >> 1;                 // columns of this half
                const int cb = half * ncols;
                if (stg.kind == 1) {                          // N == 128: bias + residual + LN2 -> h_out (+ operand)
                    float x[64];
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        float v[32];
                        tmem_ld32(t_row + (uint32_t)(cb + ch * 32), v);
#pragma unroll
                        for (int q = 0; q < 32; ++q) x[ch * 32 + q] = v[q];
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float rsd[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (valid) ldg256f(p.x1 + (size_t)r * 128 + cb + 8 * q, rsd);
                        const float4 ba = __ldg(reinterpret_cast<const float4*>(stg.bias + cb) + 2 * q);
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(stg.bias + cb) + 2 * q + 1);
                        x[8 * q + 0] += ba.x + rsd[0]; x[8 * q + 1] += ba.y + rsd[1];
                        x[8 * q + 2] += ba.z + rsd[2]; x[8 * q + 3] += ba.w + rsd[3];
                        x[8 * q + 4] += bb.x + rsd[4]; x[8 * q + 5] += bb.y + rsd[5];
                        x[8 * q + 6] += bb.z + rsd[6]; x[8 * q + 7] += bb.w + rsd[7];
                    }
                    layer_norm64(x, p.ln2_g, p.ln2_b, cb);
                    if (valid) store_row64(p.h_out + (size_t)r * 128 + cb, x);
                    if (s + 1 < S) to_operand64(x, cb);
                } else {
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
                    epi_bar_sync();          // every TMEM read of this tile is done before the next prologue arrives
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

extern "C" int rmp_edge_mlp_bf16(const void* h_E_bf16, const void* W_bf16, int num_stages, const float* const* bias,
                                 const float* Pc, const void* Pn_bf16, const int32_t* ctr, const int32_t* nbr,
                                 const int32_t* row_ptr, int n_edges, float alpha, float* agg, float* debug,
                                 void* stream) {
    return rmp_edge_mlp_bf16_traced(h_E_bf16, W_bf16, num_stages, bias, Pc, Pn_bf16, ctr, nbr, row_ptr, n_edges, alpha,
                                    agg, debug, nullptr, stream);
}

extern "C" int rmp_edge_mlp_bf16_traced(const void* h_E_bf16, const void* W_bf16, int num_stages,
                                        const float* const* bias, const float* Pc, const void* Pn_bf16,
                                        const int32_t* ctr, const int32_t* nbr, const int32_t* row_ptr, int n_edges,
                                        float alpha, float* agg, float* debug, long long* timeline, void* stream) {
    const void* Pn = Pn_bf16;
    RMP_REQUIRE(row_ptr, RMP_E_ARG, "rmp_edge_mlp_bf16: row_ptr is null");
    RMP_REQUIRE(h_E_bf16 && W_bf16 && Pc && Pn && ctr && nbr && agg && bias, RMP_E_ARG, "rmp_edge_mlp_bf16: null pointer");
    RMP_REQUIRE(num_stages >= 1 && num_stages <= tc::kMaxStages, RMP_E_SHAPE, "rmp_edge_mlp_bf16: 1..4 message layers");
    RMP_REQUIRE(n_edges >= 0, RMP_E_ARG, "rmp_edge_mlp_bf16: bad sizes");
    RMP_REQUIRE(aligned16(h_E_bf16) && aligned16(W_bf16) && aligned16(Pc) && aligned16(Pn), RMP_E_ALIGN,
                "rmp_edge_mlp_bf16: pointers must be 16-byte aligned");
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
    const size_t smem = 1024 + (size_t)(num_stages + tc::kSlots) * tc::kTileBytes + tc::kGroups * tc::kSelBytes +
                        sizeof(tc::Barriers) + 16 + (2 * tc::kSlots + 2) * sizeof(int) +
                        tc::kMaxStages * 128 * sizeof(float) + 16;
    RMP_REQUIRE(smem <= 227 * 1024, RMP_E_SHAPE, "rmp_edge_mlp_bf16: %zu B of shared memory needed", smem);
    cudaFuncSetAttribute(tc::edge_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(n_edges, tc::kTileM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    tc::edge_mlp_tc_kernel<<<grid, tc::kEdgeThreads, smem, (cudaStream_t)stream>>>(tm_e, tm_w, p);
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
                        4 * tc::kTileM * sizeof(float);
    cudaFuncSetAttribute(tc::node_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(n_rows, tc::kTileM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    tc::node_mlp_tc_kernel<<<grid, tc::kThreads, smem, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3],
                                                                              maps[4], p);
    return check_launch("rmp_node_mlp_bf16");
}
