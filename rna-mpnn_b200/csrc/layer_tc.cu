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
    long long* timeline;             // optional clock64 trace of CTA 0 / first epilogue warp (tests, tuning)
};

// Two edge tiles are in flight per CTA: epilogue group g (8 warps) owns the tiles with local index
// i % 2 == g, the single MMA thread serves whichever tile has its next operand ready, and the TMA
// producer keeps a 3-deep ring of h_E tiles.  While one group runs a stage epilogue on the CUDA cores
// the other group's GEMM runs on the tensor pipe.
constexpr int kGroups = 2;
constexpr int kABufs = 3;
constexpr int kEdgeThreads = 64 + kGroups * kEpiThreads;       // 576
constexpr int kSelSlots = 16;                                  // node slots per tile served by the selector MMA
constexpr int kSelBytes = 2 * kSelSlots * 128;                 // [2 K-blocks][16 slots][64 edges] bf16
constexpr uint32_t kEdgeTmemCols = 512;                        // acc g at 128*g, neighbour sums g at 256 + 32*g

struct alignas(8) Barriers {
    uint64_t w_full, a_full[kABufs], a_empty[kABufs];
    uint64_t acc_full[kGroups], act_ready[kGroups], agg_full[kGroups], agg_empty[kGroups];
};

// 18 warps: the register file is allocated to a CTA in 4-warp granules (20 x 32 x 96 = 61440), so 96 is the cap
__global__ void __launch_bounds__(kEdgeThreads, 1)
edge_mlp_tc_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_w,
                   const EdgeParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;                                         // [S][2][128][64] bf16, resident
    uint8_t* sA = sW + (size_t)p.num_stages * kTileBytes;       // [3 buffers][2][128][64] bf16
    uint8_t* sSel = sA + kABufs * kTileBytes;                   // [2 groups][2][16][64] bf16 selector
    Barriers* bars = reinterpret_cast<Barriers*>(sSel + kGroups * kSelBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    int* s_int = reinterpret_cast<int*>(s_tmem + 4);            // per group: ctr[128] start[128] end[128] misc[4]
    constexpr int kIntsPerGroup = 3 * kTileM + 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.n_edges + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.num_stages;

    if (threadIdx.x == 0) {
        mbar_init(&bars->w_full, 1);
        for (int b = 0; b < kABufs; ++b) { mbar_init(&bars->a_full[b], 1); mbar_init(&bars->a_empty[b], 1); }
        for (int g = 0; g < kGroups; ++g) {
            mbar_init(&bars->acc_full[g], 1);
            mbar_init(&bars->act_ready[g], kGroups * kEpiThreads);
            mbar_init(&bars->agg_full[g], 1);
            mbar_init(&bars->agg_empty[g], kGroups * kEpiThreads);
        }
        fence_barrier_init();
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
                const int b = i % kABufs;
                const int tile = blockIdx.x + i * gridDim.x;
                mbar_wait_relaxed(&bars->a_empty[b], ((i / kABufs) & 1) ^ 1, 400);
                mbar_expect_tx(&bars->a_full[b], kTileBytes);
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d(sA + (size_t)b * kTileBytes + kb * kKBlockBytes, &tm_e, &bars->a_full[b], kb * 64,
                                tile * kTileM);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer: serves the two in-flight tiles as their operands become ready ====
        if (lane == 0) {
            constexpr uint32_t idesc_stage = instr_desc_bf16(128, 128, 0);
            constexpr uint32_t idesc_sel = instr_desc_bf16_amn(128, kSelSlots);
            int ti[kGroups], st[kGroups];
            uint32_t ph_act[kGroups], ph_ae[kGroups];
            for (int g = 0; g < kGroups; ++g) { ti[g] = g; st[g] = 0; ph_act[g] = 0; ph_ae[g] = 0; }
            mbar_wait(&bars->w_full, 0);
            int remaining = my_tiles;
            while (remaining > 0) {
                bool progressed = false;
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    if (ti[g] >= my_tiles) continue;
                    const int i = ti[g], b = i % kABufs, s = st[g];
                    bool ready;
                    if (s == 0) ready = mbar_test(&bars->a_full[b], (i / kABufs) & 1);
                    else if (s < S) ready = mbar_test(&bars->act_ready[g], ph_act[g]);
                    else ready = mbar_test(&bars->act_ready[g], ph_act[g]) && mbar_test(&bars->agg_empty[g], ph_ae[g] ^ 1);
                    if (!ready) continue;
                    progressed = true;
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + (size_t)b * kTileBytes);
                    if (s < S) {
                        if (s > 0) ph_act[g] ^= 1;
                        const uint32_t w_addr = smem_u32(sW + (size_t)s * kTileBytes);
                        const uint32_t acc = tmem_base + (uint32_t)(g * 128);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {               // K = 128 = 8 x UMMA_K(16)
                            const uint32_t off = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                            umma_bf16(acc, smem_desc_sw128(a_addr + off, 16, 1024),
                                      smem_desc_sw128(w_addr + off, 16, 1024), idesc_stage, k > 0);
                        }
                        umma_commit(&bars->acc_full[g]);
                        st[g] = s + 1;
                    } else {
                        // neighbour sums, transposed: aggT[ch][slot] = sum_r m3[r][ch] * Sel[slot][r].
                        // A = m3 read MN-major in place, B = 16-slot selector (K-major).
                        ph_act[g] ^= 1; ph_ae[g] ^= 1;
                        const int* misc = s_int + g * kIntsPerGroup + 3 * kTileM;
                        if (*reinterpret_cast<const volatile int*>(misc + 1) == 0) {   // selector path (<= 16 node slots)
                            const uint32_t sel_addr = smem_u32(sSel + g * kSelBytes);
                            const uint32_t agg = tmem_base + 256u + (uint32_t)(g * 32);
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
                        st[g] = 0;
                        ti[g] = i + kGroups;
                        --remaining;
                    }
                }
                if (!progressed) __nanosleep(120);     // do not steal issue slots from the epilogue warps
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: 16 warps, time-multiplexed over the two in-flight tiles ==========
        // warp -> (TMEM lane group = warp % 4, column quarter): thread = one edge row x 32 columns.  All 16
        // warps run stage s of tile A, then stage s of tile B; while they work on one tile the tensor pipe
        // runs the next GEMM of the other, so neither side waits for the other in steady state.
        const int ew = warp - 2;
        const int quarter = ew >> 2;                     // 32-column slice; K-block = quarter / 2
        const int lane_grp = warp & 3;                   // TMEM lanes this warp may touch: 32*(warp%4)..+31
        const int row = lane_grp * 32 + lane;
        const int col0 = quarter * 32;
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        uint32_t ph_acc[kGroups] = {0, 0}, ph_agg[kGroups] = {0, 0};
        int tl_n = 0;
        const bool tl_on = p.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
        auto stamp = [&](int tag) {
            if (tl_on && tl_n < 4000) { p.timeline[tl_n * 2] = tag; p.timeline[tl_n * 2 + 1] = clock64(); ++tl_n; }
        };

        // (centre, neighbour) node of this thread's edge row in local tile i; -1 marks rows past the end
        auto load_idx = [&](int i, int& c, int& nb) {
            c = -1; nb = 0;
            if (i < my_tiles) {
                const int e = (blockIdx.x + i * gridDim.x) * kTileM + row;
                if (e < p.n_edges) { c = __ldg(p.ctr + e); nb = __ldg(p.nbr + e); }
            }
        };
        int c_nx0, nb_nx0, c_nx1, nb_nx1;
        load_idx(0, c_nx0, nb_nx0);
        load_idx(1, c_nx1, nb_nx1);
        // stage biases (stages >= 1) live in shared memory: broadcast LDS instead of global loads per chunk
        float* s_bias = reinterpret_cast<float*>(s_int + kGroups * kIntsPerGroup);       // [kMaxStages][128]
        s_bias += (4 - ((smem_u32(s_bias) >> 2) & 3)) & 3;                                // 16-byte aligned for LDS.128
        for (int t = threadIdx.x - 64; t < (S - 1) * 128; t += kGroups * kEpiThreads) {
            const int st = 1 + t / 128;
            const float* bsp = st == 1 ? p.bias[1] : (st == 2 ? p.bias[2] : p.bias[3]);
            s_bias[st * 128 + (t & 127)] = __ldg(bsp + (t & 127));
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kGroups * kEpiThreads) : "memory");

        for (int i0 = 0; i0 < my_tiles; i0 += kGroups) {
            stamp(90);
            const int c_cur[kGroups] = {c_nx0, c_nx1}, nb_cur[kGroups] = {nb_nx0, nb_nx1};
            int fn[kGroups] = {0, 0}, ns[kGroups] = {0, 0};

            // stage-0 addend of this row's 32 columns: Pc[centre] + Pn[neighbour], combined while the GEMM runs
            float4 add[8];
            auto gather_addend = [&](int g) {
                if (c_cur[g] >= 0) {
                    const float* pc = p.Pc + (size_t)c_cur[g] * p.ld_pc + col0;
                    const uint16_t* pn = p.Pn + (size_t)nb_cur[g] * p.ld_pn + col0;
                    uint32_t w[2][8], f[4][8];
                    ldg256(pn, w[0]); ldg256(pn + 16, w[1]);                 // 32 bf16 = two 32-byte sectors
#pragma unroll
                    for (int q = 0; q < 4; ++q) ldg256(pc + 8 * q, f[q]);    // 32 fp32 = four sectors
#pragma unroll
                    for (int q = 0; q < 8; ++q) {                            // columns 4q .. 4q+3
                        const uint32_t w0 = w[q >> 2][(q & 3) * 2], w1 = w[q >> 2][(q & 3) * 2 + 1];
                        const uint32_t* fq = &f[q >> 1][(q & 1) * 4];
                        add[q].x = __uint_as_float(fq[0]) + __uint_as_float(w0 << 16);
                        add[q].y = __uint_as_float(fq[1]) + __uint_as_float(w0 & 0xffff0000u);
                        add[q].z = __uint_as_float(fq[2]) + __uint_as_float(w1 << 16);
                        add[q].w = __uint_as_float(fq[3]) + __uint_as_float(w1 & 0xffff0000u);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) add[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            gather_addend(0);
            stamp(95);
            // streaming index loads (HBM) are issued only now, after the L2 gather was consumed, so that nothing
            // on the critical path shares a scoreboard with them; they land during the stage epilogues
            load_idx(i0 + kGroups, c_nx0, nb_nx0);       // indices of the next pair, a whole pair of tiles ahead
            load_idx(i0 + kGroups + 1, c_nx1, nb_nx1);
#pragma unroll
            for (int g = 0; g < kGroups; ++g)            // node slots of the two tiles, consumed after stage 0
                if (i0 + g < my_tiles) {
                    const int e_base = (blockIdx.x + (i0 + g) * gridDim.x) * kTileM;
                    fn[g] = __ldg(p.ctr + e_base);
                    ns[g] = __ldg(p.ctr + min(e_base + kTileM, p.n_edges) - 1);
                }

            // one (tile, stage) epilogue; FIRST selects the stage-0 flavour at compile time so that the gathered
            // addend registers are live in that code only
            auto stage = [&](auto first_tag, int s, int g) {
                constexpr bool FIRST = decltype(first_tag)::value;
                const int i = i0 + g;
                if (i >= my_tiles) return;
                const bool valid = c_cur[g] >= 0;
                const bool bypass = (ns[g] - fn[g] + 1) > kSelSlots;
                uint8_t* a_buf = sA + (size_t)(i % kABufs) * kTileBytes;
                stamp(100 + s * 10 + g);                       // about to wait for the accumulator
                mbar_wait(&bars->acc_full[g], ph_acc[g]); ph_acc[g] ^= 1;
                tc_fence_after();
                stamp(200 + s * 10 + g);                       // accumulator ready
                float v[32];
                tmem_ld32(t_row + (uint32_t)(g * 128 + col0), v);
                stamp(300 + s * 10 + g);                       // TMEM load done
                if constexpr (FIRST) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        v[4 * q + 0] += add[q].x; v[4 * q + 1] += add[q].y;
                        v[4 * q + 2] += add[q].z; v[4 * q + 3] += add[q].w;
                    }
                } else {
                    const float4* bs = reinterpret_cast<const float4*>(s_bias + s * 128 + col0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 x = bs[q];
                        v[4 * q + 0] += x.x; v[4 * q + 1] += x.y; v[4 * q + 2] += x.z; v[4 * q + 3] += x.w;
                    }
                }
                stamp(400 + s * 10 + g);                       // gathers / bias added
                stamp(500 + s * 10 + g);
                if (p.debug != nullptr || (s == S - 1 && bypass)) {
                    // fp32 GELU: debug dump (tests) and the tiny-degree bypass that sums with vector reductions
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = gelu_tanh_fast(v[q]);
                    if (p.debug && valid) {
                        const int e = (blockIdx.x + i * gridDim.x) * kTileM + row;
                        float* dbg = p.debug + ((size_t)s * p.n_edges + e) * 128 + col0;
#pragma unroll
                        for (int q = 0; q < 32; ++q) dbg[q] = v[q];
                    }
                    if (s == S - 1 && bypass) {
                        if (valid) {
                            float* dst = p.agg + (size_t)c_cur[g] * 128 + col0;
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
                stamp(600 + s * 10 + g);                       // operand stored
                if constexpr (FIRST) {
                    if (g == 0) gather_addend(1);              // v[] is dead: start the other tile's gather
                    // selector matrix of this tile (rows = node slots, K = the tile's 128 edge rows); the MMA
                    // thread reads it (and the bypass flag) only after the last stage's act_ready
                    if (threadIdx.x == 64) (s_int + g * kIntsPerGroup + 3 * kTileM)[1] = bypass;
                    if (!bypass && row < kSelSlots && quarter < 2) {
                        const int e_base = (blockIdx.x + i * gridDim.x) * kTileM;
                        int st = 0, en = 0;
                        if (row < ns[g] - fn[g] + 1) {
                            st = max(__ldg(p.row_ptr + fn[g] + row) - e_base, 0);
                            en = min(__ldg(p.row_ptr + fn[g] + row + 1) - e_base, kTileM);
                        }
                        uint8_t* sel = sSel + g * kSelBytes + quarter * (kSelSlots * 128);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            uint32_t w[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int r0 = quarter * 64 + j * 8 + q * 2;
                                const uint32_t lo = (r0 >= st && r0 < en) ? 0x3F80u : 0u;
                                const uint32_t hi = (r0 + 1 >= st && r0 + 1 < en) ? 0x3F80u : 0u;
                                w[q] = lo | (hi << 16);
                            }
                            *reinterpret_cast<uint4*>(sel + sw128_chunk(row, j)) = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(&bars->act_ready[g]);
                stamp(700 + s * 10 + g);                       // fences + arrive done
            };
            stage(std::true_type{}, 0, 0);
            stage(std::true_type{}, 0, 1);
            for (int s = 1; s < S; ++s) {
                stage(std::false_type{}, s, 0);
                stage(std::false_type{}, s, 1);
            }

            // ---- neighbour sums out of TMEM: lane = channel, columns = node slots, 4 slots per column quarter ----
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                const int i = i0 + g;
                if (i >= my_tiles) continue;
                stamp(800 + g);
                mbar_wait(&bars->agg_full[g], ph_agg[g]); ph_agg[g] ^= 1;
                tc_fence_after();
                stamp(810 + g);
                const int nslots = ns[g] - fn[g] + 1;
                if (nslots <= kSelSlots && quarter * 4 < nslots) {       // warp-uniform
                    float v[4];
                    tmem_ld4(t_row + 256u + (uint32_t)(g * 32 + quarter * 4), v);
                    float* dst = p.agg + (size_t)(fn[g] + quarter * 4) * 128 + row;      // row == channel here
#pragma unroll
                    for (int n = 0; n < 4; ++n)
                        if (quarter * 4 + n < nslots)
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + n * 128), "f"(v[n] * p.alpha) : "memory");
                }
                tc_fence_before();
                mbar_arrive(&bars->agg_empty[g]);
                stamp(820 + g);
            }
        }
        if (tl_on) p.timeline[7999] = tl_n;
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
    const size_t smem = 1024 + (size_t)(num_stages + tc::kABufs) * tc::kTileBytes + tc::kGroups * tc::kSelBytes +
                        sizeof(tc::Barriers) + 16 + tc::kGroups * (3 * tc::kTileM + 4) * sizeof(int) +
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
