// Graph construction for the rdesign path: valid-only flattening bookkeeping,
// brute-force kNN with a warp-level top-K, and the flat edge list.
//
// Reference semantics: rdesign/model/feature.py:42-50 (_dist), :173-177 and
// :227-237 (mask_attend, shift, masked_select).  HBM-bound integer/byte work.
#include "common.cuh"

#include <cstdarg>

namespace rmp {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;

void note_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// --------------------------------------------------------------------------------
// graph index
// --------------------------------------------------------------------------------

// One block per structure: rank of every valid residue inside its structure, L_b, deg_b.
__global__ void __launch_bounds__(256) struct_rank_kernel(const float* __restrict__ mask, int N, int K,
                                                         int32_t* __restrict__ node_map,
                                                         int32_t* __restrict__ sinfo) {
    __shared__ int warp_tot[8];
    __shared__ int carry;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 256) {
        const int i = base + t;
        const bool ok = (i < N) && (mask[(size_t)b * N + i] == 1.0f);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int in_warp = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int before = carry;
        for (int q = 0; q < w; ++q) before += warp_tot[q];
        if (i < N) node_map[(size_t)b * N + i] = ok ? before + in_warp : -1;
        __syncthreads();
        if (t == 0) {
            int s = 0;
            for (int q = 0; q < 8; ++q) s += warp_tot[q];
            carry += s;
        }
        __syncthreads();
    }
    if (t == 0) {
        const int L = carry;
        sinfo[b * 4 + 0] = L;
        sinfo[b * 4 + 1] = L < K ? L : K;
    }
}

// Single block: exclusive scans over structures -> node_off, edge_off, totals.
__global__ void __launch_bounds__(1024) struct_scan_kernel(int B, int32_t* __restrict__ sinfo,
                                                          int32_t* __restrict__ counts) {
    __shared__ int wn[32], we[32];
    __shared__ int carry_n, carry_e;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) { carry_n = 0; carry_e = 0; }
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int b = base + t;
        int L = 0, E = 0;
        if (b < B) { L = sinfo[b * 4 + 0]; E = L * sinfo[b * 4 + 1]; }
        int sn = L, se = E;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, sn, o);
            const int c = __shfl_up_sync(0xffffffffu, se, o);
            if (lane >= o) { sn += a; se += c; }
        }
        if (lane == 31) { wn[w] = sn; we[w] = se; }
        __syncthreads();
        int pn = carry_n, pe = carry_e;
        for (int q = 0; q < w; ++q) { pn += wn[q]; pe += we[q]; }
        if (b < B) { sinfo[b * 4 + 2] = pn + sn - L; sinfo[b * 4 + 3] = pe + se - E; }
        __syncthreads();
        if (t == 1023) { carry_n = pn + sn; carry_e = pe + se; }
        __syncthreads();
    }
    if (t == 0) { counts[0] = carry_n; counts[1] = carry_e; }
}

__global__ void __launch_bounds__(256) node_finalize_kernel(int B, int N, const int32_t* __restrict__ sinfo,
                                                           const int32_t* __restrict__ counts,
                                                           int32_t* __restrict__ node_map,
                                                           int32_t* __restrict__ node_src,
                                                           int32_t* __restrict__ row_ptr) {
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx == 0) row_ptr[counts[0]] = counts[1];
    if (idx >= (size_t)B * N) return;
    const int b = (int)(idx / N);
    const int rank = node_map[idx];
    if (rank < 0) return;
    const int v = sinfo[b * 4 + 2] + rank;
    node_map[idx] = v;
    node_src[v] = (int32_t)idx;
    row_ptr[v] = sinfo[b * 4 + 3] + rank * sinfo[b * 4 + 1];
}

// --------------------------------------------------------------------------------
// kNN: one warp per query row, keys = (distance bits << 32 | index), ascending.
// The K best keys live in a warp-distributed sorted list: slot s is register s/32
// of lane s%32.  Candidates that beat the current K-th key are compacted into a
// per-warp ring in shared memory; every 32 of them are sorted (bitonic, across
// lanes) and merged into the list block by block, then the threshold is refreshed.
// Ordering (distance, then index) is total, so the result is deterministic and
// equals a stable ascending sort truncated to K, whatever the batching.
// --------------------------------------------------------------------------------
typedef unsigned long long u64;
constexpr u64 kKeyMax = ~0ull;
// rows handled by one warp: more rows amortise staging the structure in shared memory
__host__ __device__ inline int knn_rows_per_warp(int N) { return N <= 1024 ? 1 : (N <= 4096 ? 2 : 8); }
constexpr int kKnnWarps = 16;
constexpr int kKnnSmemPoints = 8192;      // structures up to this length are staged in smem
constexpr int kKnnRing = 64;              // candidate ring entries per warp

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ u64 umin64(u64 a, u64 b) { return a < b ? a : b; }
__device__ __forceinline__ u64 umax64(u64 a, u64 b) { return a < b ? b : a; }

// bitonic (across the 32 lanes) -> ascending
__device__ __forceinline__ u64 bitonic_merge32(u64 v, int lane) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const u64 o = shfl_xor_u64(v, m);
        v = (lane & m) ? umax64(v, o) : umin64(v, o);
    }
    return v;
}
// arbitrary -> ascending across the 32 lanes
__device__ __forceinline__ u64 bitonic_sort32(u64 v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int m = k >> 1; m > 0; m >>= 1) {
            const u64 o = shfl_xor_u64(v, m);
            const bool up = ((lane & k) == 0) || k == 32;        // direction of this lane's k-block
            const bool lower = (lane & m) == 0;
            v = (lower == up) ? umin64(v, o) : umax64(v, o);
        }
    }
    return v;
}
// merge 32 candidates (one per lane, any order) into the sorted list, dropping the 32 largest keys
template <int R>
__device__ __forceinline__ void warp_merge_batch(u64 (&slot)[R], u64 cand, int lane) {
    u64 b = bitonic_sort32(cand, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        // merge-split of two ascending blocks: the element-wise min / max against the reversed block are bitonic
        const u64 rev = __shfl_sync(0xffffffffu, b, 31 - lane);
        const u64 lo = umin64(slot[r], rev), hi = umax64(slot[r], rev);
        slot[r] = bitonic_merge32(lo, lane);
        b = bitonic_merge32(hi, lane);                           // the larger half moves up one block
    }
}

template <int R>
__device__ __forceinline__ u64 warp_slot_read(const u64 (&slot)[R], int s) {
    u64 out = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const u64 v = __shfl_sync(0xffffffffu, slot[r], s & 31);
        if (r == (s >> 5)) out = v;
    }
    return out;
}

// Adjusted distances (make_key below), exact rounding sequence, no contraction:
//   MODE 0 (rdesign/model/feature.py:42-48): sqrt.rn(((dx*dx + dy*dy) + dz*dz) + 1e-6); a pair with a masked end is
//           20001 = 10000 + (rowmax 10000 + 1).
//   MODE 1 (rnampnn/model/feature.py:218-227): the same on centroids; pairs with a padded end and the diagonal sit at
//           exactly 1e6 (1e-3 + 1e6 rounds to 1e6 in fp32) and take part in the ordering like any other candidate.

// MODE 0: rdesign (one atom per residue, self included, E_idx [B,N,K] with K <= N).
// MODE 1: rnampnn (centroid of the A atoms, self excluded, -1 sentinels, K may exceed N).
template <int R, bool SMEM, int MODE>
__global__ void __launch_bounds__(kKnnWarps * 32) knn_kernel(const float* __restrict__ X,
                                                             const float* __restrict__ mask, int N, int A,
                                                             int atom, int K, int rows_per_warp,
                                                             int64_t* __restrict__ E_idx, float* __restrict__ D_out) {
    extern __shared__ float4 s_pts[];
    __shared__ u64 s_ring[kKnnWarps][kKnnRing];
    __shared__ int s_len;
    const int b = blockIdx.y, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const float* Xb = X + (size_t)b * N * A * 3 + (size_t)atom * 3;
    const float* mb = mask + (size_t)b * N;
    auto load_point = [&](int j) -> float4 {
        const float* p = Xb + (size_t)j * A * 3;
        const float m = mb[j] == 1.0f ? 1.0f : 0.0f;
        if (MODE == 0) return make_float4(p[0], p[1], p[2], m);
        // coords.mean(dim=2): torch-CPU adds the seven atoms in the order 0,4,5,6,1,2,3, then divides
        // (oracle/rnampnn_oracle.py: centroids); other atom counts are summed in index order
        float c[3];
#pragma unroll
        for (int x = 0; x < 3; ++x) {
            float s = p[x];
            if (A == 7) {
                s = __fadd_rn(s, p[12 + x]); s = __fadd_rn(s, p[15 + x]); s = __fadd_rn(s, p[18 + x]);
                s = __fadd_rn(s, p[3 + x]);  s = __fadd_rn(s, p[6 + x]);  s = __fadd_rn(s, p[9 + x]);
            } else {
                for (int a = 1; a < A; ++a) s = __fadd_rn(s, p[a * 3 + x]);
            }
            c[x] = __fdiv_rn(s, (float)A);
        }
        return make_float4(c[0], c[1], c[2], m);
    };
    if (SMEM) {
        for (int j = t; j < N; j += kKnnWarps * 32) s_pts[j] = load_point(j);
        __syncthreads();
    }
    if (MODE == 1) {                                         // L_b = number of valid residues of this structure
        int cnt = 0;
        for (int base = 0; base < N; base += kKnnWarps * 32) {
            const int j = base + t;
            cnt += __syncthreads_count(j < N && mb[j] == 1.0f);
        }
        if (t == 0) s_len = cnt;
        __syncthreads();
    }
    u64* ring = s_ring[w];
    const unsigned lt_mask = (1u << lane) - 1u;
    const int row0 = (blockIdx.x * kKnnWarps + w) * rows_per_warp;
    for (int rr = 0; rr < rows_per_warp; ++rr) {
        const int i = row0 + rr;
        if (i >= N) break;                                   // warp-uniform
        const float4 pi = SMEM ? s_pts[i] : load_point(i);
        u64 slot[R];
#pragma unroll
        for (int r = 0; r < R; ++r) slot[r] = kKeyMax;
        u64 thresh = kKeyMax;
        int head = 0, cnt = 0;                               // ring state (warp-uniform)
        // Candidate blocks are visited outward from the query's own position (i/32, then +1, -1, +2, ...): along a chain
        // the sequence neighbours are the likely spatial neighbours, so the running K-th key tightens within the first
        // few blocks and almost nothing passes it afterwards.  Scanning j = 0..N-1 instead approaches the query
        // monotonically -- nearly every candidate below i beats the current K-th key and pays for a sort / merge
        // (measured 0.8 ms at 8 x 4096, K = 48).  The (distance, index) order is total, so the result is the same.
        const int nblk = (N + 31) >> 5;
        int blk_lo = i >> 5, blk_hi = (i >> 5) + 1;
        // squared prefilter: a valid pair with s > T^2 (1 + 2^-20) has sqrt.rn(s) > T = the current K-th distance and
        // can never enter the list, so its IEEE square root is skipped (the margin is 16x the rounding of T*T and of
        // sqrt.rn: every candidate that could tie or beat T still takes the exact path)
        float s_thr = INFINITY;
        auto next_block = [&](int tb) -> int {
            if ((((tb & 1) == 0) && blk_lo >= 0) || blk_hi >= nblk) return blk_lo--;
            return blk_hi++;
        };
        auto make_key = [&](int j) -> u64 {
            if (j >= N) return kKeyMax;
            const float4 pj = SMEM ? s_pts[j] : load_point(j);
            if (pi.w == 0.0f || pj.w == 0.0f || (MODE == 1 && j == i))
                return ((u64)__float_as_uint(MODE == 0 ? 20001.0f : 1e6f) << 32) | (unsigned)j;
            const float dx = __fsub_rn(pj.x, pi.x), dy = __fsub_rn(pj.y, pi.y), dz = __fsub_rn(pj.z, pi.z);
            const float sq = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)), 1e-6f);
            if (sq > s_thr) return kKeyMax;
            return ((u64)__float_as_uint(__fsqrt_rn(sq)) << 32) | (unsigned)j;
        };
        auto offer = [&](u64 key) {                            // warp-uniform control flow
            const bool pass = key < thresh;
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (m == 0) return;
            if (pass) ring[(head + cnt + __popc(m & lt_mask)) & (kKnnRing - 1)] = key;
            cnt += __popc(m);
            if (cnt >= 32) {
                __syncwarp();
                const u64 cand = ring[(head + lane) & (kKnnRing - 1)];
                __syncwarp();
                head = (head + 32) & (kKnnRing - 1);
                cnt -= 32;
                warp_merge_batch<R>(slot, cand, lane);
                thresh = warp_slot_read<R>(slot, K - 1);
                if (thresh != kKeyMax) {
                    const float T = __uint_as_float((unsigned)(thresh >> 32));
                    s_thr = __fmul_rn(__fmul_rn(T, T), 1.0000009537f);          // 1 + 2^-20
                }
            }
        };
        // two candidate blocks per iteration: their distance chains are independent and hide each other's latency
        for (int tb = 0; tb < nblk; tb += 2) {
            const int blk_a = next_block(tb);
            const int blk_b = tb + 1 < nblk ? next_block(tb + 1) : -1;
            const u64 key_a = make_key((blk_a << 5) + lane);
            const u64 key_b = blk_b >= 0 ? make_key((blk_b << 5) + lane) : kKeyMax;
            offer(key_a);
            offer(key_b);
        }
        if (cnt > 0) {                                       // the last, partial batch
            __syncwarp();
            const u64 cand = lane < cnt ? ring[(head + lane) & (kKnnRing - 1)] : kKeyMax;
            __syncwarp();
            warp_merge_batch<R>(slot, cand, lane);
        }
        const size_t o = ((size_t)b * N + i) * K;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int s = r * 32 + lane;
            if (s < K) {
                if (MODE == 0) {
                    E_idx[o + s] = (int64_t)(slot[r] & 0xffffffffull);
                } else {
                    // feature.py:236-254: columns past N, "self", slots k > valid_neighbours and padded rows -> -1
                    const int j = (int)(slot[r] & 0xffffffffull);
                    const int valid_nb = pi.w != 0.0f ? max(s_len - 1, 0) : 0;
                    const bool dead = slot[r] == kKeyMax || j == i || s > valid_nb || pi.w == 0.0f;
                    E_idx[o + s] = dead ? (int64_t)-1 : (int64_t)j;
                }
                if (D_out) D_out[o + s] = __uint_as_float((unsigned)(slot[r] >> 32));
            }
        }
    }
}

template <int R, int MODE>
static int launch_knn(const float* X, const float* mask, int B, int N, int A, int atom, int K,
                      int64_t* E_idx, float* D, cudaStream_t st) {
    const int rpw = knn_rows_per_warp(N);
    dim3 grid(ceil_div(N, kKnnWarps * rpw), B);
    if (N <= kKnnSmemPoints) {
        const size_t smem = (size_t)N * sizeof(float4);
        auto k = knn_kernel<R, true, MODE>;
        // the kernel also holds ~4 KB of static shared memory (candidate rings): opt in before the sum reaches 48 KB
        if (smem + 8 * 1024 > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, kKnnWarps * 32, smem, st>>>(X, mask, N, A, atom, K, rpw, E_idx, D);
    } else {
        knn_kernel<R, false, MODE><<<grid, kKnnWarps * 32, 0, st>>>(X, mask, N, A, atom, K, rpw, E_idx, D);
    }
    return check_launch(MODE == 0 ? "rmp_knn" : "rmp_res_graph");
}

// --------------------------------------------------------------------------------
// flat edge list
// --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) edge_list_kernel(const int64_t* __restrict__ E_idx,
                                                       const int32_t* __restrict__ node_map,
                                                       const int32_t* __restrict__ sinfo,
                                                       const int32_t* __restrict__ row_ptr, int B, int N, int K,
                                                       int ne_cap, int32_t* __restrict__ ctr,
                                                       int32_t* __restrict__ nbr, int64_t* __restrict__ E_flat) {
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (size_t)B * N * K) return;
    const int k = (int)(idx % K);
    const size_t bi = idx / K;
    const int b = (int)(bi / N);
    const int v = node_map[bi];
    if (v < 0 || k >= sinfo[b * 4 + 1]) return;
    const int e = row_ptr[v] + k;
    if (e >= ne_cap) return;
    const int j = (int)E_idx[idx];
    const int u = node_map[(size_t)b * N + j];
    ctr[e] = v;
    nbr[e] = u;
    if (E_flat) {
        E_flat[e] = v;
        E_flat[(size_t)ne_cap + e] = u;
    }
}

}  // namespace rmp

using namespace rmp;

extern "C" const char* rmp_version(void) { return "rmpnn_b200 0.1 (sm_100a)"; }
extern "C" const char* rmp_last_error_string(void) { return g_err; }
extern "C" int rmp_abi_struct_size(void) { return (int)sizeof(rmp_mlp_desc); }
extern "C" uint64_t rmp_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int rmp_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) {
        set_error("rmp_device_arch: %s", cudaGetErrorString(e));
        return -(int)e;
    }
    return major * 10 + minor;
}

extern "C" int rmp_graph_index(const float* mask, int B, int N, int K, int32_t* node_map, int32_t* node_src,
                               int32_t* sinfo, int32_t* row_ptr, int32_t* counts, void* stream) {
    RMP_REQUIRE(mask && node_map && node_src && sinfo && row_ptr && counts, RMP_E_ARG, "rmp_graph_index: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && K > 0, RMP_E_ARG, "rmp_graph_index: B, N, K must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    struct_rank_kernel<<<B, 256, 0, st>>>(mask, N, K, node_map, sinfo);
    struct_scan_kernel<<<1, 1024, 0, st>>>(B, sinfo, counts);
    const size_t total = (size_t)B * N;
    node_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(B, N, sinfo, counts, node_map, node_src,
                                                                           row_ptr);
    return check_launch("rmp_graph_index", 3);
}

extern "C" int rmp_knn(const float* X, const float* mask, int B, int N, int A, int atom, int K, int64_t* E_idx,
                       float* D_neighbors, void* stream) {
    RMP_REQUIRE(X && mask && E_idx, RMP_E_ARG, "rmp_knn: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && A > 0 && atom >= 0 && atom < A, RMP_E_ARG, "rmp_knn: bad sizes");
    RMP_REQUIRE(K > 0 && K <= N, RMP_E_SHAPE, "rmp_knn: K must be in [1, N] (pass min(top_k, N)), got %d", K);
    RMP_REQUIRE(K <= RMP_MAX_K, RMP_E_SHAPE, "rmp_knn: K=%d exceeds RMP_MAX_K=%d", K, RMP_MAX_K);
    RMP_REQUIRE(B <= 65535, RMP_E_SHAPE, "rmp_knn: B=%d exceeds 65535 structures per call", B);
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 32) return launch_knn<1, 0>(X, mask, B, N, A, atom, K, E_idx, D_neighbors, st);
    if (K <= 64) return launch_knn<2, 0>(X, mask, B, N, A, atom, K, E_idx, D_neighbors, st);
    return launch_knn<4, 0>(X, mask, B, N, A, atom, K, E_idx, D_neighbors, st);
}

extern "C" int rmp_res_graph(const float* coords, const float* mask, int B, int N, int A, int K,
                             int64_t* edge_index, float* D_neighbors, void* stream) {
    RMP_REQUIRE(coords && mask && edge_index, RMP_E_ARG, "rmp_res_graph: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && A > 0, RMP_E_ARG, "rmp_res_graph: bad sizes");
    RMP_REQUIRE(K > 0 && K <= RMP_MAX_K, RMP_E_SHAPE, "rmp_res_graph: K=%d outside [1, %d]", K, RMP_MAX_K);
    RMP_REQUIRE(B <= 65535, RMP_E_SHAPE, "rmp_res_graph: B=%d exceeds 65535 structures per call", B);
    cudaStream_t st = (cudaStream_t)stream;
    if (K <= 32) return launch_knn<1, 1>(coords, mask, B, N, A, 0, K, edge_index, D_neighbors, st);
    if (K <= 64) return launch_knn<2, 1>(coords, mask, B, N, A, 0, K, edge_index, D_neighbors, st);
    return launch_knn<4, 1>(coords, mask, B, N, A, 0, K, edge_index, D_neighbors, st);
}

extern "C" int rmp_edge_list(const int64_t* E_idx, const int32_t* node_map, const int32_t* node_src,
                             const int32_t* sinfo, const int32_t* row_ptr, const int32_t* counts, int B, int N,
                             int K, int ne_cap, int32_t* ctr, int32_t* nbr, int64_t* E_flat, void* stream) {
    (void)node_src;
    (void)counts;
    RMP_REQUIRE(E_idx && node_map && sinfo && row_ptr && ctr && nbr, RMP_E_ARG, "rmp_edge_list: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && K > 0 && ne_cap >= 0, RMP_E_ARG, "rmp_edge_list: bad sizes");
    if (ne_cap == 0) return RMP_OK;
    const size_t total = (size_t)B * N * K;
    edge_list_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        E_idx, node_map, sinfo, row_ptr, B, N, K, ne_cap, ctr, nbr, E_flat);
    return check_launch("rmp_edge_list");
}
