// Kernels of the secondary (rnampnn) path, SURVEY.md section 8a rows a-10..a-16 (fp32 CUDA-core arm).
//
// Reference semantics: rnampnn/model/feature.py:205-571 (graph, 28 intra-residue and 90 cross-residue raw
// features), rnampnn/model/functional.py:7-48 (GraphNormalization), :153-172 (RNABert: self-attention over
// a 4500-row padding), rnampnn/model/mpnn.py:154-294 (ResMPNN).  The reference keeps the padded [B,N,..]
// layout end to end and so do these kernels: rows are dense residues b*N+i, edges are (row, slot k).
// The Linear stacks run on the fused row-MLP kernel (rowmlp.cu); this file holds what is not a Linear.
#include "geom.cuh"

#include <cuda_bf16.h>

namespace rmp {

constexpr int kMaxAtoms = 7;

// --------------------------------------------------------------------------------
// per-edge bookkeeping from edge_index [B,N,K] (int64, -1 sentinel)
//   nbr_row [B*N*K] : dense row of the neighbour, -1 -> row b*N+0 (mpnn.py:173-174, "safe" index)
//   edge_keep       : 1 for edges that keep their embedding (edge_index != -1 and mask_i == 1), feature.py:564-569
//   edge_w          : (edge_index != -1) / max(#valid slots of the row, 1)   (mpnn.py:191, :215-219)
// --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) res_edge_rows_kernel(const int64_t* __restrict__ edge_index,
                                                           const float* __restrict__ mask, int B, int N, int K,
                                                           int32_t* __restrict__ ctr_row, int32_t* __restrict__ nbr_row,
                                                           float* __restrict__ edge_keep, float* __restrict__ edge_w) {
    const size_t row = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= (size_t)B * N) return;
    const int b = (int)(row / N);
    const int64_t* ei = edge_index + row * K;
    int cnt = 0;
    for (int k = 0; k < K; ++k) cnt += ei[k] != -1;
    const float inv = 1.0f / (float)(cnt > 0 ? cnt : 1);
    const float mi = mask[row] == 0.0f ? 0.0f : 1.0f;
    for (int k = 0; k < K; ++k) {
        const int64_t j = ei[k];
        const bool ok = j != -1;
        ctr_row[row * K + k] = (int32_t)row;
        nbr_row[row * K + k] = b * N + (ok ? (int)j : 0);
        edge_keep[row * K + k] = ok ? mi : 0.0f;
        edge_w[row * K + k] = ok ? inv : 0.0f;
    }
}

// --------------------------------------------------------------------------------
// intra-residue raw features (a-11): one thread per dense residue
//   [nd*(nd-1)/2 distances | na-2 bond-angle cosines | nh-3 bond-plane cosines]
// --------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) res_inside_kernel(const float* __restrict__ coords, const float* __restrict__ mask,
                                                        int rows, int A, int nd, int na, int nh, int raw_dim,
                                                        float* __restrict__ raw) {
    const int row = blockIdx.x * 128 + threadIdx.x;
    if (row >= rows) return;
    const float* p = coords + (size_t)row * A * 3;
    float* out = raw + (size_t)row * raw_dim;
    const bool ok = mask[row] != 0.0f;
    V3 c[kMaxAtoms];
#pragma unroll
    for (int a = 0; a < kMaxAtoms; ++a) c[a] = a < A ? ld3(p + a * 3) : V3{0.f, 0.f, 0.f};
    int o = 0;
    // feature.py:316-329: upper triangle, row-major (a < b), padded rows -> 1e6
#pragma unroll
    for (int a = 0; a < kMaxAtoms; ++a)
#pragma unroll
        for (int b2 = a + 1; b2 < kMaxAtoms; ++b2)
            if (b2 < nd) out[o++] = ok ? eps_dist(c[a], c[b2]) : 1e6f;
    // feature.py:349-357: cos between consecutive bond vectors, dot / (|a||b| + 1e-6), * mask
    V3 v[kMaxAtoms - 1];
    float nrm[kMaxAtoms - 1];
#pragma unroll
    for (int a = 0; a < kMaxAtoms - 1; ++a) {
        v[a] = c[a + 1] - c[a];
        nrm[a] = sqrtf(dot(v[a], v[a]));
    }
#pragma unroll
    for (int a = 0; a < kMaxAtoms - 2; ++a)
        if (a < na - 2) out[o++] = ok ? dot(v[a], v[a + 1]) / (nrm[a] * nrm[a + 1] + 1e-6f) : 0.f;
    // feature.py:377-383: unit bond vectors (eps 1e-6) -> unit plane normals (eps 1e-6) -> consecutive dots
    V3 nor[kMaxAtoms - 2];
#pragma unroll
    for (int a = 0; a < kMaxAtoms - 2; ++a) nor[a] = unit_eps(cross(unit_eps(v[a], 1e-6f), unit_eps(v[a + 1], 1e-6f)), 1e-6f);
#pragma unroll
    for (int a = 0; a < kMaxAtoms - 3; ++a)
        if (a < nh - 3) out[o++] = ok ? dot(nor[a + 1], nor[a]) : 0.f;
}

// --------------------------------------------------------------------------------
// cross-residue raw features (a-12): 64 edges per block, staged in smem, written coalesced
//   [nd*nd distances | (na-1)^2 bond cosines | (nh-2)^2 plane-normal cosines]
// --------------------------------------------------------------------------------
constexpr int kEdgeTile = 64;

__global__ void __launch_bounds__(kEdgeTile) res_cross_kernel(const float* __restrict__ coords,
                                                             const float* __restrict__ mask,
                                                             const int64_t* __restrict__ edge_index, int B, int N,
                                                             int A, int K, int nd, int na, int nh, int edge_dim,
                                                             float* __restrict__ edge_raw) {
    extern __shared__ float s_feat[];                  // [kEdgeTile][edge_dim + 1]
    const size_t total = (size_t)B * N * K;
    const size_t e0 = (size_t)blockIdx.x * kEdgeTile;
    const size_t e = e0 + threadIdx.x;
    const int stride = edge_dim + 1;
    if (e < total) {
        float* f = s_feat + (size_t)threadIdx.x * stride;
        const size_t row = e / K;
        const int b = (int)(row / N);
        const int64_t j = edge_index[e];
        const bool valid = j != -1 && mask[row] != 0.0f;   // feature.py:420: mask_i * (edge_index != -1)
        if (!valid) {
            for (int q = 0; q < nd * nd; ++q) f[q] = 1e6f;
            for (int q = nd * nd; q < edge_dim; ++q) f[q] = 0.f;
        } else {
            const float* pi = coords + row * A * 3;
            const float* pj = coords + ((size_t)b * N + (size_t)j) * A * 3;
            V3 ci[kMaxAtoms], cj[kMaxAtoms];
#pragma unroll
            for (int a = 0; a < kMaxAtoms; ++a) {
                ci[a] = a < A ? ld3(pi + a * 3) : V3{0.f, 0.f, 0.f};
                cj[a] = a < A ? ld3(pj + a * 3) : V3{0.f, 0.f, 0.f};
            }
            int o = 0;
            // feature.py:414-418: d(centre atom a, neighbour atom b), index a*nd + b
#pragma unroll
            for (int a = 0; a < kMaxAtoms; ++a)
#pragma unroll
                for (int b2 = 0; b2 < kMaxAtoms; ++b2)
                    if (a < nd && b2 < nd) f[o++] = eps_dist(ci[a], cj[b2]);
            V3 vi[kMaxAtoms - 1], vj[kMaxAtoms - 1];
#pragma unroll
            for (int a = 0; a < kMaxAtoms - 1; ++a) { vi[a] = ci[a + 1] - ci[a]; vj[a] = cj[a + 1] - cj[a]; }
            // feature.py:451-464: unit bond vectors (F.normalize eps 1e-12), index a*(na-1) + b
            {
                V3 ui[kMaxAtoms - 1], uj[kMaxAtoms - 1];
#pragma unroll
                for (int a = 0; a < kMaxAtoms - 1; ++a) { ui[a] = unit_eps(vi[a], 1e-12f); uj[a] = unit_eps(vj[a], 1e-12f); }
#pragma unroll
                for (int a = 0; a < kMaxAtoms - 1; ++a)
#pragma unroll
                    for (int b2 = 0; b2 < kMaxAtoms - 1; ++b2)
                        if (a < na - 1 && b2 < na - 1) f[o++] = dot(ui[a], uj[b2]);
            }
            // feature.py:493-512: normals of consecutive UN-normalised bond vectors (eps 1e-6), index a*(nh-2) + b
            {
                V3 ni[kMaxAtoms - 2], nj[kMaxAtoms - 2];
#pragma unroll
                for (int a = 0; a < kMaxAtoms - 2; ++a) {
                    ni[a] = unit_eps(cross(vi[a], vi[a + 1]), 1e-6f);
                    nj[a] = unit_eps(cross(vj[a], vj[a + 1]), 1e-6f);
                }
#pragma unroll
                for (int a = 0; a < kMaxAtoms - 2; ++a)
#pragma unroll
                    for (int b2 = 0; b2 < kMaxAtoms - 2; ++b2)
                        if (a < nh - 2 && b2 < nh - 2) f[o++] = dot(ni[a], nj[b2]);
            }
        }
    }
    __syncthreads();
    const size_t n_here = total - e0 < (size_t)kEdgeTile ? total - e0 : (size_t)kEdgeTile;
    float* dst = edge_raw + e0 * edge_dim;
    for (size_t q = threadIdx.x; q < n_here * edge_dim; q += kEdgeTile) {
        const int r = (int)(q / edge_dim), cidx = (int)(q % edge_dim);
        dst[q] = s_feat[(size_t)r * stride + cidx];
    }
}

// --------------------------------------------------------------------------------
// GraphNormalization (a-13), ONE launch: a thread-block cluster per structure.
//   pass 1: every CTA of the cluster sums x = (a + b) and x^2 over its share of the structure's VALID rows per channel
//           (fp64, fixed order); the partials are exchanged through distributed shared memory (cluster.map_shared_rank)
//           and every CTA adds them in rank order, so the statistics are deterministic (the two-launch version of
//           round 1 used fp64 atomics into a global workspace: 3 graph nodes per norm, 31 % of the rnampnn step);
//   pass 2: mean = S1/n, var = (S2 - n mean^2 + n_pad mean^2)/n  [padded rows add mean^2 each, functional.py:33-38],
//           out = ((x - mean)/sqrt(var + 1e-6) * scale + shift) * mask over the same rows (re-read from L2; short
//           shares stay in registers).
// x = a + b fuses the residual adds that precede every GraphNorm in the reference (mpnn.py:222, functional.py:165).
// The two-launch kernels below remain for structures longer than one cluster covers well and as the checked fallback
// when the cluster launch is refused.
// --------------------------------------------------------------------------------
constexpr int kGnRows = 32;       // rows per block (both kernels)

__global__ void __launch_bounds__(256) graph_norm_stats_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                              const float* __restrict__ mask, int N,
                                                              double* __restrict__ stats) {
    __shared__ double s_part[2][128];
    __shared__ int s_cnt;
    const int g = blockIdx.y, c = threadIdx.x & 127, rh = threadIdx.x >> 7;
    const int r0 = blockIdx.x * kGnRows + rh * 16;
    const size_t base = (size_t)g * N;
    // 16 rows per thread, every load issued before the first use
    float x[16];
    unsigned ok = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int r = r0 + q;
        x[q] = 0.f;
        if (r < N && mask[base + r] != 0.0f) {
            ok |= 1u << q;
            x[q] = a[(base + r) * 128 + c];
        }
    }
    if (b) {
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (ok >> q & 1) x[q] += b[(base + r0 + q) * 128 + c];
    }
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        s1 += (double)x[q];                                   // masked rows hold exact zeros
        s2 += (double)x[q] * (double)x[q];
    }
    if (rh == 1) { s_part[0][c] = s1; s_part[1][c] = s2; }
    if (threadIdx.x == 128) s_cnt = __popc(ok);
    __syncthreads();
    if (rh == 0) {
        s1 += s_part[0][c];
        s2 += s_part[1][c];
        const int n = __popc(ok) + s_cnt;
        if (n == 0) return;
        double* st = stats + (size_t)g * 257;
        atomicAdd(st + c, s1);
        atomicAdd(st + 128 + c, s2);
        if (c == 0) atomicAdd(st + 256, (double)n);
    }
}

__global__ void __launch_bounds__(256) graph_norm_apply_kernel(const float* __restrict__ a, float* __restrict__ b,
                                                              const float* __restrict__ mask, int N, int n_total,
                                                              const double* __restrict__ stats,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int zero_b,
                                                              float* __restrict__ out, uint2* __restrict__ out16) {
    __shared__ __align__(16) float s_mean[128], s_mul[128];     // mean, standard deviation
    const int g = blockIdx.y;
    if (threadIdx.x < 128) {                                   // per-channel statistics of this graph, once per block
        const int c = threadIdx.x;
        const double* st = stats + (size_t)g * 257;
        const double cnt = st[256];
        const double n = cnt > 0.0 ? cnt : 1.0;
        const double mean = st[c] / n;
        double var = (st[128 + c] - n * mean * mean + ((double)n_total - cnt) * mean * mean) / n;
        var = var > 0.0 ? var : 0.0;
        const float sd = sqrtf((float)var + 1e-6f);
        s_mean[c] = (float)mean;
        s_mul[c] = sd;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float4 mu = reinterpret_cast<const float4*>(s_mean)[lane];
    const float4 iv = reinterpret_cast<const float4*>(s_mul)[lane];
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + lane);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + lane);
    const size_t base = (size_t)g * N;
    float4 x[4], y[4];
    bool live[4], inside[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = blockIdx.x * kGnRows + w * 4 + q;
        inside[q] = r < N;
        live[q] = inside[q] && mask[base + r] != 0.0f;
        x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        y[q] = x[q];
        if (live[q]) x[q] = reinterpret_cast<const float4*>(a + (base + r) * 128)[lane];
        if (live[q] && b) y[q] = reinterpret_cast<const float4*>(b + (base + r) * 128)[lane];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (!inside[q]) continue;
        const size_t row = base + blockIdx.x * kGnRows + w * 4 + q;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live[q]) {
            // (x - mean) / sd * scale + shift, the division as in the reference (functional.py:42-43)
            o.x = (x[q].x + y[q].x - mu.x) / iv.x * sc.x + sh.x;
            o.y = (x[q].y + y[q].y - mu.y) / iv.y * sc.y + sh.y;
            o.z = (x[q].z + y[q].z - mu.z) / iv.z * sc.z + sh.z;
            o.w = (x[q].w + y[q].w - mu.w) / iv.w * sc.w + sh.w;
        }
        reinterpret_cast<float4*>(out + row * 128)[lane] = o;
        if (out16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            out16[row * 32 + lane] = make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
        }
        if (b && zero_b) reinterpret_cast<float4*>(b + row * 128)[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Fused GraphNorm: grid (CL, B), cluster (CL, 1, 1); CTA `rank` owns rows [rank * share, (rank + 1) * share) of structure g.
// 256 threads, thread = channel c = tid & 127 x row phase tid >> 7 (rows of a thread are 2 apart).  KEEP: share <= 32, the
// thread's 16 rows are loaded once, all loads in flight, and stay in registers for the normalisation (one memory round
// trip per norm); otherwise the rows are streamed twice (L2-resident the second time).
template <bool KEEP>
__global__ void __launch_bounds__(256) graph_norm_fused_kernel(const float* __restrict__ a, float* __restrict__ b,
                                                              const float* __restrict__ mask, int N, int n_total, int share,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int zero_b,
                                                              float* __restrict__ out, uint16_t* __restrict__ out16) {
    __shared__ double s_sum[2][128];           // this CTA's partial sums: read by the whole cluster
    __shared__ double s_half[2][128];
    __shared__ int s_n, s_nh;
    __shared__ float s_mean[128], s_sd[128];
    unsigned rank, cl;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(cl));
    const int g = blockIdx.y, c = threadIdx.x & 127, rh = threadIdx.x >> 7;
    const size_t base = (size_t)g * N;
    const int r_begin = (int)rank * share, r_end = min(N, r_begin + share);

    // ---- pass 1: per-channel sums over this CTA's valid rows (fp64, row order)
    double s1 = 0.0, s2 = 0.0;
    int n = 0;
    float keep[KEEP ? 16 : 1];
    unsigned keep_ok = 0;
    if constexpr (KEEP) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int r = r_begin + rh + 2 * q;
            keep[q] = 0.f;
            if (r < r_end && mask[base + r] != 0.0f) {
                keep_ok |= 1u << q;
                keep[q] = a[(base + r) * 128 + c];
            }
        }
        if (b) {
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (keep_ok >> q & 1) keep[q] += b[(base + r_begin + rh + 2 * q) * 128 + c];
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            s1 += (double)keep[q];                               // masked rows hold exact zeros
            s2 += (double)keep[q] * (double)keep[q];
        }
        n = __popc(keep_ok);
    } else {
        for (int r0 = r_begin + rh; r0 < r_end; r0 += 16) {     // eight rows in flight per thread
            float x[8];
            bool ok[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = r0 + 2 * q;
                ok[q] = r < r_end && mask[base + r] != 0.0f;
                x[q] = ok[q] ? a[(base + r) * 128 + c] : 0.f;
            }
            if (b) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (ok[q]) x[q] += b[(base + r0 + 2 * q) * 128 + c];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                s1 += (double)x[q];
                s2 += (double)x[q] * (double)x[q];
                n += ok[q] ? 1 : 0;
            }
        }
    }
    if (rh == 1) { s_half[0][c] = s1; s_half[1][c] = s2; }
    if (threadIdx.x == 128) s_nh = n;
    __syncthreads();
    if (rh == 0) {
        s_sum[0][c] = s1 + s_half[0][c];
        s_sum[1][c] = s2 + s_half[1][c];
        if (c == 0) s_n = n + s_nh;
    }
    // ---- exchange: all partials of the cluster, added in rank order by every CTA
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rh == 0) {
        double t1 = 0.0, t2 = 0.0;
        int cnt_i = 0;
        for (unsigned k = 0; k < cl; ++k) {
            uint32_t p1, p2, pn;
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(p1) : "r"((uint32_t)__cvta_generic_to_shared(&s_sum[0][c])), "r"(k));
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(p2) : "r"((uint32_t)__cvta_generic_to_shared(&s_sum[1][c])), "r"(k));
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(pn) : "r"((uint32_t)__cvta_generic_to_shared(&s_n)), "r"(k));
            double v1, v2;
            int vn;
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v1) : "r"(p1));
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v2) : "r"(p2));
            asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(vn) : "r"(pn));
            t1 += v1; t2 += v2; cnt_i += vn;
        }
        const double cnt = (double)cnt_i;
        const double nn = cnt > 0.0 ? cnt : 1.0;
        const double mean = t1 / nn;
        double var = (t2 - nn * mean * mean + ((double)n_total - cnt) * mean * mean) / nn;
        var = var > 0.0 ? var : 0.0;
        s_mean[c] = (float)mean;
        s_sd[c] = sqrtf((float)var + 1e-6f);
    }
    // nobody may leave (and release its shared memory) before every CTA has read the partials
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    __syncthreads();

    // ---- pass 2: normalise this CTA's rows; (x - mean) / sd * scale + shift, the division as in the reference
    // (functional.py:42-43).  A row is written by 128 consecutive threads (coalesced 512 bytes).
    const float mu = s_mean[c], sd = s_sd[c], sc = __ldg(scale + c), sh = __ldg(shift + c);
    auto emit = [&](int r, bool live, float x) {
        const size_t row = base + r;
        const float o = live ? (x - mu) / sd * sc + sh : 0.f;
        out[row * 128 + c] = o;
        if (out16) {
            const float o_hi = __shfl_down_sync(0xffffffffu, o, 1);
            if ((c & 1) == 0) {
                const __nv_bfloat162 pk = __floats2bfloat162_rn(o, o_hi);
                reinterpret_cast<uint32_t*>(out16)[(row * 128 + c) >> 1] = *reinterpret_cast<const uint32_t*>(&pk);
            }
        }
        if (b && zero_b) b[row * 128 + c] = 0.f;
    };
    if constexpr (KEEP) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int r = r_begin + rh + 2 * q;
            if (r < r_end) emit(r, (keep_ok >> q & 1) != 0, keep[q]);      // uniform per warp (one row per 128 threads)
        }
    } else {
        for (int r0 = r_begin + rh; r0 < r_end; r0 += 16) {
            float x[8];
            bool ok[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int r = r0 + 2 * q;
                ok[q] = r < r_end && mask[base + r] != 0.0f;
                x[q] = ok[q] ? a[(base + r) * 128 + c] : 0.f;
            }
            if (b) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (ok[q]) x[q] += b[(base + r0 + 2 * q) * 128 + c];
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r0 + 2 * q < r_end) emit(r0 + 2 * q, ok[q], x[q]);
        }
    }
}

// --------------------------------------------------------------------------------
// Self-attention over the valid rows of each structure (a-15).  One thread per query row, one block per
// (128 queries, head, structure); keys/values of the head stream through shared memory in tiles of 128 rows,
// online softmax in chunks of 8 keys.  The reference pads every structure to 4500 rows and masks the padded
// keys (functional.py:161-168); only the L_b valid keys carry weight, so only those are visited.
//   qkv [B*N, 3E] = in_proj(x): q | k | v, head h at columns h*HD.  out [B*N, E] (padded query rows -> 0).
// --------------------------------------------------------------------------------
constexpr int kAttnTile = 128;

// packed fp32x2 arithmetic (FFMA2 / FMUL2 on sm_100): one instruction per two channels
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t"
        ".reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Scores are kept in the log2 domain (q is pre-scaled by log2(e)/sqrt(HD)), so the softmax needs one MUFU.EX2 per key
// instead of a full expf; dot products and the P.V update run on packed fp32x2 FMAs.  ex2.approx is accurate to 2^-22
// relative, far inside the 1e-4 logit tolerance of the fp32 arm.
template <int HD>
__global__ void __launch_bounds__(kAttnTile) res_attention_kernel(const float* __restrict__ qkv,
                                                                 const float* __restrict__ mask, int N, int E,
                                                                 float* __restrict__ out) {
    __shared__ __align__(16) float s_k[kAttnTile][HD];
    __shared__ __align__(16) float s_v[kAttnTile][HD];
    __shared__ float s_ok[kAttnTile];
    const int b = blockIdx.z, h = blockIdx.y, t = threadIdx.x;
    const int qi = blockIdx.x * kAttnTile + t;
    const size_t base = (size_t)b * N;
    const bool q_ok = qi < N && mask[base + qi] != 0.0f;
    float2 q[HD / 2], acc[HD / 2];
    const float qs = 1.4426950408889634f / sqrtf((float)HD);
#pragma unroll
    for (int d = 0; d < HD / 2; ++d) { q[d] = make_float2(0.f, 0.f); acc[d] = make_float2(0.f, 0.f); }
    if (q_ok) {
        const float4* qp = reinterpret_cast<const float4*>(qkv + (base + qi) * 3 * E + h * HD);
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 v = qp[d4];
            q[2 * d4] = make_float2(v.x * qs, v.y * qs);
            q[2 * d4 + 1] = make_float2(v.z * qs, v.w * qs);
        }
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < N; k0 += kAttnTile) {
        __syncthreads();
        {
            const int kj = k0 + t;
            const bool ok = kj < N && mask[base + kj] != 0.0f;
            s_ok[t] = ok ? 1.f : 0.f;
            if (ok) {
                const float4* kp = reinterpret_cast<const float4*>(qkv + (base + kj) * 3 * E + E + h * HD);
                const float4* vp = reinterpret_cast<const float4*>(qkv + (base + kj) * 3 * E + 2 * E + h * HD);
#pragma unroll
                for (int d4 = 0; d4 < HD / 4; ++d4) {
                    reinterpret_cast<float4*>(s_k[t])[d4] = kp[d4];
                    reinterpret_cast<float4*>(s_v[t])[d4] = vp[d4];
                }
            }
        }
        __syncthreads();
        const int kn = min(kAttnTile, N - k0);
        for (int c0 = 0; c0 < kn; c0 += 8) {
            float s[8];
            float cmax = -INFINITY;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int kk = c0 + u;
                s[u] = -INFINITY;
                if (kk < kn && s_ok[kk] != 0.f) {               // uniform across the block
                    float2 da = make_float2(0.f, 0.f), db = make_float2(0.f, 0.f);
#pragma unroll
                    for (int d4 = 0; d4 < HD / 4; ++d4) {
                        const float4 kv = reinterpret_cast<const float4*>(s_k[kk])[d4];
                        da = ffma2(q[2 * d4], make_float2(kv.x, kv.y), da);
                        db = ffma2(q[2 * d4 + 1], make_float2(kv.z, kv.w), db);
                    }
                    const float d0 = (da.x + da.y) + (db.x + db.y);
                    s[u] = d0;
                    cmax = fmaxf(cmax, d0);
                }
            }
            if (cmax == -INFINITY) continue;
            const float m_new = fmaxf(m, cmax);
            const float corr = ex2_fast(m - m_new);             // 2^-inf = 0 on the first chunk
            l *= corr;
            const float2 c2 = make_float2(corr, corr);
#pragma unroll
            for (int d = 0; d < HD / 2; ++d) acc[d] = fmul2(acc[d], c2);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (s[u] == -INFINITY) continue;
                const float p = ex2_fast(s[u] - m_new);
                l += p;
                const float2 p2 = make_float2(p, p);
#pragma unroll
                for (int d4 = 0; d4 < HD / 4; ++d4) {
                    const float4 vv = reinterpret_cast<const float4*>(s_v[c0 + u])[d4];
                    acc[2 * d4] = ffma2(p2, make_float2(vv.x, vv.y), acc[2 * d4]);
                    acc[2 * d4 + 1] = ffma2(p2, make_float2(vv.z, vv.w), acc[2 * d4 + 1]);
                }
            }
            m = m_new;
        }
    }
    if (qi < N) {
        float* op = out + (base + qi) * E + h * HD;
        const float inv = (q_ok && l > 0.f) ? 1.0f / l : 0.f;
        // 32-byte stores: full sectors per instruction (a thread's head slice is HD * 4 = 32..128 contiguous bytes)
#pragma unroll
        for (int d8 = 0; d8 < HD / 8; ++d8) {
            const float2 a0 = acc[4 * d8], a1 = acc[4 * d8 + 1], a2 = acc[4 * d8 + 2], a3 = acc[4 * d8 + 3];
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op + 8 * d8), "f"(a0.x * inv),
                         "f"(a0.y * inv), "f"(a1.x * inv), "f"(a1.y * inv), "f"(a2.x * inv), "f"(a2.y * inv),
                         "f"(a3.x * inv), "f"(a3.y * inv)
                         : "memory");
        }
    }
}

// rows of x [rows, width] times a per-row factor (logits *= mask, functional.py:88)
__global__ void __launch_bounds__(256) scale_rows_kernel(float* __restrict__ x, const float* __restrict__ w, size_t rows,
                                                        int width) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= rows * width) return;
    x[i] *= w[i / width];
}

}  // namespace rmp

using namespace rmp;

extern "C" int rmp_res_edge_rows(const int64_t* edge_index, const float* mask, int B, int N, int K, int32_t* ctr_row,
                                 int32_t* nbr_row, float* edge_keep, float* edge_w, void* stream) {
    RMP_REQUIRE(edge_index && mask && ctr_row && nbr_row && edge_keep && edge_w, RMP_E_ARG, "rmp_res_edge_rows: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && K > 0, RMP_E_ARG, "rmp_res_edge_rows: bad sizes");
    const size_t rows = (size_t)B * N;
    res_edge_rows_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(edge_index, mask, B, N, K, ctr_row,
                                                                                          nbr_row, edge_keep, edge_w);
    return check_launch("rmp_res_edge_rows");
}

extern "C" int rmp_res_raw_features(const float* coords, const float* mask, const int64_t* edge_index, int B, int N,
                                    int A, int K, int n_inside_dist, int n_inside_angle, int n_inside_dihedral,
                                    int n_cross_dist, int n_cross_angle, int n_cross_dihedral, float* raw,
                                    float* edge_raw, void* stream) {
    RMP_REQUIRE(coords && mask, RMP_E_ARG, "rmp_res_raw_features: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && A >= 4 && A <= kMaxAtoms, RMP_E_SHAPE, "rmp_res_raw_features: A=%d outside [4, %d]", A,
                kMaxAtoms);
    cudaStream_t st = (cudaStream_t)stream;
    int launches = 0;
    if (raw) {
        const int nd = n_inside_dist, na = n_inside_angle, nh = n_inside_dihedral;
        RMP_REQUIRE(nd >= 2 && nd <= A && na >= 3 && na <= A && nh >= 4 && nh <= A, RMP_E_SHAPE,
                    "rmp_res_raw_features: inside atom counts (%d,%d,%d) need 2<=d, 3<=a, 4<=h, all <= A", nd, na, nh);
        const int raw_dim = nd * (nd - 1) / 2 + na - 2 + nh - 3;
        const int rows = B * N;
        res_inside_kernel<<<ceil_div(rows, 128), 128, 0, st>>>(coords, mask, rows, A, nd, na, nh, raw_dim, raw);
        ++launches;
    }
    if (edge_raw) {
        const int nd = n_cross_dist, na = n_cross_angle, nh = n_cross_dihedral;
        RMP_REQUIRE(edge_index && K > 0, RMP_E_ARG, "rmp_res_raw_features: edge_index missing");
        RMP_REQUIRE(nd >= 2 && nd <= A && na >= 3 && na <= A && nh >= 4 && nh <= A, RMP_E_SHAPE,
                    "rmp_res_raw_features: cross atom counts (%d,%d,%d) need 2<=d, 3<=a, 4<=h, all <= A", nd, na, nh);
        const int edge_dim = nd * nd + (na - 1) * (na - 1) + (nh - 2) * (nh - 2);
        const size_t total = (size_t)B * N * K;
        const size_t smem = (size_t)kEdgeTile * (edge_dim + 1) * sizeof(float);
        res_cross_kernel<<<(unsigned)((total + kEdgeTile - 1) / kEdgeTile), kEdgeTile, smem, st>>>(
            coords, mask, edge_index, B, N, A, K, nd, na, nh, edge_dim, edge_raw);
        ++launches;
    }
    return check_launch("rmp_res_raw_features", launches);
}

extern "C" int rmp_graph_norm(const float* a, float* b, const float* mask, int B, int N, int n_total,
                              const float* scale, const float* shift, int zero_b, double* stats_ws, float* out,
                              void* out_bf16, void* stream) {
    RMP_REQUIRE(a && mask && scale && shift && stats_ws && out, RMP_E_ARG, "rmp_graph_norm: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && n_total >= N, RMP_E_ARG, "rmp_graph_norm: bad sizes (n_total must be >= N)");
    RMP_REQUIRE(aligned16(a) && aligned16(out) && (!b || aligned16(b)), RMP_E_ALIGN, "rmp_graph_norm: 16-byte alignment");
    RMP_REQUIRE(B <= 65535, RMP_E_SHAPE, "rmp_graph_norm: B=%d exceeds 65535 structures per call", B);
    cudaStream_t st = (cudaStream_t)stream;
    // one cluster of up to 8 CTAs per structure (32 rows or more per CTA); out may alias a (the passes of a CTA touch its
    // own rows only, and pass 2 starts after every CTA of the cluster has finished pass 1)
    static const bool two_launch = getenv("RMP_GRAPH_NORM_TWO_LAUNCH") != nullptr;
    if (!two_launch) {
        int cl = 1;
        while (cl < 8 && cl * kGnRows < N) cl <<= 1;
        const int share = ceil_div(N, cl);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(cl, B);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, share <= 32 ? graph_norm_fused_kernel<true> : graph_norm_fused_kernel<false>, a, b,
                                            mask, N, n_total, share, scale, shift, zero_b, out,
                                            reinterpret_cast<uint16_t*>(out_bf16));
        if (le == cudaSuccess) return check_launch("rmp_graph_norm");
        cudaGetLastError();          // cluster launch refused: fall through to the two-launch version
    }
    cudaError_t e = cudaMemsetAsync(stats_ws, 0, (size_t)B * 257 * sizeof(double), st);
    if (e != cudaSuccess) { set_error("rmp_graph_norm: memset: %s", cudaGetErrorString(e)); return (int)e; }
    graph_norm_stats_kernel<<<dim3(ceil_div(N, kGnRows), B), 256, 0, st>>>(a, b, mask, N, stats_ws);
    graph_norm_apply_kernel<<<dim3(ceil_div(N, kGnRows), B), 256, 0, st>>>(a, b, mask, N, n_total, stats_ws, scale, shift,
                                                                            zero_b, out, reinterpret_cast<uint2*>(out_bf16));
    return check_launch("rmp_graph_norm", 2);
}

extern "C" int rmp_res_attention(const float* qkv, const float* mask, int B, int N, int E, int heads, float* out,
                                 void* stream) {
    RMP_REQUIRE(qkv && mask && out, RMP_E_ARG, "rmp_res_attention: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && heads > 0 && E > 0 && E % heads == 0, RMP_E_ARG, "rmp_res_attention: bad sizes");
    RMP_REQUIRE(aligned16(qkv) && aligned32(out), RMP_E_ALIGN, "rmp_res_attention: qkv 16-byte, out 32-byte alignment (256-bit stores)");
    RMP_REQUIRE(B <= 65535 && heads <= 65535, RMP_E_SHAPE, "rmp_res_attention: grid too large");
    const int hd = E / heads;
    dim3 grid(ceil_div(N, kAttnTile), heads, B);
    cudaStream_t st = (cudaStream_t)stream;
    switch (hd) {
        case 8: res_attention_kernel<8><<<grid, kAttnTile, 0, st>>>(qkv, mask, N, E, out); break;
        case 16: res_attention_kernel<16><<<grid, kAttnTile, 0, st>>>(qkv, mask, N, E, out); break;
        case 32: res_attention_kernel<32><<<grid, kAttnTile, 0, st>>>(qkv, mask, N, E, out); break;
        default:
            set_error("rmp_res_attention: head dim %d not in {8,16,32}", hd);
            return RMP_E_SHAPE;
    }
    return check_launch("rmp_res_attention");
}

extern "C" int rmp_scale_rows(float* x, const float* w, int64_t rows, int width, void* stream) {
    RMP_REQUIRE(x && w && rows >= 0 && width > 0, RMP_E_ARG, "rmp_scale_rows: bad arguments");
    if (rows == 0) return RMP_OK;
    const size_t total = (size_t)rows * width;
    scale_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, w, (size_t)rows, width);
    return check_launch("rmp_scale_rows");
}
