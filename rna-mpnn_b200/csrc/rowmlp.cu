// fp32 CUDA-core path: fused row-tile MLP (chained Linear + exact GELU) with
// gather/concat input and fused epilogues, add+LayerNorm, and the readout head.
//
// This is the RMP_PREC_FP32 arm of the layer path (1e-4 logit parity with the
// reference's CPU fp32 path) and the embedding / readout stages of both arms.
// Reference semantics: rdesign/model/feature.py:223-224 (embeddings + Normalize),
// rdesign/model/rdesign.py:85 + rdesign/model/mpnn.py:31-38 (layer),
// rdesign/model/functional.py:79-96,123-127 (Normalize, Readout).
//
// Activations live in shared memory feature-major, act[c][r] with row stride
// ROWS+4, so the register-tiled GEMM reads 4 consecutive rows as one float4 and
// the next stage consumes the previous stage's output in place.
#include "common.cuh"

namespace rmp {

constexpr int kThreads = 256;
constexpr int kKC = 32;          // k-chunk of the weight pipeline
constexpr int kNB = 128;         // output columns per pass

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct MlpSmemPlan {
    int cap0, cap1;              // channel capacity of the two activation buffers
};

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// EXT = false is the instantiation of the rdesign path (no per-row scale, no in-place residual epilogue): keeping the
// rnampnn extensions out of it preserves its register allocation and FFMA-loop schedule (the merged kernel ran the
// message MLP 15 % slower).
template <int ROWS, bool EXT>
__global__ void __launch_bounds__(kThreads) row_mlp_kernel(const rmp_mlp_desc d, const int cap0, const int cap1) {
    constexpr int RS = ROWS + 4;                 // smem row stride (floats)
    constexpr int TR = ROWS / 16;                // rows per thread tile: 4 (ROWS=64) or 2 (ROWS=32)
    extern __shared__ __align__(16) float smem[];
    float* buf0 = smem;
    float* buf1 = buf0 + (size_t)cap0 * RS;
    float* wbuf = buf1 + (size_t)cap1 * RS;      // [2][kKC][kNB]
    int* s_seg = reinterpret_cast<int*>(wbuf + 2 * kKC * kNB);

    const int t = threadIdx.x;
    const int nrows = d.rows_dev ? min(d.rows_dev[d.rows_slot], d.rows_cap) : d.rows_cap;
    const int r0 = blockIdx.x * ROWS;
    if (r0 >= nrows) return;
    const int rv = min(ROWS, nrows - r0);

    // ---------------- stage the (gathered, concatenated) input, feature-major ----------------
    {
        const int r = t % ROWS;
        const int part = t / ROWS;
        constexpr int PARTS = kThreads / ROWS;
        int coff = 0;
        for (int s = 0; s < d.num_segs; ++s) {
            const int wdt = d.seg_width[s];
            const float* src = nullptr;
            if (r < rv) {
                const int row = d.seg_idx[s] ? d.seg_idx[s][r0 + r] : (r0 + r);
                src = d.seg_src[s] + (size_t)row * wdt;
            }
            if ((wdt & 3) == 0) {
                for (int c4 = part; c4 < wdt / 4; c4 += PARTS) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (src) v = __ldg(reinterpret_cast<const float4*>(src) + c4);
                    float* dst = buf0 + (size_t)(coff + c4 * 4) * RS + r;
                    dst[0] = v.x; dst[RS] = v.y; dst[2 * RS] = v.z; dst[3 * RS] = v.w;
                }
            } else {
                for (int c = part; c < wdt; c += PARTS) buf0[(size_t)(coff + c) * RS + r] = src ? __ldg(src + c) : 0.f;
            }
            coff += wdt;
        }
        const int kpad = round_up(coff, kKC);
        for (int c = coff + part; c < kpad; c += PARTS) buf0[(size_t)c * RS + r] = 0.f;
        if (d.epilogue == RMP_EPI_SEGSUM && t < ROWS) s_seg[t] = (t < rv) ? d.seg_id[r0 + t] : -1;
    }
    __syncthreads();

    // ---------------- chained Linear (+GELU) stages ----------------
    const int tx = t & 15, ty = t >> 4;
    float* in = buf0;
    float* out = buf1;
    for (int s = 0; s < d.num_stages; ++s) {
        const int kpad = round_up(d.k_in[s], kKC);
        const int nout = d.n_out[s];
        const float* __restrict__ Wt = d.Wt[s];
        const int nchunks = kpad / kKC;
        for (int nb = 0; nb < nout; nb += kNB) {
            // accumulators as fp32 pairs: the inner product runs on packed fma.rn.f32x2 (same rounding as fmaf, half the
            // issue slots -- the loop is issue-bound, not FMA-pipe-bound)
            float2 acc2[TR][4];
#pragma unroll
            for (int i = 0; i < TR; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);

            auto prefetch = [&](int chunk, int stage) {
                float* dst = wbuf + stage * kKC * kNB;
                const float* src = Wt + (size_t)chunk * kKC * nout + nb;
#pragma unroll
                for (int q = 0; q < (kKC * kNB / 4) / kThreads; ++q) {
                    const int idx = q * kThreads + t;        // 16-byte unit
                    const int kk = idx / (kNB / 4), c4 = idx % (kNB / 4);
                    cp_async16(dst + kk * kNB + c4 * 4, src + (size_t)kk * nout + c4 * 4);
                }
                cp_async_commit();
            };
            prefetch(0, 0);
            for (int ch = 0; ch < nchunks; ++ch) {
                if (ch + 1 < nchunks) {
                    prefetch(ch + 1, (ch + 1) & 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                const float* w = wbuf + (ch & 1) * kKC * kNB;
                const float* a_base = in + (size_t)ch * kKC * RS + ty * TR;
#pragma unroll 8
                for (int kk = 0; kk < kKC; ++kk) {
                    float a[TR];
                    if (TR == 4) {
                        const float4 av = *reinterpret_cast<const float4*>(a_base + kk * RS);
                        a[0] = av.x; a[1] = av.y; a[2 % TR] = av.z; a[3 % TR] = av.w;
                    } else {
                        const float2 av = *reinterpret_cast<const float2*>(a_base + kk * RS);
                        a[0] = av.x; a[1] = av.y;
                    }
                    const float4 b0 = *reinterpret_cast<const float4*>(w + kk * kNB + tx * 4);
                    const float4 b1 = *reinterpret_cast<const float4*>(w + kk * kNB + 64 + tx * 4);
                    const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                          make_float2(b1.z, b1.w)};
#pragma unroll
                    for (int i = 0; i < TR; ++i) {
                        const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc2[i][j] = ffma2(aa, bv[j], acc2[i][j]);
                    }
                }
                __syncthreads();
            }
            // bias + activation, write feature-major for the next stage / the epilogue
            const float* __restrict__ bias = d.bias[s];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = nb + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                const float bj = bias ? __ldg(bias + col) : 0.f;
                float v[TR];
#pragma unroll
                for (int i = 0; i < TR; ++i) {
                    v[i] = ((j & 1) ? acc2[i][j >> 1].y : acc2[i][j >> 1].x) + bj;
                    if (d.gelu[s]) v[i] = gelu_erf(v[i]);
                }
                float* dst = out + (size_t)col * RS + ty * TR;
                if (TR == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2 % TR], v[3 % TR]);
                else *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
            }
        }
        __syncthreads();
        float* tmp = in; in = out; out = tmp;
    }
    const float* fin = in;                                   // [n][RS]
    const int n = d.n_out[d.num_stages - 1];
    const int lane = t & 31, w = t >> 5;

    // ---------------- epilogues ----------------
    if (d.epilogue == RMP_EPI_SEGSUM) {
        // n == 128.  thread c sums its channel over each run of equal seg ids; one
        // atomic per (run, channel) onto the pre-zeroed destination.
        if (t < 128) {
            float run = 0.f;
            int cur = s_seg[0];
            for (int r = 0; r < rv; ++r) {
                const int sg = s_seg[r];
                if (sg != cur) {
                    atomicAdd(d.out + (size_t)cur * 128 + t, run * d.alpha);
                    run = 0.f;
                    cur = sg;
                }
                if (EXT && d.row_scale) run += fin[(size_t)t * RS + r] * __ldg(d.row_scale + r0 + r);
                else run += fin[(size_t)t * RS + r];
            }
            atomicAdd(d.out + (size_t)cur * 128 + t, run * d.alpha);
        }
        return;
    }
    // row-wise epilogues: one warp per row, lanes stride over channels
    for (int r = w; r < rv; r += kThreads / 32) {
        float* orow = d.out + (size_t)(r0 + r) * n;
        if (d.epilogue == RMP_EPI_STORE) {
            if (EXT && d.row_scale) {
                const float rs = __ldg(d.row_scale + r0 + r);
                for (int c = lane; c < n; c += 32) orow[c] = fin[(size_t)c * RS + r] * rs;
            } else {
                for (int c = lane; c < n; c += 32) orow[c] = fin[(size_t)c * RS + r];
            }
            continue;
        }
        if (EXT && d.epilogue == RMP_EPI_RESID_ADD) {
            const float* rrow = d.resid + (size_t)(r0 + r) * n;
            for (int c = lane; c < n; c += 32) orow[c] = rrow[c] + fin[(size_t)c * RS + r];
            continue;
        }
        // NORMALIZE / RESID_LN operate on n == 128 channels: 4 per lane
        float y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            y[q] = fin[(size_t)c * RS + r];
            if (d.epilogue == RMP_EPI_RESID_LN) y[q] += __ldg(d.resid + (size_t)(r0 + r) * 128 + c);
        }
        const float mu = warp_sum(y[0] + y[1] + y[2] + y[3]) * (1.f / 128.f);
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) { y[q] -= mu; ss += y[q] * y[q]; }
        ss = warp_sum(ss);
        float inv;
        if (d.epilogue == RMP_EPI_NORMALIZE) {
            const float sigma = sqrtf(ss * (1.f / 127.f) + 1e-6f);     // unbiased variance
            inv = 1.f / (sigma + 1e-6f);
        } else {
            inv = 1.f / sqrtf(ss * (1.f / 128.f) + 1e-5f);             // nn.LayerNorm
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            orow[c] = __ldg(d.gain + c) * (y[q] * inv) + __ldg(d.beta + c);
        }
    }
}

// out = LayerNorm(a + b) over 128 channels, one warp per row, float4 per lane.
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ a, float* __restrict__ b,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps,
                                                           const int32_t* __restrict__ rows_dev, int rows_slot,
                                                           int rows_cap, int zero_b, int variant,
                                                           float* __restrict__ out) {
    const int nrows = rows_dev ? min(rows_dev[rows_slot], rows_cap) : rows_cap;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const int lane = threadIdx.x & 31;
    float4 x = reinterpret_cast<const float4*>(a + (size_t)r * 128)[lane];
    if (b) {
        float4* bp = reinterpret_cast<float4*>(b + (size_t)r * 128) + lane;
        const float4 y = *bp;
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
        if (zero_b) *bp = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float mu = warp_sum(x.x + x.y + x.z + x.w) * (1.f / 128.f);
    x.x -= mu; x.y -= mu; x.z -= mu; x.w -= mu;
    const float ss = warp_sum(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w);
    // variant 0: nn.LayerNorm (biased variance); 1: reference Normalize (unbiased, eps inside and outside)
    const float inv = variant ? 1.f / (sqrtf(ss * (1.f / 127.f) + eps) + eps) : 1.f / sqrtf(ss * (1.f / 128.f) + eps);
    const float4 g = reinterpret_cast<const float4*>(gamma)[lane];
    const float4 bt = reinterpret_cast<const float4*>(beta)[lane];
    reinterpret_cast<float4*>(out + (size_t)r * 128)[lane] =
        make_float4(g.x * (x.x * inv) + bt.x, g.y * (x.y * inv) + bt.y, g.z * (x.z * inv) + bt.z,
                    g.w * (x.w * inv) + bt.w);
}

// Philox-4x32-10 counter RNG: one draw per (seed, node).
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t ctr0) {
    uint32_t c[4] = {ctr0, 0u, 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return (float)(c[0] >> 8) * (1.0f / 16777216.0f);          // [0,1)
}

// Final Linear(k_in -> 4) + softmax + argmax / categorical sample.  One warp per node.
__global__ void __launch_bounds__(256) readout_head_kernel(const float* __restrict__ h, int k_in,
                                                          const float* __restrict__ W, const float* __restrict__ b,
                                                          float temperature, uint64_t seed,
                                                          const int32_t* __restrict__ rows_dev, int rows_slot,
                                                          int rows_cap, const float* __restrict__ row_mask,
                                                          float* __restrict__ logits,
                                                          float* __restrict__ probs, int64_t* __restrict__ samples) {
    const int nrows = rows_dev ? min(rows_dev[rows_slot], rows_cap) : rows_cap;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const int lane = threadIdx.x & 31;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < k_in; c += 32) {
        const float x = h[(size_t)r * k_in + c];
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = fmaf(x, __ldg(W + o * k_in + c), acc[o]);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] = warp_sum(acc[o]) + __ldg(b + o);
    if (lane != 0) return;
    if (row_mask) {                                            // logits *= mask (rnampnn/model/functional.py:88)
        const float mk = row_mask[r];
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] *= mk;
    }
    float p[4];
    const float m = fmaxf(fmaxf(acc[0], acc[1]), fmaxf(acc[2], acc[3]));
    float z = 0.f;
#pragma unroll
    for (int o = 0; o < 4; ++o) { p[o] = expf(acc[o] - m); z += p[o]; }
#pragma unroll
    for (int o = 0; o < 4; ++o) p[o] /= z;
    if (logits) reinterpret_cast<float4*>(logits)[r] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (probs) reinterpret_cast<float4*>(probs)[r] = make_float4(p[0], p[1], p[2], p[3]);
    if (samples) {
        int pick = 0;
        if (temperature <= 0.f) {
            float best = p[0];
#pragma unroll
            for (int o = 1; o < 4; ++o)
                if (p[o] > best) { best = p[o]; pick = o; }           // first maximum, like argmax
        } else {
            float q[4], zq = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) { q[o] = expf((acc[o] - m) / temperature); zq += q[o]; }
            const float u = philox_uniform(seed, (uint32_t)r) * zq;
            float cdf = 0.f;
            pick = 3;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                cdf += q[o];
                if (u < cdf) { pick = o; break; }
            }
        }
        samples[r] = pick;
    }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                                       int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        const unsigned lo = __float_as_uint(v.x), lo2 = __float_as_uint(v.y);
        (void)lo; (void)lo2;
        unsigned short o[4];
        const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            unsigned u = __float_as_uint(f[q]);
            u += 0x7fffu + ((u >> 16) & 1u);                           // round to nearest even
            o[q] = (unsigned short)(u >> 16);
        }
        *reinterpret_cast<uint2*>(dst + i) = make_uint2((unsigned)o[0] | ((unsigned)o[1] << 16),
                                                        (unsigned)o[2] | ((unsigned)o[3] << 16));
    } else {
        for (int64_t j = i; j < n; ++j) {
            unsigned u = __float_as_uint(src[j]);
            u += 0x7fffu + ((u >> 16) & 1u);
            dst[j] = (unsigned short)(u >> 16);
        }
    }
}

}  // namespace rmp

using namespace rmp;

extern "C" int rmp_row_mlp(const rmp_mlp_desc* dp, void* stream) {
    RMP_REQUIRE(dp, RMP_E_ARG, "rmp_row_mlp: null descriptor");
    const rmp_mlp_desc& d = *dp;
    RMP_REQUIRE(d.num_segs >= 1 && d.num_segs <= RMP_MAX_SEGS, RMP_E_ARG, "rmp_row_mlp: num_segs=%d", d.num_segs);
    RMP_REQUIRE(d.num_stages >= 1 && d.num_stages <= RMP_MAX_STAGES, RMP_E_ARG, "rmp_row_mlp: num_stages=%d",
                d.num_stages);
    RMP_REQUIRE(d.out && d.rows_cap >= 0, RMP_E_ARG, "rmp_row_mlp: bad output / rows");
    if (d.rows_cap == 0) return RMP_OK;
    int k0 = 0;
    for (int s = 0; s < d.num_segs; ++s) {
        RMP_REQUIRE(d.seg_src[s] && d.seg_width[s] > 0, RMP_E_ARG, "rmp_row_mlp: segment %d", s);
        RMP_REQUIRE(!d.seg_bf16[s], RMP_E_ARG, "rmp_row_mlp: bf16 segments are a tensor-core arm feature");
        RMP_REQUIRE((d.seg_width[s] & 3) != 0 || aligned16(d.seg_src[s]), RMP_E_ALIGN,
                    "rmp_row_mlp: segment %d not 16-byte aligned", s);
        k0 += d.seg_width[s];
    }
    RMP_REQUIRE(k0 == d.k_in[0], RMP_E_SHAPE, "rmp_row_mlp: segments give %d inputs, stage 0 expects %d", k0,
                d.k_in[0]);
    int cap0 = round_up(k0, kKC), cap1 = 0, wmax = 0;
    for (int s = 0; s < d.num_stages; ++s) {
        RMP_REQUIRE(d.Wt[s] && aligned16(d.Wt[s]), RMP_E_ALIGN, "rmp_row_mlp: Wt[%d] null or unaligned", s);
        RMP_REQUIRE(d.n_out[s] > 0 && d.n_out[s] % kNB == 0, RMP_E_SHAPE, "rmp_row_mlp: n_out[%d]=%d not a multiple of %d",
                    s, d.n_out[s], kNB);
        RMP_REQUIRE(s == 0 || d.k_in[s] == d.n_out[s - 1], RMP_E_SHAPE, "rmp_row_mlp: stage %d input width mismatch", s);
        if (s & 1) cap0 = cap0 > d.n_out[s] ? cap0 : d.n_out[s];
        else cap1 = cap1 > d.n_out[s] ? cap1 : d.n_out[s];
        wmax = wmax > d.n_out[s] ? wmax : d.n_out[s];
    }
    const int n_last = d.n_out[d.num_stages - 1];
    if (d.epilogue != RMP_EPI_STORE && d.epilogue != RMP_EPI_RESID_ADD)
        RMP_REQUIRE(n_last == 128, RMP_E_SHAPE, "rmp_row_mlp: epilogue %d needs 128 output channels", d.epilogue);
    if (d.epilogue == RMP_EPI_NORMALIZE || d.epilogue == RMP_EPI_RESID_LN)
        RMP_REQUIRE(d.gain && d.beta, RMP_E_ARG, "rmp_row_mlp: gain/beta missing");
    if (d.epilogue == RMP_EPI_RESID_LN || d.epilogue == RMP_EPI_RESID_ADD) RMP_REQUIRE(d.resid, RMP_E_ARG, "rmp_row_mlp: resid missing");
    if (d.epilogue == RMP_EPI_SEGSUM) RMP_REQUIRE(d.seg_id, RMP_E_ARG, "rmp_row_mlp: seg_id missing");
    if (cap1 == 0) cap1 = kNB;
    cudaStream_t st = (cudaStream_t)stream;
    auto smem_for = [&](int rows) {
        return (size_t)(cap0 + cap1) * (rows + 4) * sizeof(float) + 2 * kKC * kNB * sizeof(float) + rows * sizeof(int);
    };
    const size_t kMaxSmem = 227 * 1024;
    const bool ext = d.row_scale != nullptr || d.epilogue == RMP_EPI_RESID_ADD;
    auto launch = [&](auto kern, int rows) {
        const size_t smem = smem_for(rows);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<ceil_div(d.rows_cap, rows), kThreads, smem, st>>>(d, cap0, cap1);
    };
    if (smem_for(64) <= kMaxSmem) {
        if (ext) launch(row_mlp_kernel<64, true>, 64);
        else launch(row_mlp_kernel<64, false>, 64);
    } else {
        RMP_REQUIRE(smem_for(32) <= kMaxSmem, RMP_E_SHAPE, "rmp_row_mlp: layer widths need %zu B of shared memory", smem_for(32));
        if (ext) launch(row_mlp_kernel<32, true>, 32);
        else launch(row_mlp_kernel<32, false>, 32);
    }
    return check_launch("rmp_row_mlp");
}

extern "C" int rmp_add_layernorm(const float* a, float* b, const float* gamma, const float* beta, float eps,
                                 const int32_t* rows_dev, int rows_slot, int rows_cap, int zero_b, int variant,
                                 float* out, void* stream) {
    RMP_REQUIRE(a && gamma && beta && out, RMP_E_ARG, "rmp_add_layernorm: null pointer");
    RMP_REQUIRE(aligned16(a) && aligned16(out) && aligned16(gamma) && aligned16(beta) && (!b || aligned16(b)),
                RMP_E_ALIGN, "rmp_add_layernorm: pointers must be 16-byte aligned");
    if (rows_cap <= 0) return RMP_OK;
    add_layernorm_kernel<<<ceil_div(rows_cap, 8), 256, 0, (cudaStream_t)stream>>>(a, b, gamma, beta, eps, rows_dev,
                                                                                  rows_slot, rows_cap, zero_b, variant, out);
    return check_launch("rmp_add_layernorm");
}

extern "C" int rmp_readout_head(const float* h, int k_in, const float* W, const float* b, float temperature,
                                uint64_t seed, const int32_t* rows_dev, int rows_slot, int rows_cap, float* logits,
                                float* probs, int64_t* samples, void* stream) {
    RMP_REQUIRE(h && W && b && k_in > 0, RMP_E_ARG, "rmp_readout_head: null pointer");
    RMP_REQUIRE((!logits || aligned16(logits)) && (!probs || aligned16(probs)), RMP_E_ALIGN,
                "rmp_readout_head: logits/probs must be 16-byte aligned");
    if (rows_cap <= 0) return RMP_OK;
    readout_head_kernel<<<ceil_div(rows_cap, 8), 256, 0, (cudaStream_t)stream>>>(
        h, k_in, W, b, temperature, seed, rows_dev, rows_slot, rows_cap, nullptr, logits, probs, samples);
    return check_launch("rmp_readout_head");
}

extern "C" int rmp_readout_head_masked(const float* h, int k_in, const float* W, const float* b, const float* row_mask,
                                       float temperature, uint64_t seed, int rows, float* logits, float* probs,
                                       int64_t* samples, void* stream) {
    RMP_REQUIRE(h && W && b && row_mask && k_in > 0, RMP_E_ARG, "rmp_readout_head_masked: null pointer");
    RMP_REQUIRE((!logits || aligned16(logits)) && (!probs || aligned16(probs)), RMP_E_ALIGN,
                "rmp_readout_head_masked: logits/probs must be 16-byte aligned");
    if (rows <= 0) return RMP_OK;
    readout_head_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        h, k_in, W, b, temperature, seed, nullptr, 0, rows, row_mask, logits, probs, samples);
    return check_launch("rmp_readout_head_masked");
}

extern "C" int rmp_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
    RMP_REQUIRE(src && dst && n >= 0, RMP_E_ARG, "rmp_cast_bf16: bad arguments");
    RMP_REQUIRE(aligned16(src) && aligned16(dst), RMP_E_ALIGN, "rmp_cast_bf16: pointers must be 16-byte aligned");
    if (n == 0) return RMP_OK;
    const int64_t quads = (n + 3) / 4;
    cast_bf16_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, reinterpret_cast<uint16_t*>(dst), n);
    return check_launch("rmp_cast_bf16");
}
