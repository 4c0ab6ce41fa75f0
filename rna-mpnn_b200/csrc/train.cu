// Training-path kernels of the rdesign message MLP (BASELINE configs[3]).
//
// The per-edge part of MPNNLayer's backward (rdesign/model/mpnn.py:15-19, :33-34: three Linear + GELU + Dropout on
// [h_E | h_V[ctr] | h_V[nbr]], scatter_sum / scale) is 96 % of the training step's flops and bytes.  It is
// decomposed into a few HBM-bound passes over bf16 [Ne,128] tensors and tcgen05 GEMMs:
//   forward   z_s = a_{s-1} W_s^T + b_s            rmp_row_mlp_tc (pre-activations kept as bf16 for the backward)
//             a_s = Dropout(GELU(z_s))             rmp_act_fwd_bf16        (Philox mask from (seed, stream, element))
//   backward  dz_s = da_s * mask/(1-p) * GELU'(z_s)  rmp_act_bwd_bf16      (mask regenerated, + column sums = db_s;
//                                                     the last stage gathers da from d(agg)[ctr] * alpha itself)
//             da_{s-1} = dz_s W_s                  rmp_row_mlp_tc with the transposed weight
//             dW_s += dz_s^T a_{s-1}               rmp_gemm_dw_bf16        (tcgen05, both operands MN-major from smem)
//             d(Pc) = segment sums of dz_0 by centre / neighbour: rmp_scatter_add_rows
// plus rmp_adam_step, a fused Adam update over the flat parameter / gradient buffers.
#include "tc_common.cuh"
#include "dropout.cuh"

namespace rmp {
namespace train {

// (the counter-based dropout mask generator lives in dropout.cuh)
// GELU (tanh form, MUFU.TANH) and its derivative; the same pair is used in both directions, so the backward is the
// exact derivative of what the forward computed (|tanh form - erf form| <= 5e-4, below the bf16 rounding of a)
__device__ __forceinline__ void gelu_tanh_both(float x, float& y, float& dy) {
    const float x2 = x * x;
    const float u = x * fmaf(x2, 0.0356774081f, 0.7978845608f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float hx = 0.5f * x;
    y = fmaf(hx, t, hx);
    // d/dx: 0.5 (1 + t) + 0.5 x (1 - t^2) (c0 + 3 c1 x^2)
    dy = fmaf(hx * (1.f - t * t), fmaf(x2, 0.1070322243f, 0.7978845608f), 0.5f + 0.5f * t);
}
__device__ __forceinline__ float gelu_tanh_y(float x) { float y, d; gelu_tanh_both(x, y, d); return y; }
__device__ __forceinline__ float gelu_tanh_dy(float x) { float y, d; gelu_tanh_both(x, y, d); return d; }
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// a = Dropout(GELU(z)); one thread per 8 consecutive elements (16 bytes)
__global__ void __launch_bounds__(256) act_fwd_kernel(const uint4* __restrict__ z, uint4* __restrict__ a, size_t n_groups,
                                                      uint32_t thresh16, float keep_scale, uint64_t seed, const uint64_t* __restrict__ seed_dev, uint32_t stream) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    if (seed_dev) seed += __ldg(seed_dev);
    const uint4 zi = z[g];
    const uint32_t zw[4] = {zi.x, zi.y, zi.z, zi.w};
    const uint32_t keep = keep_mask8(seed, g, stream, thresh16);
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float lo = (keep >> (2 * j)) & 1u ? gelu_tanh_y(bf_lo(zw[j])) * keep_scale : 0.f;
        const float hi = (keep >> (2 * j + 1)) & 1u ? gelu_tanh_y(bf_hi(zw[j])) * keep_scale : 0.f;
        o[j] = tc::pack_bf16(lo, hi);
    }
    a[g] = make_uint4(o[0], o[1], o[2], o[3]);
}

// dz = da * mask/(1-p) * GELU'(z), da either a bf16 tensor or gathered from dagg[ctr[row]] * alpha (fp32 rows);
// column sums of dz are accumulated into db.  Block = 256 threads = 16 rows x 16 column groups of 8 (width 128).
__global__ void __launch_bounds__(256) act_bwd_kernel(const uint4* __restrict__ z, const uint4* __restrict__ da,
                                                      const float* __restrict__ dagg, const int32_t* __restrict__ ctr,
                                                      float alpha, uint4* __restrict__ dz, float* __restrict__ db,
                                                      int rows, int rows_per_block, uint32_t thresh16, float keep_scale,
                                                      uint64_t seed, const uint64_t* __restrict__ seed_dev, uint32_t stream) {
    __shared__ float s_db[16][128];
    if (seed_dev) seed += __ldg(seed_dev);
    const int cg = threadIdx.x & 15, rl = threadIdx.x >> 4;          // column group (8 columns), row lane
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(rows, r0 + rows_per_block);
    for (int r = r0 + rl; r < r1; r += 16) {
        const size_t g = (size_t)r * 16 + cg;
        const uint4 zi = z[g];
        const uint32_t zw[4] = {zi.x, zi.y, zi.z, zi.w};
        float d[8];
        if (da) {
            const uint4 di = da[g];
            const uint32_t dw[4] = {di.x, di.y, di.z, di.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { d[2 * j] = bf_lo(dw[j]); d[2 * j + 1] = bf_hi(dw[j]); }
        } else {
            const float4* src = reinterpret_cast<const float4*>(dagg + (size_t)__ldg(ctr + r) * 128 + cg * 8);
            const float4 x = __ldg(src), y = __ldg(src + 1);
            d[0] = x.x * alpha; d[1] = x.y * alpha; d[2] = x.z * alpha; d[3] = x.w * alpha;
            d[4] = y.x * alpha; d[5] = y.y * alpha; d[6] = y.z * alpha; d[7] = y.w * alpha;
        }
        const uint32_t keep = keep_mask8(seed, g, stream, thresh16);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float lo = (keep >> (2 * j)) & 1u ? d[2 * j] * keep_scale * gelu_tanh_dy(bf_lo(zw[j])) : 0.f;
            const float hi = (keep >> (2 * j + 1)) & 1u ? d[2 * j + 1] * keep_scale * gelu_tanh_dy(bf_hi(zw[j])) : 0.f;
            o[j] = tc::pack_bf16(lo, hi);
            acc[2 * j] += bf_lo(o[j]); acc[2 * j + 1] += bf_hi(o[j]);      // the bias gradient sums what the GEMMs will see
        }
        dz[g] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (db) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s_db[rl][cg * 8 + j] = acc[j];
        __syncthreads();
        if (threadIdx.x < 128) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) t += s_db[k][threadIdx.x];
            atomicAdd(db + threadIdx.x, t);
        }
    }
}

// out[idx[r], :] += x[r, :]  (x bf16 [rows,128], out fp32): one thread per 8 columns, red.global.v4
__global__ void __launch_bounds__(256) scatter_add_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ idx,
                                                          float* __restrict__ out, int rows) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)rows * 16) return;
    const int r = (int)(g >> 4), cg = (int)(g & 15);
    const uint4 xi = x[g];
    float* dst = out + (size_t)__ldg(idx + r) * 128 + cg * 8;
    tc::red_add_v4(dst, bf_lo(xi.x), bf_hi(xi.x), bf_lo(xi.y), bf_hi(xi.y));
    tc::red_add_v4(dst + 4, bf_lo(xi.z), bf_hi(xi.z), bf_lo(xi.w), bf_hi(xi.w));
}

// nn.LayerNorm over 128 channels with the residual add in front of it (rdesign/model/mpnn.py:35,37), training forward:
//   x = a + b (b fp32 or bf16, optional) -> x_out (kept for the backward), y = gamma * (x - mean) * rstd + beta, mean, rstd.
// One warp per row, one float4 per lane.
template <bool B_BF16>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const float4* __restrict__ a, const void* __restrict__ b_,
                                                         const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                                         float eps, int rows, float4* __restrict__ x_out, float4* __restrict__ y,
                                                         float* __restrict__ mean, float* __restrict__ rstd) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    float4 v = a[(size_t)r * 32 + lane];
    if (b_) {
        if constexpr (B_BF16) {
            const uint2 w = reinterpret_cast<const uint2*>(b_)[(size_t)r * 32 + lane];
            v.x += bf_lo(w.x); v.y += bf_hi(w.x); v.z += bf_lo(w.y); v.w += bf_hi(w.y);
        } else {
            const float4 w = reinterpret_cast<const float4*>(b_)[(size_t)r * 32 + lane];
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        x_out[(size_t)r * 32 + lane] = v;
    }
    const float mu = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
    const float c0 = v.x - mu, c1 = v.y - mu, c2 = v.z - mu, c3 = v.w - mu;
    const float var = warp_sum((c0 * c0 + c1 * c1) + (c2 * c2 + c3 * c3)) * (1.f / 128.f);
    const float rs = 1.f / sqrtf(var + eps);
    const float4 g = __ldg(gamma + lane), be = __ldg(beta + lane);
    y[(size_t)r * 32 + lane] = make_float4(fmaf(g.x, c0 * rs, be.x), fmaf(g.y, c1 * rs, be.y), fmaf(g.z, c2 * rs, be.z),
                                           fmaf(g.w, c3 * rs, be.w));
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
}

// LayerNorm backward in one pass: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) with g = dy * gamma, plus the parameter
// gradients dgamma += sum dy * xhat, dbeta += sum dy (per-warp registers -> shared -> one red per column and block); dx16 is
// an optional bf16 copy of dx (the gradient of a bf16 residual branch).
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float4* __restrict__ gamma, int rows, int rows_per_block,
                                                     float4* __restrict__ dx, uint2* __restrict__ dx16,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float s_acc[2][8][128];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 gm = __ldg(gamma + lane);
    float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
    const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int r = r0 + w; r < r1; r += 8) {
        const float4 d = dy[(size_t)r * 32 + lane], v = x[(size_t)r * 32 + lane];
        const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
        const float xh[4] = {(v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs};
        const float dd[4] = {d.x, d.y, d.z, d.w}, gg[4] = {gm.x, gm.y, gm.z, gm.w};
        float g[4], t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            g[j] = dd[j] * gg[j]; t1 += g[j]; t2 += g[j] * xh[j];
            ag[j] += dd[j] * xh[j]; ab[j] += dd[j];
        }
        t1 = warp_sum(t1) * (1.f / 128.f); t2 = warp_sum(t2) * (1.f / 128.f);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = rs * (g[j] - t1 - xh[j] * t2);
        dx[(size_t)r * 32 + lane] = make_float4(o[0], o[1], o[2], o[3]);
        if (dx16) dx16[(size_t)r * 32 + lane] = make_uint2(tc::pack_bf16(o[0], o[1]), tc::pack_bf16(o[2], o[3]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_acc[0][w][lane * 4 + j] = ag[j]; s_acc[1][w][lane * 4 + j] = ab[j]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        float t = 0.f, u = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t += s_acc[0][k][threadIdx.x]; u += s_acc[1][k][threadIdx.x]; }
        atomicAdd(dgamma + threadIdx.x, t);
        atomicAdd(dbeta + threadIdx.x, u);
    }
}

// Column sums over the rows of a [rows, width] matrix (bias gradients of the per-node Linear / LayerNorm modules):
//   s0[c] += sum_r a[r][c]                                     a fp32 or bf16
//   s1[c] += sum_r a[r][c] * (x[r][c] - mean[r]) * rstd[r]     (x fp32, optional: the LayerNorm weight gradient)
// Thread = 4 consecutive columns x a row lane; per-block partial sums through shared memory, one red per column and block.
template <bool A_BF16>
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ a_, const float* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd, int rows,
                                                     int width, int rows_per_block, float* __restrict__ s0,
                                                     float* __restrict__ s1) {
    __shared__ float4 sh[2][256];
    const int ncg = width >> 2;                                  // column groups of 4: 32 (width 128) .. 64 (width 256)
    const int cg = threadIdx.x % ncg, rl = threadIdx.x / ncg, nrl = 256 / ncg;
    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int r = r0 + rl; r < r1; r += nrl) {
        float4 v;
        if constexpr (A_BF16) {
            const uint2 w = reinterpret_cast<const uint2*>(a_)[(size_t)r * ncg + cg];
            v = make_float4(bf_lo(w.x), bf_hi(w.x), bf_lo(w.y), bf_hi(w.y));
        } else {
            v = reinterpret_cast<const float4*>(a_)[(size_t)r * ncg + cg];
        }
        t0.x += v.x; t0.y += v.y; t0.z += v.z; t0.w += v.w;
        if (x) {
            const float4 xv = reinterpret_cast<const float4*>(x)[(size_t)r * ncg + cg];
            const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
            t1.x += v.x * ((xv.x - mu) * rs); t1.y += v.y * ((xv.y - mu) * rs);
            t1.z += v.z * ((xv.z - mu) * rs); t1.w += v.w * ((xv.w - mu) * rs);
        }
    }
    sh[0][threadIdx.x] = t0; sh[1][threadIdx.x] = t1;
    __syncthreads();
    if (threadIdx.x < ncg) {
        float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
        for (int k = 0; k < nrl; ++k) {
            const float4 p0 = sh[0][k * ncg + threadIdx.x], p1 = sh[1][k * ncg + threadIdx.x];
            u0.x += p0.x; u0.y += p0.y; u0.z += p0.z; u0.w += p0.w;
            u1.x += p1.x; u1.y += p1.y; u1.z += p1.z; u1.w += p1.w;
        }
        tc::red_add_v4(s0 + 4 * threadIdx.x, u0.x, u0.y, u0.z, u0.w);
        if (x) tc::red_add_v4(s1 + 4 * threadIdx.x, u1.x, u1.y, u1.z, u1.w);
    }
}

// out[v, :] = sum of the rows row_ptr[v] .. row_ptr[v+1]-1 of x (bf16 [rows,128]; the edges of a node are contiguous,
// SURVEY section 8a a-2): the scatter_sum by CENTRE node without atomics.  Half a warp per node: lane = 8 columns (one
// 16-byte load per row), rows summed in order in fp32 -> deterministic; one pass over x at HBM rate where the red.global
// version moved every row through the L2 atomic units (70 -> ~25 us on [409,600 x 128]).
// With `perm` the segment's rows are x[perm[r]]: the scatter by NEIGHBOUR node as a gather through the edges sorted by
// neighbour (the transposed CSR of the graph, built once per step), deterministic and without atomics as well.
template <bool PERM>
__global__ void __launch_bounds__(256) segment_sum_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ perm,
                                                          const int32_t* __restrict__ row_ptr, float* __restrict__ out, int n_seg) {
    const int v = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    const int cg = threadIdx.x & 15;
    if (v >= n_seg) return;
    const int r0 = __ldg(row_ptr + v), r1 = __ldg(row_ptr + v + 1);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int r = r0;
    for (; r + 4 <= r1; r += 4) {                    // four independent 16-byte loads in flight
        uint4 xi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) xi[q] = x[(size_t)(PERM ? __ldg(perm + r + q) : r + q) * 16 + cg];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            acc[0] += bf_lo(xi[q].x); acc[1] += bf_hi(xi[q].x); acc[2] += bf_lo(xi[q].y); acc[3] += bf_hi(xi[q].y);
            acc[4] += bf_lo(xi[q].z); acc[5] += bf_hi(xi[q].z); acc[6] += bf_lo(xi[q].w); acc[7] += bf_hi(xi[q].w);
        }
    }
    for (; r < r1; ++r) {
        const uint4 xi = x[(size_t)(PERM ? __ldg(perm + r) : r) * 16 + cg];
        acc[0] += bf_lo(xi.x); acc[1] += bf_hi(xi.x); acc[2] += bf_lo(xi.y); acc[3] += bf_hi(xi.y);
        acc[4] += bf_lo(xi.z); acc[5] += bf_hi(xi.z); acc[6] += bf_lo(xi.w); acc[7] += bf_hi(xi.w);
    }
    float4* dst = reinterpret_cast<float4*>(out + (size_t)v * 128 + cg * 8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// Adam (torch.optim.Adam semantics, no weight decay / amsgrad): p -= lr * m_hat / (sqrt(v_hat) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                   float4* __restrict__ v, size_t n4, float grad_scale, float lr, float b1,
                                                   float b2, float eps, float bc1, float bc2_sqrt,
                                                   const int* __restrict__ step_dev) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    if (step_dev) {                 // graph replay: the step number lives on the device
        const float t = (float)__ldg(step_dev);
        bc1 = 1.0f - powf(b1, t);
        bc2_sqrt = sqrtf(1.0f - powf(b2, t));
    }
    float4 pi = p[i], mi = m[i], vi = v[i];
    const float4 gi = g[i];
    const float step = lr / bc1;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= grad_scale;
        mm = b1 * mm + (1.f - b1) * gg;
        vv = b2 * vv + (1.f - b2) * gg * gg;
        pp -= step * mm / (sqrtf(vv) / bc2_sqrt + eps);
    };
    upd(pi.x, gi.x, mi.x, vi.x); upd(pi.y, gi.y, mi.y, vi.y); upd(pi.z, gi.z, mi.z, vi.z); upd(pi.w, gi.w, mi.w, vi.w);
    p[i] = pi; m[i] = mi; v[i] = vi;
}

// Reference Normalize (rdesign/model/functional.py:79-96) over 128 channels, forward and backward, one warp per row:
//   y = gain * (x - mu) / (sqrt(var_unbiased + eps) + eps) + bias
// forward : x fp32 -> y bf16 (the layers' h_E operand)
// backward: dy fp32 (d h_E accumulated over the layers) -> dx bf16 (operand of the weight-gradient GEMM), and the
//           column sums d gain = sum dy * xhat, d bias = sum dy, d b = sum dx (bias gradient of the Linear in front)
__global__ void __launch_bounds__(256) normalize_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gain,
                                                            const float4* __restrict__ bias, uint2* __restrict__ y16, int rows,
                                                            float eps) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float4 v = x[(size_t)r * 32 + lane];
    const float mu = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
    const float a = v.x - mu, b = v.y - mu, c = v.z - mu, d = v.w - mu;
    const float var = warp_sum((a * a + b * b) + (c * c + d * d)) * (1.f / 127.f);
    const float inv = 1.f / (sqrtf(var + eps) + eps);
    const float4 g = __ldg(gain + lane), be = __ldg(bias + lane);
    y16[(size_t)r * 32 + lane] = make_uint2(tc::pack_bf16(fmaf(g.x, a * inv, be.x), fmaf(g.y, b * inv, be.y)),
                                            tc::pack_bf16(fmaf(g.z, c * inv, be.z), fmaf(g.w, d * inv, be.w)));
}

__global__ void __launch_bounds__(256) normalize_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ dy,
                                                            const float4* __restrict__ gain, uint2* __restrict__ dx16,
                                                            float* __restrict__ dgain, float* __restrict__ dbias,
                                                            float* __restrict__ dbsum, int rows, int rows_per_block, float eps) {
    __shared__ float s_acc[3][8][128];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 g = __ldg(gain + lane);
    float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f}, ax[4] = {0.f, 0.f, 0.f, 0.f};
    const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int r = r0 + w; r < r1; r += 8) {
        const float4 v = x[(size_t)r * 32 + lane], d = dy[(size_t)r * 32 + lane];
        const float mu = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
        const float xc[4] = {v.x - mu, v.y - mu, v.z - mu, v.w - mu};
        const float var = warp_sum((xc[0] * xc[0] + xc[1] * xc[1]) + (xc[2] * xc[2] + xc[3] * xc[3])) * (1.f / 127.f);
        const float sigma = sqrtf(var + eps), s = sigma + eps, inv = 1.f / s;
        const float dd[4] = {d.x, d.y, d.z, d.w}, gg[4] = {g.x, g.y, g.z, g.w};
        float dxh[4], t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { dxh[j] = dd[j] * gg[j]; t1 += dxh[j] * xc[j]; t2 += dxh[j]; }
        t1 = warp_sum(t1); t2 = warp_sum(t2);
        // d s = -sum(dxh * xc) / s^2;  d xc_i = dxh_i / s + d s * xc_i / (sigma (n - 1));  d x_i = d xc_i - mean(d xc)
        const float cvar = -t1 * inv * inv / (sigma * 127.f), mean_dxc = t2 * inv * (1.f / 128.f);
        float dx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dx[j] = dxh[j] * inv + cvar * xc[j] - mean_dxc;
            ag[j] += dd[j] * xc[j] * inv; ab[j] += dd[j];
        }
        const uint2 o = make_uint2(tc::pack_bf16(dx[0], dx[1]), tc::pack_bf16(dx[2], dx[3]));
        dx16[(size_t)r * 32 + lane] = o;
        ax[0] += bf_lo(o.x); ax[1] += bf_hi(o.x); ax[2] += bf_lo(o.y); ax[3] += bf_hi(o.y);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_acc[0][w][4 * lane + j] = ag[j]; s_acc[1][w][4 * lane + j] = ab[j]; s_acc[2][w][4 * lane + j] = ax[j]; }
    __syncthreads();
    if (threadIdx.x < 128) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t0 += s_acc[0][k][threadIdx.x]; t1 += s_acc[1][k][threadIdx.x]; t2 += s_acc[2][k][threadIdx.x]; }
        atomicAdd(dgain + threadIdx.x, t0); atomicAdd(dbias + threadIdx.x, t1); atomicAdd(dbsum + threadIdx.x, t2);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// dW[n][k] += sum_r dZ[r][n] X[r][k]   (dZ, X bf16 [rows,128]; dW fp32 [128,128])
// The contraction runs over the rows, so both operand tiles are read MN-major from their natural row-major
// [128 rows][128 cols] layout (two 128-byte-swizzled [128][64] blocks per tile, exactly what TMA delivers):
// A = dZ^T (M = n), B = X^T (N = k), K = 16 rows per tcgen05.mma.  One persistent CTA per SM accumulates its row
// tiles in TMEM and adds its 128x128 partial to dW with red.global at the end.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kDwStages = 3;
constexpr int kDwThreads = 192;                   // warp 0 TMA, warp 1 MMA, warps 2..5 drain (one TMEM lane quadrant each)
struct alignas(8) DwBarriers { uint64_t full[kDwStages], empty[kDwStages], done; };

__global__ void __launch_bounds__(kDwThreads, 1)
gemm_dw_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_x, float* __restrict__ dW,
               int rows) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                                    // [stages][2][128][64] dZ tiles
    uint8_t* sB = sA + kDwStages * kTileBytes;             // [stages][2][128][64] X tiles
    DwBarriers* bars = reinterpret_cast<DwBarriers*>(sB + kDwStages * kTileBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (rows + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDwStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        mbar_init(&bars->done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    if (warp == 0) {
        if (elect_one()) {
            tma_prefetch_desc(&tm_dz); tma_prefetch_desc(&tm_x);
            for (int i = 0; i < my_tiles; ++i) {
                const int st = i % kDwStages;
                const int tile = blockIdx.x + i * gridDim.x;
                mbar_wait(&bars->empty[st], ((i / kDwStages) & 1) ^ 1);
                mbar_expect_tx(&bars->full[st], 2 * kTileBytes);      // rows past the end are zero-filled by TMA
                for (int kb = 0; kb < 2; ++kb) {
                    tma_load_2d(sA + (size_t)st * kTileBytes + kb * kKBlockBytes, &tm_dz, &bars->full[st], kb * 64, tile * kTileM);
                    tma_load_2d(sB + (size_t)st * kTileBytes + kb * kKBlockBytes, &tm_x, &bars->full[st], kb * 64, tile * kTileM);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = instr_desc_bf16(128, 128, 1) | (1u << 15);     // A and B MN-major
            for (int i = 0; i < my_tiles; ++i) {
                const int st = i % kDwStages;
                mbar_wait(&bars->full[st], (i / kDwStages) & 1);
                tc_fence_after();
                const uint64_t d_a = smem_desc_sw128(smem_u32(sA + (size_t)st * kTileBytes), kKBlockBytes, 1024);
                const uint64_t d_b = smem_desc_sw128(smem_u32(sB + (size_t)st * kTileBytes), kKBlockBytes, 1024);
#pragma unroll
                for (int k = 0; k < 8; ++k) {                      // 128 rows = 8 x UMMA_K(16); 16 rows = 2048 B
                    const uint64_t off = (uint64_t)((k * 2048) >> 4);
                    umma_bf16(tmem_base, d_a + off, d_b + off, idesc, (uint32_t)((i | k) > 0));
                }
                umma_commit(&bars->empty[st]);
            }
            umma_commit(&bars->done);
        }
        __syncwarp();
    } else if (my_tiles > 0) {
        // drain: thread = output row n (TMEM lane), 128 columns k in four chunks
        const int lane_grp = warp & 3;
        const int n = lane_grp * 32 + lane;
        mbar_wait(&bars->done, 0);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v[32];
            tmem_ld32(t_row + (uint32_t)(c * 32), v);
            float* dst = dW + (size_t)n * 128 + c * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 128);
}


// ------------------------------------------------------------------------------------------------------------------
// Z[r][n] = sum_k A[r][k] W[n][k] (+ bias[n]) (+ Pc[ctr[r]][n] + Pn[nbr[r]][n])      A bf16 [rows,128], W bf16 [128,128]
// One Linear stage of the training path as a lean, HBM-bound GEMM: A tiles by TMA through a 4-slot ring, W resident,
// one tcgen05 128x128x128 GEMM per tile into one of two TMEM accumulators, 8 epilogue warps (thread = row x 64
// columns) that overlap the next tile's GEMM.  Epilogue outputs (any subset):
//   z16   bf16 [rows,128]   the pre-activation (kept for the backward)
//   a16   bf16 [rows,128]   Dropout(GELU(z16)): the activation pass fused in (forward)
//   acc32 fp32 [rows,128]   += Z: gradient accumulation over the layers (d h_E)
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGrStages = 3;
constexpr int kGrEpiWarps = 16;
constexpr int kGrThreads = 128 + 32 * kGrEpiWarps;         // warp 0 TMA, warp 1 MMA, warps 2..3 idle, warps 4..19 epilogue
struct alignas(8) GrBarriers { uint64_t w_full, full[kGrStages], empty[kGrStages], acc_full[2], acc_empty[2]; };
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(tc::smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void gr_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(512) : "memory"); }

struct GrParams {
    const float* bias;
    const uint16_t* Pc; const uint16_t* Pn; const int32_t* ctr; const int32_t* nbr;      // bf16 [Nv,128] per-node projections
    uint16_t* z16; uint16_t* a16; float* acc32;
    int rows;
    uint32_t thresh16; float keep_scale; uint64_t seed; const uint64_t* seed_dev; uint32_t stream;
    int w_mn_major;                 // 1: W is used transposed (Z = A W): read MN-major, no transposed copy needed
    // K-concatenated mode (k_blocks > 1): Z = sum_l A_l W_l with A = [k_blocks][a_block_rows][128] and W = [k_blocks][128][128];
    // W_l travels through the ring with A_l (the slots are the resident-W buffer and the two staging buffers, which the
    // mode does not use: its only output is acc32)
    int k_blocks, a_block_rows;
    int acc32_store;                // 1: acc32 = Z (no read-modify-write)
};

__global__ void __launch_bounds__(kGrThreads, 1)
gemm_rows_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_act, const GrParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem;                                    // [2][128][64] bf16
    uint8_t* sA = sW + kTileBytes;                         // [stages][2][128][64]
    uint8_t* sZ = sA + kGrStages * kTileBytes;             // [2][128][64] bf16 staging of the z16 tile (TMA store source)
    uint8_t* sAct = sZ + kTileBytes;                       // the same for the a16 tile
    GrBarriers* bars = reinterpret_cast<GrBarriers*>(sAct + kTileBytes);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (p.rows + kTileM - 1) / kTileM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (threadIdx.x == 0) {
        mbar_init(&bars->w_full, 1);
        for (int i = 0; i < kGrStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 32 * kGrEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    if (warp == 0) {
        if (elect_one()) {
            bool my_tiles_done = false;
            tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_w);
            if (p.k_blocks > 1) {
                uint32_t n = 0;
                for (int i = 0; i < my_tiles; ++i) {
                    const int tile = blockIdx.x + i * gridDim.x;
                    for (int l = 0; l < p.k_blocks; ++l, ++n) {
                        const int st = (int)(n % kGrStages);
                        uint8_t* w_slot = st == 0 ? sW : (st == 1 ? sZ : sAct);
                        mbar_wait(&bars->empty[st], ((n / kGrStages) & 1) ^ 1);
                        mbar_expect_tx(&bars->full[st], 2 * kTileBytes);
                        for (int kb = 0; kb < 2; ++kb) {
                            tma_load_2d(sA + (size_t)st * kTileBytes + kb * kKBlockBytes, &tm_a, &bars->full[st], kb * 64,
                                        l * p.a_block_rows + tile * kTileM);
                            tma_load_2d(w_slot + kb * kKBlockBytes, &tm_w, &bars->full[st], kb * 64, l * 128);
                        }
                    }
                }
                my_tiles_done = true;
            }
            if (!my_tiles_done) {
            mbar_expect_tx(&bars->w_full, kTileBytes);
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(sW + kb * kKBlockBytes, &tm_w, &bars->w_full, kb * 64, 0);
            for (int i = 0; i < my_tiles; ++i) {
                const int st = i % kGrStages;
                const int tile = blockIdx.x + i * gridDim.x;
                mbar_wait(&bars->empty[st], ((i / kGrStages) & 1) ^ 1);
                mbar_expect_tx(&bars->full[st], kTileBytes);
                for (int kb = 0; kb < 2; ++kb)
                    tma_load_2d(sA + (size_t)st * kTileBytes + kb * kKBlockBytes, &tm_a, &bars->full[st], kb * 64, tile * kTileM);
            }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = instr_desc_bf16(128, 128, p.w_mn_major);
            if (p.k_blocks > 1) {
                uint32_t n = 0;
                for (int i = 0; i < my_tiles; ++i) {
                    const int ab = i & 1;
                    mbar_wait(&bars->acc_empty[ab], ((i >> 1) & 1) ^ 1);
                    for (int l = 0; l < p.k_blocks; ++l, ++n) {
                        const int st = (int)(n % kGrStages);
                        mbar_wait(&bars->full[st], (n / kGrStages) & 1);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(sA + (size_t)st * kTileBytes);
                        const uint32_t w_addr = smem_u32(st == 0 ? sW : (st == 1 ? sZ : sAct));
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t off = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                            const uint64_t d_w = p.w_mn_major ? smem_desc_sw128(w_addr + (uint32_t)(k * 2048), kKBlockBytes, 1024)
                                                              : smem_desc_sw128(w_addr + off, 16, 1024);
                            umma_bf16(tmem_base + (uint32_t)(ab * 128), smem_desc_sw128(a_addr + off, 16, 1024), d_w, idesc,
                                      (uint32_t)((l | k) > 0));
                        }
                        umma_commit(&bars->empty[st]);
                    }
                    umma_commit(&bars->acc_full[ab]);
                }
            } else {
            mbar_wait(&bars->w_full, 0);
            for (int i = 0; i < my_tiles; ++i) {
                const int st = i % kGrStages, ab = i & 1;
                mbar_wait(&bars->acc_empty[ab], ((i >> 1) & 1) ^ 1);
                mbar_wait(&bars->full[st], (i / kGrStages) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + (size_t)st * kTileBytes), w_addr = smem_u32(sW);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t off = (uint32_t)((k >> 2) * kKBlockBytes + (k & 3) * 32);
                    // W as stored is [n][k] (Z = A W^T, K-major B); for Z = A W the same bytes are B = W read MN-major:
                    // rows of the tile are the contraction index, 16 of them (2048 B) per MMA
                    const uint64_t d_w = p.w_mn_major ? smem_desc_sw128(w_addr + (uint32_t)(k * 2048), kKBlockBytes, 1024)
                                                      : smem_desc_sw128(w_addr + off, 16, 1024);
                    umma_bf16(tmem_base + (uint32_t)(ab * 128), smem_desc_sw128(a_addr + off, 16, 1024), d_w, idesc, (uint32_t)(k > 0));
                }
                umma_commit(&bars->empty[st]);
                umma_commit(&bars->acc_full[ab]);
            }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int lane_grp = warp & 3;
        const int row = lane_grp * 32 + lane;
        const int col0 = (ew >> 2) * 32;                   // this thread's 32 columns
        const uint64_t seed = p.seed + (p.seed_dev ? __ldg(p.seed_dev) : 0ull);
        const uint32_t t_row = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        float bias[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) bias[q] = p.bias ? __ldg(p.bias + col0 + q) : 0.f;
        for (int i = 0; i < my_tiles; ++i) {
            const int ab = i & 1;
            const int tile = blockIdx.x + i * gridDim.x;
            const int r = tile * kTileM + row;
            const bool valid = r < p.rows;
            // gathered per-node projections (the h_V[E_idx] part of W1): requested before the accumulator wait
            float add[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) add[q] = bias[q];
            if (p.Pc && valid) {
                uint32_t pc[16], pn[16];
                const uint16_t* pcr = p.Pc + (size_t)__ldg(p.ctr + r) * 128 + col0;
                const uint16_t* pnr = p.Pn + (size_t)__ldg(p.nbr + r) * 128 + col0;
                ldg256(pcr, reinterpret_cast<uint32_t(&)[8]>(pc[0])); ldg256(pcr + 16, reinterpret_cast<uint32_t(&)[8]>(pc[8]));
                ldg256(pnr, reinterpret_cast<uint32_t(&)[8]>(pn[0])); ldg256(pnr + 16, reinterpret_cast<uint32_t(&)[8]>(pn[8]));
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    add[2 * q] += bf_lo(pc[q]) + bf_lo(pn[q]);
                    add[2 * q + 1] += bf_hi(pc[q]) + bf_hi(pn[q]);
                }
            }
            mbar_wait(&bars->acc_full[ab], (i >> 1) & 1);
            tc_fence_after();
            float v[32];
            tmem_ld32(t_row + (uint32_t)(ab * 128 + col0), v);
            tc_fence_before();
            mbar_arrive(&bars->acc_empty[ab]);             // the accumulator is in registers: free for the next GEMM
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += add[q];
            if (p.acc32 && valid) {
                float* dst = p.acc32 + (size_t)r * 128 + col0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float t[8];
                    if (p.acc32_store) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) t[k] = v[8 * q + k];
                    } else {
                        ldg256f(dst + 8 * q, t);
#pragma unroll
                        for (int k = 0; k < 8; ++k) t[k] += v[8 * q + k];
                    }
                    stg256f(dst + 8 * q, t);
                }
            }
            // bf16 outputs: staged in shared memory in the 128-byte-swizzled tile layout and written with TMA stores
            // (a thread-per-row 32-byte store pattern touches 32 lines per instruction and was LSU-bound)
            uint32_t zp[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) zp[q] = pack_bf16(v[2 * q], v[2 * q + 1]);
            if (i > 0) {                                   // the previous tile's TMA stores have read the staging buffers
                if (threadIdx.x == 128) tma_store_wait_read();
                gr_epi_bar();
            }
            const int kb = col0 >> 6, j0 = (col0 & 63) >> 3;
            if (p.z16) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(sZ + kb * kKBlockBytes + sw128_chunk(row, j0 + j)) =
                        make_uint4(zp[4 * j], zp[4 * j + 1], zp[4 * j + 2], zp[4 * j + 3]);
            }
            if (p.a16) {                                   // a = Dropout(GELU(bf16(z))): what rmp_act_fwd_bf16 would write
                uint32_t ap[16];
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    const uint32_t keep = keep_mask8(seed, (uint64_t)r * 16 + (uint64_t)(col0 / 8 + g8), p.stream, p.thresh16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t w = zp[4 * g8 + j];
                        const float lo = (keep >> (2 * j)) & 1u ? gelu_tanh_y(bf_lo(w)) * p.keep_scale : 0.f;
                        const float hi = (keep >> (2 * j + 1)) & 1u ? gelu_tanh_y(bf_hi(w)) * p.keep_scale : 0.f;
                        ap[4 * g8 + j] = pack_bf16(lo, hi);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(sAct + kb * kKBlockBytes + sw128_chunk(row, j0 + j)) =
                        make_uint4(ap[4 * j], ap[4 * j + 1], ap[4 * j + 2], ap[4 * j + 3]);
            }
            if (p.z16 || p.a16) {
                fence_proxy_async();                       // staging bytes -> visible to the TMA engine
                gr_epi_bar();
                if (threadIdx.x == 128) {                  // rows past the end of the tensor are clipped by the TMA unit
                    for (int b2 = 0; b2 < 2; ++b2) {
                        if (p.z16) tma_store_2d(&tm_z, sZ + b2 * kKBlockBytes, b2 * 64, tile * kTileM);
                        if (p.a16) tma_store_2d(&tm_act, sAct + b2 * kKBlockBytes, b2 * 64, tile * kTileM);
                    }
                    tma_store_commit();
                }
            }
        }
        if (threadIdx.x == 128 && my_tiles > 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace train
}  // namespace rmp

using namespace rmp;

static uint32_t drop_threshold(float p) { return p <= 0.f ? 0u : (uint32_t)(p * 256.0f + 0.5f); }
static float keep_scale_of(float p) { return 256.0f / (256.0f - (float)drop_threshold(p)); }

extern "C" int rmp_act_fwd_bf16(const void* z_bf16, void* a_bf16, int64_t rows, int width, float p_drop, uint64_t seed,
                                const uint64_t* seed_dev, uint32_t stream_id, void* stream) {
    RMP_REQUIRE(z_bf16 && a_bf16 && rows >= 0 && width > 0 && width % 8 == 0, RMP_E_ARG, "rmp_act_fwd_bf16: bad arguments");
    RMP_REQUIRE(aligned16(z_bf16) && aligned16(a_bf16), RMP_E_ALIGN, "rmp_act_fwd_bf16: 16-byte alignment");
    RMP_REQUIRE(p_drop >= 0.f && p_drop < 1.f, RMP_E_ARG, "rmp_act_fwd_bf16: dropout probability in [0,1)");
    const size_t groups = (size_t)rows * width / 8;
    if (groups == 0) return RMP_OK;
    train::act_fwd_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(z_bf16), reinterpret_cast<uint4*>(a_bf16), groups, drop_threshold(p_drop),
        keep_scale_of(p_drop), seed, seed_dev, stream_id);
    return check_launch("rmp_act_fwd_bf16");
}

extern "C" int rmp_act_bwd_bf16(const void* z_bf16, const void* da_bf16, const float* dagg, const int32_t* ctr, float alpha,
                                void* dz_bf16, float* db, int rows, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                                uint32_t stream_id, void* stream) {
    RMP_REQUIRE(z_bf16 && dz_bf16 && rows >= 0, RMP_E_ARG, "rmp_act_bwd_bf16: null pointer");
    RMP_REQUIRE((da_bf16 != nullptr) != (dagg != nullptr && ctr != nullptr), RMP_E_ARG,
                "rmp_act_bwd_bf16: give either da_bf16 or (dagg, ctr)");
    RMP_REQUIRE(aligned16(z_bf16) && aligned16(dz_bf16) && (!da_bf16 || aligned16(da_bf16)) && (!dagg || aligned16(dagg)),
                RMP_E_ALIGN, "rmp_act_bwd_bf16: 16-byte alignment");
    if (rows == 0) return RMP_OK;
    const int rows_per_block = 128;
    train::act_bwd_kernel<<<ceil_div(rows, rows_per_block), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(z_bf16), reinterpret_cast<const uint4*>(da_bf16), dagg, ctr, alpha,
        reinterpret_cast<uint4*>(dz_bf16), db, rows, rows_per_block, drop_threshold(p_drop), keep_scale_of(p_drop), seed,
        seed_dev, stream_id);
    return check_launch("rmp_act_bwd_bf16");
}

extern "C" int rmp_scatter_add_rows(const void* x_bf16, const int32_t* idx, float* out, int rows, void* stream) {
    RMP_REQUIRE(x_bf16 && idx && out && rows >= 0, RMP_E_ARG, "rmp_scatter_add_rows: null pointer");
    RMP_REQUIRE(aligned16(x_bf16) && aligned16(out), RMP_E_ALIGN, "rmp_scatter_add_rows: 16-byte alignment");
    if (rows == 0) return RMP_OK;
    const size_t groups = (size_t)rows * 16;
    train::scatter_add_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(x_bf16), idx, out, rows);
    return check_launch("rmp_scatter_add_rows");
}

extern "C" int rmp_segment_sum_rows(const void* x_bf16, const int32_t* row_ptr, float* out, int n_seg, void* stream) {
    RMP_REQUIRE(x_bf16 && row_ptr && out && n_seg >= 0, RMP_E_ARG, "rmp_segment_sum_rows: null pointer");
    RMP_REQUIRE(aligned16(x_bf16) && aligned16(out), RMP_E_ALIGN, "rmp_segment_sum_rows: 16-byte alignment");
    if (n_seg == 0) return RMP_OK;
    const size_t threads = (size_t)n_seg * 16;
    train::segment_sum_kernel<false><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(x_bf16), nullptr, row_ptr, out, n_seg);
    return check_launch("rmp_segment_sum_rows");
}

extern "C" int rmp_segment_sum_rows_perm(const void* x_bf16, const int32_t* perm, const int32_t* row_ptr, float* out, int n_seg,
                                         void* stream) {
    RMP_REQUIRE(x_bf16 && perm && row_ptr && out && n_seg >= 0, RMP_E_ARG, "rmp_segment_sum_rows_perm: null pointer");
    RMP_REQUIRE(aligned16(x_bf16) && aligned16(out), RMP_E_ALIGN, "rmp_segment_sum_rows_perm: 16-byte alignment");
    if (n_seg == 0) return RMP_OK;
    const size_t threads = (size_t)n_seg * 16;
    train::segment_sum_kernel<true><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(x_bf16), perm, row_ptr, out, n_seg);
    return check_launch("rmp_segment_sum_rows_perm");
}

extern "C" int rmp_add_layernorm_fwd(const float* a, const void* b, int b_is_bf16, const float* gamma, const float* beta, float eps,
                                     int rows, float* x_out, float* y, float* mean, float* rstd, void* stream) {
    RMP_REQUIRE(a && gamma && beta && y && mean && rstd && rows >= 0 && (!b || x_out), RMP_E_ARG, "rmp_add_layernorm_fwd: null pointer");
    RMP_REQUIRE(aligned16(a) && aligned16(gamma) && aligned16(beta) && aligned16(y) && (!b || (aligned16(b) && aligned16(x_out))),
                RMP_E_ALIGN, "rmp_add_layernorm_fwd: 16-byte alignment");
    if (rows == 0) return RMP_OK;
    auto launch = [&](auto kernel) {
        kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4*>(a), b, reinterpret_cast<const float4*>(gamma), reinterpret_cast<const float4*>(beta), eps,
            rows, reinterpret_cast<float4*>(x_out), reinterpret_cast<float4*>(y), mean, rstd);
    };
    if (b && b_is_bf16) launch(train::add_ln_fwd_kernel<true>); else launch(train::add_ln_fwd_kernel<false>);
    return check_launch("rmp_add_layernorm_fwd");
}

extern "C" int rmp_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int rows,
                                 float* dx, void* dx_bf16, float* dgamma, float* dbeta, void* stream) {
    RMP_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && rows >= 0, RMP_E_ARG, "rmp_layernorm_bwd: null pointer");
    RMP_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(gamma) && aligned16(dx) && (!dx_bf16 || aligned16(dx_bf16)), RMP_E_ALIGN,
                "rmp_layernorm_bwd: 16-byte alignment");
    if (rows == 0) return RMP_OK;
    const int rows_per_block = 64;
    train::ln_bwd_kernel<<<ceil_div(rows, rows_per_block), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x), mean, rstd, reinterpret_cast<const float4*>(gamma),
        rows, rows_per_block, reinterpret_cast<float4*>(dx), reinterpret_cast<uint2*>(dx_bf16), dgamma, dbeta);
    return check_launch("rmp_layernorm_bwd");
}

extern "C" int rmp_colsum(const void* a, int a_is_bf16, const float* x, const float* mean, const float* rstd, int rows, int width,
                          float* s0, float* s1, void* stream) {
    RMP_REQUIRE(a && s0 && rows >= 0, RMP_E_ARG, "rmp_colsum: null pointer");
    RMP_REQUIRE((x == nullptr) == (s1 == nullptr) && (!x || (mean && rstd)), RMP_E_ARG, "rmp_colsum: x, mean, rstd, s1 go together");
    RMP_REQUIRE(width >= 4 && width <= 1024 && width % 4 == 0 && 256 % (width / 4) == 0, RMP_E_SHAPE,
                "rmp_colsum: width must be 4 * (a divisor of 256), got %d", width);
    RMP_REQUIRE(aligned16(a) && aligned16(s0) && (!x || (aligned16(x) && aligned16(s1))), RMP_E_ALIGN, "rmp_colsum: 16-byte alignment");
    if (rows == 0) return RMP_OK;
    const int rows_per_block = 64;
    const unsigned grid = (unsigned)ceil_div(rows, rows_per_block);
    if (a_is_bf16)
        train::colsum_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a, x, mean, rstd, rows, width, rows_per_block, s0, s1);
    else
        train::colsum_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a, x, mean, rstd, rows, width, rows_per_block, s0, s1);
    return check_launch("rmp_colsum");
}

extern "C" int rmp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float grad_scale,
                             float lr, float beta1, float beta2, float eps, int step, const int* step_dev, void* stream) {
    RMP_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && (step >= 1 || step_dev), RMP_E_ARG, "rmp_adam_step: bad arguments");
    RMP_REQUIRE(n % 4 == 0 && aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), RMP_E_ALIGN,
                "rmp_adam_step: buffers must be 16-byte aligned and a multiple of 4 elements (pad the flat buffer)");
    if (n == 0) return RMP_OK;
    const float bc1 = 1.0f - powf(beta1, (float)(step < 1 ? 1 : step)), bc2 = 1.0f - powf(beta2, (float)(step < 1 ? 1 : step));
    const size_t n4 = (size_t)n / 4;
    train::adam_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(exp_avg),
        reinterpret_cast<float4*>(exp_avg_sq), n4, grad_scale, lr, beta1, beta2, eps, bc1, sqrtf(bc2), step_dev);
    return check_launch("rmp_adam_step");
}

extern "C" int rmp_gemm_dw_bf16(const void* dZ_bf16, const void* X_bf16, float* dW, int rows, void* stream) {
    RMP_REQUIRE(dZ_bf16 && X_bf16 && dW && rows >= 0, RMP_E_ARG, "rmp_gemm_dw_bf16: null pointer");
    RMP_REQUIRE(aligned16(dZ_bf16) && aligned16(X_bf16) && aligned16(dW), RMP_E_ALIGN, "rmp_gemm_dw_bf16: 16-byte alignment");
    if (rows == 0) return RMP_OK;
    CUtensorMap tm_dz, tm_x;
    int rc = tc::make_map_bf16_rows128(&tm_dz, dZ_bf16, (uint64_t)rows);
    if (rc) return rc;
    rc = tc::make_map_bf16_rows128(&tm_x, X_bf16, (uint64_t)rows);
    if (rc) return rc;
    const size_t smem = 1024 + 2 * (size_t)train::kDwStages * tc::kTileBytes + sizeof(train::DwBarriers) + 16;
    cudaFuncSetAttribute(train::gemm_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(rows, tc::kTileM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    train::gemm_dw_kernel<<<grid, train::kDwThreads, smem, (cudaStream_t)stream>>>(tm_dz, tm_x, dW, rows);
    return check_launch("rmp_gemm_dw_bf16");
}

extern "C" int rmp_gemm_rows_bf16(const void* A_bf16, const void* W_bf16, const float* bias, const void* Pc, const void* Pn,
                                  const int32_t* ctr, const int32_t* nbr, void* z_bf16, void* a_bf16, float* acc32, int rows,
                                  int w_ld, int w_transposed, float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id,
                                  void* stream) {
    RMP_REQUIRE(A_bf16 && W_bf16 && rows >= 0 && (z_bf16 || a_bf16 || acc32), RMP_E_ARG, "rmp_gemm_rows_bf16: null pointer / no output");
    RMP_REQUIRE((Pc == nullptr) == (Pn == nullptr) && (!Pc || (ctr && nbr)), RMP_E_ARG, "rmp_gemm_rows_bf16: Pc, Pn, ctr, nbr go together");
    RMP_REQUIRE(aligned16(A_bf16) && aligned16(W_bf16) && (!bias || aligned16(bias)) && (!Pc || (aligned32(Pc) && aligned32(Pn))) &&
                (!z_bf16 || aligned32(z_bf16)) && (!a_bf16 || aligned32(a_bf16)) && (!acc32 || aligned32(acc32)), RMP_E_ALIGN,
                "rmp_gemm_rows_bf16: inputs 16-byte, outputs 32-byte aligned");
    RMP_REQUIRE(p_drop >= 0.f && p_drop < 1.f, RMP_E_ARG, "rmp_gemm_rows_bf16: dropout probability in [0,1)");
    if (rows == 0) return RMP_OK;
    CUtensorMap tm_a, tm_w;
    int rc = tc::make_map_bf16_rows128(&tm_a, A_bf16, (uint64_t)rows);
    if (rc) return rc;
    {   // W: 128 x 128 window of a row-major matrix with w_ld elements per row (a column block of a wider weight)
        tc::EncodeTiledFn fn = tc::encode_fn();
        RMP_REQUIRE(fn != nullptr, RMP_E_ARCH, "cuTensorMapEncodeTiled entry point not found");
        const int ld = w_ld > 0 ? w_ld : 128;
        RMP_REQUIRE(ld >= 128 && ld % 8 == 0, RMP_E_SHAPE, "rmp_gemm_rows_bf16: w_ld must be >= 128 and a multiple of 8");
        const cuuint64_t dims[2] = {128, 128};
        const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
        const cuuint32_t box[2] = {64, 128};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = fn(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(W_bf16), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RMP_REQUIRE(r == CUDA_SUCCESS, RMP_E_ARG, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
    train::GrParams p{};
    p.bias = bias; p.Pc = reinterpret_cast<const uint16_t*>(Pc); p.Pn = reinterpret_cast<const uint16_t*>(Pn); p.ctr = ctr; p.nbr = nbr;
    p.z16 = reinterpret_cast<uint16_t*>(z_bf16); p.a16 = reinterpret_cast<uint16_t*>(a_bf16); p.acc32 = acc32; p.rows = rows;
    p.thresh16 = drop_threshold(p_drop); p.keep_scale = keep_scale_of(p_drop); p.seed = seed; p.seed_dev = seed_dev;
    p.stream = stream_id; p.w_mn_major = w_transposed ? 1 : 0; p.k_blocks = 1;
    CUtensorMap tm_z, tm_act;
    memset(&tm_z, 0, sizeof(tm_z)); memset(&tm_act, 0, sizeof(tm_act));
    if (z_bf16) { rc = tc::make_map_bf16_rows128(&tm_z, z_bf16, (uint64_t)rows); if (rc) return rc; }
    if (a_bf16) { rc = tc::make_map_bf16_rows128(&tm_act, a_bf16, (uint64_t)rows); if (rc) return rc; }
    const size_t smem = 1024 + (size_t)(3 + train::kGrStages) * tc::kTileBytes + sizeof(train::GrBarriers) + 16;
    cudaFuncSetAttribute(train::gemm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(rows, tc::kTileM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    train::gemm_rows_kernel<<<grid, train::kGrThreads, smem, (cudaStream_t)stream>>>(tm_a, tm_w, tm_z, tm_act, p);
    return check_launch("rmp_gemm_rows_bf16");
}

extern "C" int rmp_gemm_rows_kcat_bf16(const void* A_bf16, const void* W_bf16, int k_blocks, float* out32, int rows,
                                       int w_transposed, void* stream) {
    RMP_REQUIRE(A_bf16 && W_bf16 && out32 && rows >= 0, RMP_E_ARG, "rmp_gemm_rows_kcat_bf16: null pointer");
    RMP_REQUIRE(k_blocks >= 2 && k_blocks <= 64, RMP_E_SHAPE, "rmp_gemm_rows_kcat_bf16: 2..64 K blocks");
    RMP_REQUIRE((int64_t)k_blocks * rows < (int64_t)1 << 31, RMP_E_SHAPE, "rmp_gemm_rows_kcat_bf16: k_blocks * rows exceeds int32");
    RMP_REQUIRE(aligned16(A_bf16) && aligned16(W_bf16) && aligned32(out32), RMP_E_ALIGN,
                "rmp_gemm_rows_kcat_bf16: inputs 16-byte, output 32-byte aligned");
    if (rows == 0) return RMP_OK;
    CUtensorMap tm_a, tm_w, tm_z, tm_act;
    int rc = tc::make_map_bf16_rows128(&tm_a, A_bf16, (uint64_t)k_blocks * (uint64_t)rows);
    if (rc) return rc;
    rc = tc::make_map_bf16_rows128(&tm_w, W_bf16, (uint64_t)k_blocks * 128);
    if (rc) return rc;
    memset(&tm_z, 0, sizeof(tm_z)); memset(&tm_act, 0, sizeof(tm_act));
    train::GrParams p{};
    p.acc32 = out32; p.acc32_store = 1; p.rows = rows; p.w_mn_major = w_transposed ? 1 : 0;
    p.k_blocks = k_blocks; p.a_block_rows = rows; p.keep_scale = 1.f;
    const size_t smem = 1024 + (size_t)(3 + train::kGrStages) * tc::kTileBytes + sizeof(train::GrBarriers) + 16;
    cudaFuncSetAttribute(train::gemm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_tiles = ceil_div(rows, tc::kTileM);
    const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
    train::gemm_rows_kernel<<<grid, train::kGrThreads, smem, (cudaStream_t)stream>>>(tm_a, tm_w, tm_z, tm_act, p);
    return check_launch("rmp_gemm_rows_kcat_bf16");
}

extern "C" int rmp_normalize_fwd_bf16(const float* x, const float* gain, const float* bias, void* y_bf16, int rows, float eps,
                                      void* stream) {
    RMP_REQUIRE(x && gain && bias && y_bf16 && rows >= 0, RMP_E_ARG, "rmp_normalize_fwd_bf16: null pointer");
    RMP_REQUIRE(aligned16(x) && aligned16(gain) && aligned16(bias) && aligned16(y_bf16), RMP_E_ALIGN, "rmp_normalize_fwd_bf16: 16-byte alignment");
    if (rows == 0) return RMP_OK;
    train::normalize_fwd_kernel<<<ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(gain), reinterpret_cast<const float4*>(bias),
        reinterpret_cast<uint2*>(y_bf16), rows, eps);
    return check_launch("rmp_normalize_fwd_bf16");
}

extern "C" int rmp_normalize_bwd_bf16(const float* x, const float* dy, const float* gain, void* dx_bf16, float* dgain, float* dbias,
                                      float* dbsum, int rows, float eps, void* stream) {
    RMP_REQUIRE(x && dy && gain && dx_bf16 && dgain && dbias && dbsum && rows >= 0, RMP_E_ARG, "rmp_normalize_bwd_bf16: null pointer");
    RMP_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(gain) && aligned16(dx_bf16), RMP_E_ALIGN, "rmp_normalize_bwd_bf16: 16-byte alignment");
    if (rows == 0) return RMP_OK;
    const int rows_per_block = 256;
    train::normalize_bwd_kernel<<<ceil_div(rows, rows_per_block), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(gain),
        reinterpret_cast<uint2*>(dx_bf16), dgain, dbias, dbsum, rows, rows_per_block, eps);
    return check_launch("rmp_normalize_bwd_bf16");
}
