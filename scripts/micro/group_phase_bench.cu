// Round-2 schedule microbenchmark: would more independently phased epilogue groups hide the hand-off latency of the edge
// kernel?  Emulates the kernel's epilogue side only: G groups of `WPG` warps each loop over tile-stages
//     [latency phase: a chain of dependent TMEM round trips (~Z cycles, almost no issue slots)]  ->  bar.sync(group)
//     [body: per thread 128 / (WPG / 4) accumulator columns: ld x32 -> packed tanh-GELU -> st x16, as shipped]
// and reports cycles per tile-stage of the whole CTA (148 CTAs, clock64 of the slowest).  G = 1, WPG = 16 is the shipped
// staggered schedule (all 16 warps in lockstep); G = 2 / 3 / 4 with WPG = 8 are decoupled groups (3 and 4 need
// accumulators narrower than fp32: 512 TMEM columns hold 3 x 128 + 2 x 64 today).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include "tmem_wrappers.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t gelu_bf16x2(float a, float b) {
    const __nv_bfloat162 x = __floats2bfloat162_rn(a, b);
    const __nv_bfloat162 c1 = __float2bfloat162_rn(0.0356774081f), c0 = __float2bfloat162_rn(0.7978845608f);
    const __nv_bfloat162 inner = __hmul2(x, __hfma2(__hmul2(x, x), c1, c0));
    uint32_t t;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&inner)));
    const __nv_bfloat162 hx = __hmul2(x, __float2bfloat162_rn(0.5f));
    const __nv_bfloat162 y = __hfma2(hx, *reinterpret_cast<const __nv_bfloat162*>(&t), hx);
    return *reinterpret_cast<const uint32_t*>(&y);
}
__device__ __forceinline__ uint32_t tm_ld_x1_wait(uint32_t a) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(r) : "r"(a) : "memory");
    return r;
}

template <int WPG>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, long long* cyc, int iters, int groups, int chain, int skew) {
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int g = warp / WPG, wg = warp % WPG;
    const uint32_t lane_base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int kChunks = 128 / ((WPG / 4) * 32);                 // 32-column chunks per thread and tile-stage
    const uint32_t col0 = (uint32_t)((g & 1) * 128 + (wg >> 2) * kChunks * 32);      // accumulator columns of this thread
    const uint32_t ycol = 256u + (uint32_t)((g & 1) * 64 + (wg >> 2) * kChunks * 16);
    uint32_t acc = 0;
    // optional initial skew between the groups (in latency-chain links)
    for (int s = 0; s < skew * g; ++s) acc ^= tm_ld_x1_wait(lane_base + 300u + (acc & 1u));
    const long long t0 = clock64();
    _Pragma("unroll 1") for (int it = 0; it < iters; ++it) {
        // latency phase: `chain` dependent TMEM round trips (accumulator wait, tcgen05.ld round trip, fences, arrive ...)
        _Pragma("unroll 1") for (int c = 0; c < chain; ++c) acc ^= tm_ld_x1_wait(lane_base + 300u + (acc & 1u)) & 1u;
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(WPG * 32) : "memory");
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) {
            uint32_t r[32];
            tm_ld_x32(lane_base + col0 + (uint32_t)(ch * 32), r);
            tm_wait_ld_x32(r);
            uint32_t o[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) o[q] = gelu_bf16x2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
            tm_st_x16(lane_base + ycol + (uint32_t)(ch * 16), o);
            acc ^= o[0] & 1u;
        }
        wait_st();
    }
    const long long t1 = clock64();
    out[blockIdx.x * 1024 + threadIdx.x] = acc;
    if (threadIdx.x % (WPG * 32) == 0) cyc[blockIdx.x * 4 + g] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
}

template <int WPG>
void run(int groups, int chain, int skew) {
    uint32_t* d; long long* c;
    cudaMalloc(&d, 148 * 1024 * 4);
    cudaMalloc(&c, 148 * 4 * 8);
    cudaMemset(c, 0, 148 * 4 * 8);
    const int iters = 1500;
    k<WPG><<<148, groups * WPG * 32>>>(d, c, 10, groups, chain, skew);
    k<WPG><<<148, groups * WPG * 32>>>(d, c, iters, groups, chain, skew);
    long long cyc[148 * 4];
    cudaMemcpy(cyc, c, sizeof(cyc), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148 * 4; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    cudaError_t e = cudaDeviceSynchronize();
    printf("groups %d x %2d warps, latency chain %3d links, skew %3d: %8.1f cycles per group iteration, %8.1f cycles per tile-stage%s\n",
           groups, WPG, chain, skew, (double)mx / iters, (double)mx / iters / groups, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(c);
}

int main() {
    for (int chain : {0, 20, 40}) {
        run<16>(1, chain, 0);          // shipped: 16 warps in lockstep
        run<8>(2, chain, 0);           // decoupled pairs
        run<8>(2, chain, chain / 2);
        run<8>(3, chain, 0);           // three groups (needs narrower accumulators in the real kernel)
        run<8>(3, chain, chain / 3);
        run<8>(4, chain, 0);
        run<16>(2, chain, 0);          // two full 16-warp sets
        run<16>(2, chain, chain / 2);
    }
    return 0;
}
