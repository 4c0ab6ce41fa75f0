// Round-2 TMEM / epilogue microbenchmark (replaces scripts/tmem_bench.cu and scripts/tmem_shape_bench.cu, whose
// `r[it & 31]` consumption let nvcc keep the loaded registers in local memory).  Every loaded register is consumed
// through a static xor tree (fully unrolled, compile-time indices) and tied to tcgen05.wait::ld; the build recipe
// (scripts/micro/run_tmem_bench2.sh) greps the SASS for STL/LDL and prints the count next to the results.
//
// One CTA per SM (148), 4 / 8 / 16 warps, every warp runs `iters` iterations of the mode's body; reported:
// cycles per iteration of CTA 0 (clock64) and TMEM bytes per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include "tmem_wrappers.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int N>
__device__ __forceinline__ uint32_t xor_all(const uint32_t (&r)[N]) {
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) a ^= r[i];
    return a;
}
// shipped packed GELU (tc_common.cuh: gelu_bf16x2)
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
// 0.5 folded into the next layer's weights: y = x + x * tanh(..)  (one multiply less per pair)
__device__ __forceinline__ uint32_t gelu2_bf16x2(float a, float b) {
    const __nv_bfloat162 x = __floats2bfloat162_rn(a, b);
    const __nv_bfloat162 c1 = __float2bfloat162_rn(0.0356774081f), c0 = __float2bfloat162_rn(0.7978845608f);
    const __nv_bfloat162 inner = __hmul2(x, __hfma2(__hmul2(x, x), c1, c0));
    uint32_t t;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&inner)));
    const __nv_bfloat162 y = __hfma2(x, *reinterpret_cast<const __nv_bfloat162*>(&t), x);
    return *reinterpret_cast<const uint32_t*>(&y);
}

enum { LD32_XOR, LD32_LIGHT, LD32X2_XOR, LD16_XOR, LD64_XOR, ST32, ST16, EPI_BODY, GELU_ONLY, SPLIT_LD_GELU, EPI_PIPE,
       EPI_BODY_FOLD };

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, int iters) {
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t c32 = base + (uint32_t)((warp >> 2) * 32);          // this warp's 32 fp32 columns
    const uint32_t y16 = base + 384u + (uint32_t)((warp >> 2) * 16);   // its 16 packed bf16 columns
    // defined TMEM contents: small floats everywhere this kernel reads
    {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(0.01f * (float)((threadIdx.x * 7 + i * 13) % 200) - 1.0f);
        for (int c = 0; c < 512; c += 32) tm_st_x32(base + c, v);
        wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t acc = threadIdx.x;
    const long long t0 = clock64();
    if (MODE == LD32_XOR) {
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) { uint32_t r[32]; tm_ld_x32(c32, r); tm_wait_ld_x32(r); acc ^= xor_all(r); }
    } else if (MODE == LD32_LIGHT) {
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) { uint32_t r[32]; tm_ld_x32(c32, r); tm_wait_ld_x32(r); acc ^= r[0] ^ r[31]; }
    } else if (MODE == LD32X2_XOR) {
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) {
            uint32_t r[32], q[32];
            tm_ld_x32(c32, r); tm_ld_x32(c32 + 128, q);
            tm_wait_ld_x32(r); tm_wait_ld_x32(q);
            acc ^= xor_all(r) ^ xor_all(q);
        }
    } else if (MODE == LD16_XOR) {
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) { uint32_t r[16]; tm_ld_x16(c32, r); tm_wait_ld_x16(r); acc ^= xor_all(r); }
    } else if (MODE == LD64_XOR) {
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) { uint32_t r[64]; tm_ld_x64(base + (uint32_t)((warp >> 2) * 64), r); tm_wait_ld_x64(r); acc ^= xor_all(r); }
    } else if (MODE == ST32) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = acc + i;
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) { v[0] ^= it; v[1] += it; tm_st_x32(c32, v); wait_st(); }
        acc ^= v[0] ^ v[1];
    } else if (MODE == ST16) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = acc + i;
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) { v[0] ^= it; v[1] += it; tm_st_x16(y16, v); wait_st(); }
        acc ^= v[0] ^ v[1];
    } else if (MODE == EPI_BODY || MODE == EPI_BODY_FOLD) {
        // the edge epilogue's inner body: accumulator -> GELU -> packed bf16 A operand back to TMEM
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) {
            uint32_t r[32], o[16];
            tm_ld_x32(c32, r); tm_wait_ld_x32(r);
#pragma unroll
            for (int q = 0; q < 16; ++q)
                o[q] = MODE == EPI_BODY ? gelu_bf16x2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]))
                                        : gelu2_bf16x2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
            tm_st_x16(y16, o); wait_st();
            acc ^= o[0];
        }
    } else if (MODE == GELU_ONLY) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.01f * (float)((threadIdx.x + i * 13) % 200) - 1.0f;
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) {
            uint32_t o[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) o[q] = gelu_bf16x2(v[2 * q], v[2 * q + 1]);
#pragma unroll
            for (int q = 0; q < 16; ++q) {      // feed the results back so that nothing is hoisted or dropped
                v[2 * q] = __uint_as_float(o[q] << 16) + 0.25f;
                v[2 * q + 1] = __uint_as_float(o[q] & 0xffff0000u) - 0.25f;
            }
        }
        acc ^= __float_as_uint(v[0]) ^ __float_as_uint(v[17]);
    } else if (MODE == SPLIT_LD_GELU) {
        // half of the warps only read TMEM, the other half only run GELU on registers: do the two overlap?
        if (((warp >> 2) & 1) == 0) {
            _Pragma("unroll 1") for (int it = 0; it < iters; ++it) { uint32_t r[32]; tm_ld_x32(c32, r); tm_wait_ld_x32(r); acc ^= r[0] ^ r[31]; }
        } else {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.01f * (float)((threadIdx.x + i * 13) % 200) - 1.0f;
            _Pragma("unroll 1") for (int it = 0; it < iters; ++it) {
                uint32_t o[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) o[q] = gelu_bf16x2(v[2 * q], v[2 * q + 1]);
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    v[2 * q] = __uint_as_float(o[q] << 16) + 0.25f;
                    v[2 * q + 1] = __uint_as_float(o[q] & 0xffff0000u) - 0.25f;
                }
            }
            acc ^= __float_as_uint(v[0]) ^ __float_as_uint(v[17]);
        }
    } else if (MODE == EPI_PIPE) {
        // software-pipelined epilogue body: the next accumulator chunk is in flight while this one goes through GELU
        uint32_t r[32], n[32];
        tm_ld_x32(c32, r); tm_wait_ld_x32(r);
        _Pragma("unroll 1") for (int it = 0; it < iters; ++it) {
            tm_ld_x32(c32 + 128, n);
            uint32_t o[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) o[q] = gelu_bf16x2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
            tm_st_x16(y16, o);
            tm_wait_ld_x32(n);
#pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = n[q];
            wait_st();
            acc ^= o[0];
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 512 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, int ld_bytes_per_warp_iter) {
    uint32_t* d; long long* c;
    cudaMalloc(&d, 148 * 512 * 4);
    cudaMalloc(&c, 148 * 8);
    printf("%-44s", name);
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 2000;
        k<MODE><<<148, warps * 32>>>(d, c, 10);                 // warm-up (instruction cache)
        k<MODE><<<148, warps * 32>>>(d, c, iters);
        long long cyc[148];
        cudaMemcpy(cyc, c, sizeof(cyc), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
        const double per = (double)mx / iters;
        printf(" | %2dw %7.1f cyc/iter", warps, per);
        if (ld_bytes_per_warp_iter) printf(" %6.0f B/clk/SM", (double)ld_bytes_per_warp_iter * warps / per);
    }
    printf("\n");
    cudaFree(d); cudaFree(c);
}
int main() {
    run<LD32_XOR>("ld x32 (4 KB/warp) + wait, xor of all 32", 4096);
    run<LD32_LIGHT>("ld x32 + wait, r[0]^r[31]", 4096);
    run<LD32X2_XOR>("2 x ld x32 in flight + waits, xor of all 64", 8192);
    run<LD16_XOR>("ld x16 (2 KB/warp) + wait, xor of all 16", 2048);
    run<LD64_XOR>("ld x64 (8 KB/warp) + wait, xor of all 64", 8192);
    run<ST32>("st x32 (4 KB/warp) + wait::st", 4096);
    run<ST16>("st x16 (2 KB/warp) + wait::st", 2048);
    run<GELU_ONLY>("GELU bf16x2 of 32 values (registers only)", 0);
    run<EPI_BODY>("ld x32 -> GELU -> st x16 (edge epilogue body)", 0);
    run<EPI_BODY_FOLD>("ld x32 -> GELU (0.5 folded) -> st x16", 0);
    run<EPI_PIPE>("same, next ld in flight during GELU", 0);
    run<SPLIT_LD_GELU>("half the warps ld x32, half GELU-only", 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
