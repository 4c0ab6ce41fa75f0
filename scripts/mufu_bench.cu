// Tuning aid: per-SM throughput of the activation primitives the tcgen05 epilogues can use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench scripts/mufu_bench.cu && /tmp/mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, float seed) {
    // 8 independent chains per thread so the pipe, not the latency, is measured
    float f[8];
    uint32_t u[8];
    for (int i = 0; i < 8; ++i) { f[i] = seed + 0.01f * (threadIdx.x + i); u[i] = __float_as_uint(f[i]) | 0x3c003c00u; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
            if (MODE == 1) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u[i]));
            if (MODE == 2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u[i]));
            if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            if (MODE == 4) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
            if (MODE == 5) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
            if (MODE == 6) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(u[i]));
            if (MODE == 7) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[i]));
            if (MODE == 8) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(u[i]));
            if (MODE == 9) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            if (MODE == 10) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "+r"(u[i]) : "f"(f[i]));
            if (MODE == 11) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(*reinterpret_cast<unsigned long long*>(&u[i & 6])));
        }
    }
    uint32_t acc = 0;
    for (int i = 0; i < 8; ++i) acc ^= u[i] ^ __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int elems_per_op) {
    uint32_t* d;
    cudaMalloc(&d, 148 * 4 * 512 * 4);
    const int iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148, 512>>>(d, 16, 0.5f);
    cudaEventRecord(a);
    k<MODE><<<148, 512>>>(d, iters, 0.5f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    // warp-instructions per SM = 16 warps * iters * 8; assume 1.965 GHz
    const double cyc = ms * 1e-3 * 1.965e9;
    const double winst = 16.0 * iters * 8;
    printf("%-28s %8.3f ms  %6.2f cyc/warp-inst/SM  -> %6.1f lanes/clk/SM, %6.1f elems/clk/SM\n", name, ms, cyc / winst,
           32.0 * winst / cyc, 32.0 * elems_per_op * winst / cyc);
    cudaFree(d);
}

int main() {
    run<0>("tanh.approx.f32", 1);
    run<1>("tanh.approx.bf16x2", 2);
    run<2>("tanh.approx.f16x2", 2);
    run<3>("ex2.approx.ftz.f32", 1);
    run<4>("ex2.approx.ftz.bf16x2", 2);
    run<5>("ex2.approx.f16x2", 2);
    run<6>("fma.rn.bf16x2", 2);
    run<7>("fma.rn.f32", 1);
    run<8>("fma.rn.f16x2", 2);
    run<9>("rcp.approx.ftz.f32", 1);
    run<10>("cvt.rn.bf16x2.f32", 2);
    run<11>("fma.rn.f32x2", 2);
    return 0;
}
