// Shared helpers for the rmpnn_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "rmpnn_b200.h"

namespace rmp {

void set_error(const char* fmt, ...);
void note_launches(int n);      // bumps the counter behind rmp_launch_count()

inline int check_launch(const char* what, int kernels = 1) {
    note_launches(kernels);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return RMP_OK;
}

#define RMP_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            rmp::set_error(__VA_ARGS__);        \
            return (code);                      \
        }                                       \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
// pointers touched by 256-bit vector accesses (ld/st.global.v8)
inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- warp helpers -----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// packed fp32x2 FMA (FFMA2 on sm_100): d = a * b + c on two lanes of a 64-bit register pair, each rounded like fmaf
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t"
        ".reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

// exact (erf) GELU, nn.GELU() default (rdesign/model/mpnn.py:17)
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

}  // namespace rmp
