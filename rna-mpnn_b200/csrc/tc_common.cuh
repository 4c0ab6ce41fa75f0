// tcgen05 / TMEM / TMA / mbarrier PTX wrappers, UMMA descriptors and small epilogue helpers shared
// by the bf16 tensor-core kernels (layer_tc.cu, embed_tc.cu).  sm_100a only.
#pragma once
#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <type_traits>

namespace rmp {
namespace tc {

constexpr int kTileM = 128;                 // edges per tile (UMMA M)
constexpr int kKBlockBytes = kTileM * 128;  // one [128 rows][64 bf16] swizzle-128B block
constexpr int kTileBytes = 2 * kKBlockBytes;
constexpr int kMaxStages = 4;
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr uint32_t kTmemCols = 256;         // [0,128) stage accumulator, [128,256) neighbour sums

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// One thread of a CONVERGED warp (elect.sync).  With `if (lane == 0)` ptxas cannot tell that a single thread is active and
// wraps every warp-level instruction (UTCHMMA, UTMALDG, UTCBAR) in an ELECT / BRA.U.ANY loop over the active lanes: ~80
// cycles per tcgen05.mma in the issuing thread (700-1,100 per 8-step GEMM, profiles/r2_edge_issue_cost.txt); with the
// elect.sync predicate the instructions are emitted back to back.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, bf16 operands, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, 128-byte swizzle, Blackwell version bit (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor kind::f16: D fp32, A/B bf16, A K-major, B K-major or MN-major, M x N tile.
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk `j` (8 bf16) of row `r` inside one [128][64] swizzle-128B block
__device__ __forceinline__ uint32_t sw128_chunk(int r, int j) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}

// GELU for the bf16 arm: tanh form on the hardware tanh unit (MUFU.TANH), 5 FP32 ops + 1 MUFU.
// |tanh-form - erf-form| <= 5e-4 abs, below the bf16 rounding (2^-9 relative) applied to the result;
// the fp32 arm (rowmlp.cu) keeps the exact erf form.
__device__ __forceinline__ float gelu_tanh_fast(float x) {
    const float x2 = x * x;
    const float inner = x * fmaf(x2, 0.0356774081f, 0.7978845608f);     // sqrt(2/pi) * (x + 0.044715 x^3)
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

// Packed GELU of two fp32 values -> bf16x2 (tanh form, bf16 FMAs + MUFU.TANH.BF16).  The result feeds a
// bf16 tensor-core operand, so evaluating the polynomial in bf16 adds ~1 extra bf16 ulp to a value that
// is rounded to bf16 anyway; measured end-to-end |dlogit| stays < 4e-3 (budget 2e-2).
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
// GELU + pack of 32 consecutive columns, written as the next stage's swizzled A operand
__device__ __forceinline__ void store_gelu_chunk(uint8_t* a_buf, int row, int col0, const float (&v)[32]);

// 256-bit read-only global load (LDG.E.256): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void ldg256(const void* ptr, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(ptr));
}

__device__ __forceinline__ void ldg256f(const float* ptr, float (&r)[8]) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "l"(ptr));
}
__device__ __forceinline__ void stg256f(float* ptr, const float* r) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "f"(r[0]), "f"(r[1]), "f"(r[2]),
                 "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
                 : "memory");
}
__device__ __forceinline__ void stg256u(void* ptr, const uint32_t* r) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// fire-and-forget fp32 vector reduction into global memory (REDG.E.ADD.F32x4)
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
    asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void store_gelu_chunk(uint8_t* a_buf, int row, int col0, const float (&v)[32]) {
    uint8_t* kb = a_buf + (col0 >> 6) * kKBlockBytes;
    const int j0 = (col0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(kb + sw128_chunk(row, j0 + j)) =
            make_uint4(gelu_bf16x2(v[8 * j + 0], v[8 * j + 1]), gelu_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                       gelu_bf16x2(v[8 * j + 4], v[8 * j + 5]), gelu_bf16x2(v[8 * j + 6], v[8 * j + 7]));
}

__device__ __forceinline__ void store_act_chunk(uint8_t* a_buf, int row, int col0, const float (&v)[32]) {
    // 32 consecutive columns starting at col0 -> K-block col0/64, 16-byte chunks (col0%64)/8 ..
    uint8_t* kb = a_buf + (col0 >> 6) * kKBlockBytes;
    const int j0 = (col0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint4 pk = make_uint4(pack_bf16(v[8 * j + 0], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                    pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        *reinterpret_cast<uint4*>(kb + sw128_chunk(row, j0 + j)) = pk;
    }
}


__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// wait used by the role threads that are not latency critical: poll, then sleep, so the single
// spinning lane does not take issue slots from the epilogue warps of its SM sub-partition
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned ns) {
    while (!mbar_test(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void group_bar_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kEpiThreads) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// Instruction descriptor with A MN-major (the [edge][channel] activation read as A = m3^T), B K-major
__host__ __device__ constexpr uint32_t instr_desc_bf16_amn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// ------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// bf16 row-major [rows][cols] tensor, box = 64 columns x box_rows rows, 128-byte swizzle
inline int make_map_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not found"); return RMP_E_ARCH; }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {64, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return RMP_E_ARG; }
    return RMP_OK;
}
inline int make_map_bf16_rows128(CUtensorMap* map, const void* base, uint64_t rows) {
    return make_map_bf16(map, base, rows, 128, 128);
}


}  // namespace tc
}  // namespace rmp
