// Counter-based dropout mask shared by the training kernels (train.cu) and the fused training forward of the edge kernel
// (layer_tc.cu): the backward regenerates exactly the forward's mask from (seed, stream id, element group).
#pragma once
#include <cstdint>

namespace rmp {

// splitmix finalizer (stateless, so the backward regenerates exactly the forward's mask; ~25 integer instructions per
// 8 elements where Philox-4x32-10 needs ~80, which made the activation passes instruction-bound).
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}
// keep-mask of the 8 elements of group g: one hash, eight 8-bit lanes; bit j set = element j survives.
// The drop probability is quantised to thresh8 / 256 and the survivors are scaled by 256 / (256 - thresh8), so the
// expectation is exact for the quantised p (0.1 -> 26/256 = 0.1016).
__device__ __forceinline__ uint32_t keep_mask8(uint64_t seed, uint64_t g, uint32_t stream, uint32_t thresh8) {
    if (thresh8 == 0) return 0xffu;
    const uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull * (g + 1) + ((uint64_t)stream << 40));
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        m |= (uint32_t)(((lo >> (8 * j)) & 0xffu) >= thresh8) << j;
        m |= (uint32_t)(((hi >> (8 * j)) & 0xffu) >= thresh8) << (4 + j);
    }
    return m;
}

// The same mask as four 32-bit AND masks for the group's four packed bf16x2 pairs (pair j = elements 2j, 2j+1): a per-byte
// compare of the hash against the threshold and one byte permute per pair instead of eight bit extractions and eight
// selects.  The fused training forward of the edge kernel evaluates four of these per thread and tile-stage and is
// instruction-bound there (94.5 M warp instructions per launch against 45 M for the inference launch,
// profiles/r2s3_edge_trainfwd_full.txt): 64 instructions fewer per thread and tile-stage.
__device__ __forceinline__ void keep_masks_bf16x2(uint64_t seed, uint64_t g, uint32_t stream, uint32_t thresh8, uint32_t (&m)[4]) {
    if (thresh8 == 0) { m[0] = m[1] = m[2] = m[3] = 0xffffffffu; return; }
    const uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull * (g + 1) + ((uint64_t)stream << 40));
    const uint32_t t4 = thresh8 * 0x01010101u;
    const uint32_t klo = __vcmpgeu4((uint32_t)h, t4), khi = __vcmpgeu4((uint32_t)(h >> 32), t4);      // 0xff per surviving element
    m[0] = __byte_perm(klo, 0u, 0x1100); m[1] = __byte_perm(klo, 0u, 0x3322);
    m[2] = __byte_perm(khi, 0u, 0x1100); m[3] = __byte_perm(khi, 0u, 0x3322);
}

}  // namespace rmp
