// Device-side collate: packed per-residue coordinates -> the padded batch the
// featurizer consumes.  Replaces the host loop of rdesign/utils/data.py:84-119 (featurize: first six atoms,
// NaN -> 0 before padding, zero padding, prefix mask, float64 -> float32) and rnampnn/utils/data.py:110-142.
// HBM-bound byte shuffling: one thread per output element, coalesced stores.
#include "common.cuh"

namespace rmp {

template <typename T>
__global__ void __launch_bounds__(256) collate_kernel(const T* __restrict__ packed, const int64_t* __restrict__ row_start,
                                                     const int32_t* __restrict__ lengths, int B, int N, int A_in,
                                                     int A_out, float* __restrict__ X, float* __restrict__ mask) {
    const size_t per_res = (size_t)A_out * 3;
    const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (size_t)B * N * per_res) return;
    const size_t res = idx / per_res;
    const int c = (int)(idx % per_res);                       // atom * 3 + xyz, atoms 0..A_out-1 come first in the source
    const int b = (int)(res / N), i = (int)(res % N);
    const int L = lengths[b];
    float v = 0.f;
    if (i < L) {
        const T s = packed[((size_t)row_start[b] + i) * A_in * 3 + c];
        v = (float)s;
        if (v != v) v = 0.f;                                  // np.nan_to_num(nan=0.0): +-inf are left alone
    }
    X[idx] = v;
    if (c == 0) mask[res] = i < L ? 1.f : 0.f;
}

}  // namespace rmp

using namespace rmp;

extern "C" int rmp_collate(const void* packed, int packed_is_f64, const int64_t* row_start, const int32_t* lengths,
                           int B, int N, int A_in, int A_out, float* X, float* mask, void* stream) {
    RMP_REQUIRE(packed && row_start && lengths && X && mask, RMP_E_ARG, "rmp_collate: null pointer");
    RMP_REQUIRE(B > 0 && N > 0 && A_out > 0 && A_out <= A_in, RMP_E_ARG, "rmp_collate: bad sizes");
    const size_t total = (size_t)B * N * A_out * 3;
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (packed_is_f64)
        collate_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(packed), row_start, lengths, B, N,
                                                     A_in, A_out, X, mask);
    else
        collate_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(packed), row_start, lengths, B, N, A_in,
                                                    A_out, X, mask);
    return check_launch("rmp_collate");
}
