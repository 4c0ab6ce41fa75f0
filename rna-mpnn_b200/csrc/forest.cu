// Tree-ensemble readout head on the device: replaces the reference's
//     samples = self.xgb_readout.predict(h_V.to('cpu').detach().numpy())       rdesign/model/rdesign.py:153
// (and rnampnn/model/rnampnn.py:298): gradient-boosted trees (xgboost gbtree, numerical splits) over the 128-d node
// embeddings, multi:softmax -> class of the first maximal margin.  Semantics follow xgboost 2.1's CPU predictor
// (restated in oracle/forest_oracle.py): go left iff x[feature] < threshold in fp32, NaN follows default_left, a leaf's
// value is its threshold field, a group's margin is base_score + its trees' leaves accumulated in fp32 IN TREE ORDER
// (so margins are bit-identical to the sequential predictor, and so is the arg max on near ties).
//
// One warp per row: the row's features are staged in shared memory, lanes take trees t = lane, lane + 32, ... (the walks
// are independent dependent-load chains through L1/L2-resident 16-byte nodes; 400 trees x 127 nodes = 0.8 MB), leaf values
// go to a per-warp shared buffer, then lane c adds the leaves of group c in tree order.  Latency-bound pointer chasing: the
// grid is sized to fill every SM with 8-warp CTAs and rows are dealt round-robin to warps.
#include "common.cuh"

namespace rmp {

constexpr int kForestWarps = 8;

__global__ void __launch_bounds__(kForestWarps * 32)
forest_predict_kernel(const float* __restrict__ X, int rows, int ld, int n_feat, const uint4* __restrict__ nodes,
                      const int32_t* __restrict__ tree_root, int n_trees, const int32_t* __restrict__ class_ptr,
                      const int32_t* __restrict__ class_trees, int n_class, float base_score,
                      float* __restrict__ margins, int64_t* __restrict__ pred, int warps_per_cta) {
    extern __shared__ float fsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= warps_per_cta) return;
    float* s_x = fsm + (size_t)warp * (n_feat + n_trees);
    float* s_leaf = s_x + n_feat;
    const int warp_global = blockIdx.x * warps_per_cta + warp, n_warps = gridDim.x * warps_per_cta;
    for (int r = warp_global; r < rows; r += n_warps) {
        const float* x = X + (size_t)r * ld;
        for (int f = lane; f < n_feat; f += 32) s_x[f] = __ldg(x + f);
        __syncwarp();
        for (int t = lane; t < n_trees; t += 32) {
            int node = __ldg(tree_root + t);
            uint4 nd = __ldg(nodes + node);
            while ((int)nd.z >= 0) {
                const float fv = s_x[nd.y & 0x7fffffffu];
                const float thr = __uint_as_float(nd.x);
                const bool go_left = (fv != fv) ? (nd.y >> 31) != 0 : fv < thr;
                node = go_left ? (int)nd.z : (int)nd.w;
                nd = __ldg(nodes + node);
            }
            s_leaf[t] = __uint_as_float(nd.x);
        }
        __syncwarp();
        float best = 0.f;
        int best_c = 0;
        for (int c0 = 0; c0 < n_class; c0 += 32) {          // lane = output group; tree order inside a group
            const int c = c0 + lane;
            float m = base_score;
            if (c < n_class) {
                const int e = __ldg(class_ptr + c + 1);
                for (int j = __ldg(class_ptr + c); j < e; ++j) m = __fadd_rn(m, s_leaf[__ldg(class_trees + j)]);
                if (margins) margins[(size_t)r * n_class + c] = m;
            }
            // first maximal margin over the groups seen so far (ties -> lower class, like a sequential scan)
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                const float ml = __shfl_sync(0xffffffffu, m, l);
                if (c0 + l < n_class && ((c0 + l == 0) || ml > best)) { best = ml; best_c = c0 + l; }
            }
        }
        if (lane == 0 && pred) pred[r] = best_c;
        __syncwarp();
    }
}

}  // namespace rmp

extern "C" int rmp_forest_predict(const float* X, int rows, int ld, int n_feat, const void* nodes, int n_nodes,
                                  const int32_t* tree_root, int n_trees, const int32_t* class_ptr,
                                  const int32_t* class_trees, int n_class, float base_score, float* margins,
                                  int64_t* pred, void* stream) {
    using namespace rmp;
    RMP_REQUIRE(rows >= 0 && ld >= n_feat && n_feat >= 1 && n_trees >= 1 && n_class >= 1 && n_nodes >= n_trees, RMP_E_ARG,
                "rmp_forest_predict: bad sizes (rows %d, ld %d, n_feat %d, n_trees %d, n_class %d)", rows, ld, n_feat,
                n_trees, n_class);
    RMP_REQUIRE(X && nodes && tree_root && class_ptr && class_trees && (margins || pred), RMP_E_ARG,
                "rmp_forest_predict: null pointer");
    RMP_REQUIRE(aligned16(nodes), RMP_E_ALIGN, "rmp_forest_predict: nodes must be 16-byte aligned");
    if (rows == 0) return RMP_OK;
    const size_t per_warp = ((size_t)n_feat + n_trees) * sizeof(float);
    int warps = kForestWarps;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    RMP_REQUIRE(per_warp * warps <= 200 * 1024, RMP_E_SHAPE,
                "rmp_forest_predict: %d features + %d trees do not fit one warp's shared-memory slice", n_feat, n_trees);
    const size_t smem = per_warp * warps;
    cudaFuncSetAttribute(forest_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
    int grid = ceil_div(rows, warps);
    if (grid > kNumSMs * per_sm) grid = kNumSMs * per_sm;
    forest_predict_kernel<<<grid, kForestWarps * 32, smem, (cudaStream_t)stream>>>(
        X, rows, ld, n_feat, reinterpret_cast<const uint4*>(nodes), tree_root, n_trees, class_ptr, class_trees, n_class,
        base_score, margins, pred, warps);
    return check_launch("rmp_forest_predict");
}
