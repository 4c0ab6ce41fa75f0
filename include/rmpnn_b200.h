/*
 * rmpnn_b200.h -- C ABI of the B200-native RNA-MPNN structure-to-sequence forward path.
 *
 * The reference (givemeone1astkiss/RNA-MPNN) has no native layer: its hot path is a
 * sequence of ATen calls inside Python nn.Modules.  Each entry point below replaces
 * one such group of calls; the reference location is cited as file:line under the
 * reference repository root.  INTEGRATION.md shows the ctypes binding a maintainer
 * adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates all
 *     inputs, outputs and workspaces); the library never allocates or frees device
 *     memory and keeps no pointer after return;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), no hidden
 *     synchronisation, so calls can be captured in a CUDA graph;
 *   - return value: 0 = OK, negative = argument error (RMP_E_*), positive =
 *     cudaError_t of the failed launch; rmp_last_error_string() describes it;
 *   - fp32 tensors are row-major contiguous; indices are int32 inside the library
 *     and int64 only where the reference API exposes them (E_idx);
 *   - B structures, N padded length, K = min(top_k, N) neighbour slots, H = 128.
 *
 * Layout in HBM
 *   nodes are the valid residues in (b, i) order ("valid-only flattening",
 *   rdesign/model/feature.py:176-177,235-237): Nv = sum_b L_b rows.
 *   edges are (centre node, slot k) in row-major order; every valid node of
 *   structure b has deg_b = min(K, L_b) edges, so a node's edges are contiguous
 *   (CSR row_ptr) and Ne = sum_b L_b * deg_b.
 */
#ifndef RMPNN_B200_H
#define RMPNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMP_OK            0
#define RMP_E_ARG        -1   /* null pointer / non-positive size */
#define RMP_E_SHAPE      -2   /* unsupported dimension (e.g. H != 128, K > 128) */
#define RMP_E_ALIGN      -3   /* pointer not 16-byte aligned */
#define RMP_E_WORKSPACE  -4   /* workspace too small */
#define RMP_E_ARCH       -5   /* device is not sm_100 */

#define RMP_HIDDEN      128   /* hidden width the kernels are specialised for */
#define RMP_NODE_RAW    101   /* angle 12 | rbf 80 | direction 9   (rdesign/config/glob.py:8-13)  */
#define RMP_EDGE_RAW    115   /* quat 4  | rbf 96 | direction 15  (rdesign/config/glob.py:14-18) */
#define RMP_MAX_K       128   /* neighbour slots supported by the warp top-K */

/* precision of the fused layer path */
#define RMP_PREC_FP32     0   /* fp32 CUDA-core kernels: 1e-4 logit parity with the reference's CPU path */
#define RMP_PREC_BF16     1   /* bf16 operands on tcgen05 tensor cores, fp32 accumulate: 2e-2 budget */
#define RMP_PREC_FP32X3   2   /* fp32 parity on tcgen05: split bf16 operands, 3 MMAs per product (rmp_row_mlp_x3) */

const char* rmp_version(void);
const char* rmp_last_error_string(void);
/* compute capability major*10+minor of the current device, or negative error */
int rmp_device_arch(void);
/* sizeof(struct rmp_mlp_desc) as compiled -- lets a foreign-language binding verify its mirror */
int rmp_abi_struct_size(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t rmp_launch_count(void);

/* ------------------------------------------------------------------ graph index
 * Valid-only flattening bookkeeping from the mask alone.
 * Replaces: masked_select / nonzero / cumsum bookkeeping in
 *   rdesign/model/feature.py:173-177, 227-237.
 *   node_map [B*N]  : dense (b,i) -> flat node id, -1 for masked residues
 *   node_src [B*N]  : flat node id -> dense b*N+i          (first Nv entries)
 *   sinfo    [B*4]  : per structure {L_b, deg_b, node_off_b, edge_off_b}
 *   row_ptr  [B*N+1]: CSR edge offsets per flat node        (first Nv+1 entries)
 *   counts   [2]    : {Nv, Ne}
 */
int rmp_graph_index(const float* mask, int B, int N, int K,
                    int32_t* node_map, int32_t* node_src, int32_t* sinfo,
                    int32_t* row_ptr, int32_t* counts, void* stream);

/* ------------------------------------------------------------------ kNN (a-1)
 * Brute-force tiled kNN over one backbone atom per residue.
 * Replaces: RNAFeatures._dist, rdesign/model/feature.py:42-50.
 *   X          [B,N,A,3] fp32, atom `atom` of each residue is used
 *   mask       [B,N] fp32 0/1
 *   E_idx      [B,N,K] int64, ascending adjusted distance, ties -> lower index
 *   D_neighbors[B,N,K] fp32 (may be NULL)
 * Distances follow the reference's rounding sequence exactly:
 *   sqrt(((dx*dx + dy*dy) + dz*dz) + 1e-6f), no FMA contraction; pairs with a
 *   masked end are 20001 (= 10000 + rowmax(10000) + 1).  Self is included.
 */
int rmp_knn(const float* X, const float* mask, int B, int N, int A, int atom, int K,
            int64_t* E_idx, float* D_neighbors, void* stream);

/* ------------------------------------------------------------------ edge list (a-2)
 * Replaces: rdesign/model/feature.py:227-233 (shift + masked_select + cat).
 *   ctr, nbr [Ne] int32 flat node ids;  E_flat [2,Ne] int64 (row 0 centre, row 1 neighbour;
 *   may be NULL); ne_cap = capacity of the outputs in edges.
 */
int rmp_edge_list(const int64_t* E_idx, const int32_t* node_map, const int32_t* node_src,
                  const int32_t* sinfo, const int32_t* row_ptr, const int32_t* counts,
                  int B, int N, int K, int ne_cap,
                  int32_t* ctr, int32_t* nbr, int64_t* E_flat, void* stream);

/* ------------------------------------------------------------------ raw features (a-3, a-4, a-5)
 * Replaces: _dihedrals, _rbf/_get_rbf, _orientations_coarse, _quaternions,
 *   rdesign/model/feature.py:52-154, 179-220.
 *   frames   [B*N,9]  workspace: per-residue local frame (rows b1, n0, b1 x n0)
 *   rbf_mu   [16]     fp32 RBF centres (torch.linspace(0,20,16) bits)
 *   node_raw [Nv,101] angle12 | rbf80 | dir9
 *   edge_raw [Ne,115] quat4 | rbf96 | dir15
 */
int rmp_rdesign_raw_features(const float* X, int B, int N,
                             const int32_t* node_src, const int32_t* ctr, const int32_t* nbr,
                             const int32_t* counts, int nv_cap, int ne_cap,
                             const float* rbf_mu, float* frames,
                             float* node_raw, float* edge_raw, void* stream);
/* The same with edge rows 128 floats apart (columns 115..127 zero): the K = 128 input of the tensor-core row MLPs */
int rmp_rdesign_raw_features_k128(const float* X, int B, int N,
                                  const int32_t* node_src, const int32_t* ctr, const int32_t* nbr,
                                  const int32_t* counts, int nv_cap, int ne_cap,
                                  const float* rbf_mu, float* frames,
                                  float* node_raw, float* edge_raw128, void* stream);

/* ------------------------------------------------------------------ fused row MLP (a-6, a-8, a-9)
 * One launch = up to RMP_MAX_STAGES chained Linear(+exact GELU) stages over row
 * tiles, input optionally gathered/concatenated, with a fused epilogue.
 * Weights are passed TRANSPOSED and zero padded: Wt[k_pad][n_out], k_pad =
 * round_up(k_in, 32), n_out % 128 == 0 (see rmp_pack_rows()).
 */
#define RMP_MAX_STAGES 6
#define RMP_MAX_SEGS   3

#define RMP_EPI_STORE      0  /* out[r,:] = y                                                          */
#define RMP_EPI_NORMALIZE  1  /* reference Normalize: gain*(y-mu)/(sqrt(var_unbiased+1e-6)+1e-6)+bias
                                 (rdesign/model/functional.py:79-96)                                   */
#define RMP_EPI_RESID_LN   2  /* out = LayerNorm(resid + y), eps 1e-5 (rdesign/model/mpnn.py:37)        */
#define RMP_EPI_SEGSUM     3  /* out[seg[r],:] += y * alpha   (scatter_sum / scale, mpnn.py:34)         */
#define RMP_EPI_RESID_ADD  4  /* out[r,:] = resid[r,:] + y  (in-place edge update, rnampnn/model/mpnn.py:263;
                                 out may alias resid)                                                   */

typedef struct rmp_mlp_desc {
    /* input = concat over segments of src[idx ? idx[r] : r, 0:width] */
    const float*   seg_src[RMP_MAX_SEGS];
    const int32_t* seg_idx[RMP_MAX_SEGS];
    int32_t        seg_width[RMP_MAX_SEGS];
    int32_t        seg_bf16[RMP_MAX_SEGS];  /* tensor-core arm only: segment rows are bf16 (a side copy written by
                                               out_bf16 / rmp_graph_norm), gathered with 256-bit loads, no conversion */
    int32_t        num_segs;
    /* stages */
    const float*   Wt[RMP_MAX_STAGES];      /* [k_pad][n_out] */
    const float*   bias[RMP_MAX_STAGES];    /* [n_out] */
    int32_t        k_in[RMP_MAX_STAGES];
    int32_t        n_out[RMP_MAX_STAGES];
    int32_t        gelu[RMP_MAX_STAGES];
    int32_t        num_stages;
    /* epilogue */
    int32_t        epilogue;
    const float*   gain;                    /* NORMALIZE / RESID_LN weight [n] */
    const float*   beta;                    /* NORMALIZE / RESID_LN bias   [n] */
    const float*   resid;                   /* RESID_LN: [rows, n]               */
    const int32_t* seg_id;                  /* SEGSUM: destination row per input row (non-decreasing) */
    float          alpha;                   /* SEGSUM scale (1/30)               */
    const float*   row_scale;               /* optional [rows]: STORE / SEGSUM multiply row r's result by it
                                               (row masks, 1/valid-neighbour-count; rnampnn/model/mpnn.py:192,219) */
    float*         out;
    void*          out_bf16;                /* tensor-core arm, STORE / RESID_ADD: optional bf16 copy of the result rows */
    /* rows: device count (rows_dev[rows_slot]) bounded by rows_cap */
    const int32_t* rows_dev;
    int32_t        rows_slot;
    int32_t        rows_cap;
} rmp_mlp_desc;

int rmp_row_mlp(const rmp_mlp_desc* desc, void* stream);

/* out[r,:] = norm(a[r,:] + b[r,:]) over 128 channels; b may be NULL.
 *   variant 0: nn.LayerNorm, biased variance (norm1, rdesign/model/mpnn.py:35)
 *   variant 1: reference Normalize, unbiased variance, gain*(x-mu)/(sqrt(var+eps)+eps)+bias
 *              (rdesign/model/functional.py:79-96)
 * If zero_b != 0 the rows of b are zeroed after being read (re-arms the segmented-sum
 * buffer for the next layer). */
int rmp_add_layernorm(const float* a, float* b, const float* gamma, const float* beta, float eps,
                      const int32_t* rows_dev, int rows_slot, int rows_cap, int zero_b, int variant,
                      float* out, void* stream);

/* ------------------------------------------------------------------ head (a-9)
 * Final Linear(128 or hidden -> 4) + softmax + argmax / temperature sampling.
 * Replaces: Readout's last Linear (rdesign/model/functional.py:121-127) and
 *   softmax/argmax (rdesign/model/rdesign.py:112-115).
 *   W [4,k_in] (nn.Linear layout), b [4]
 *   temperature <= 0 -> samples = argmax(probs); > 0 -> one categorical draw per node from
 *   softmax(logits / T) with a counter-based generator keyed by (seed, node id).
 *   logits/probs [rows,4] fp32, samples [rows] int64 (each may be NULL)
 */
int rmp_readout_head(const float* h, int k_in, const float* W, const float* b,
                     float temperature, uint64_t seed,
                     const int32_t* rows_dev, int rows_slot, int rows_cap,
                     float* logits, float* probs, int64_t* samples, void* stream);

/* ------------------------------------------------------------------ device-side collate (section 8f)
 * Replaces the host collate loops rdesign/utils/data.py:84-119 and rnampnn/utils/data.py:110-142.
 *   packed    [sum L, A_in, 3] fp32 or fp64 (packed_is_f64): the residues of all structures back to back
 *   row_start [B] int64: first packed row of each structure of this batch; lengths [B] int32
 *   X [B,N,A_out,3] fp32: the first A_out atoms, NaN -> 0, zero padded;  mask [B,N] fp32 prefix mask.  N >= max length. */
int rmp_collate(const void* packed, int packed_is_f64, const int64_t* row_start, const int32_t* lengths,
                int B, int N, int A_in, int A_out, float* X, float* mask, void* stream);

/* ------------------------------------------------------------------ tree-ensemble readout head (section 8f rank 4)
 * Replaces `self.xgb_readout.predict(h_V.to('cpu').detach().numpy())` (rdesign/model/rdesign.py:153,
 * rnampnn/model/rnampnn.py:298): xgboost gbtree inference over the node embeddings, multi:softmax -> class.
 *   X [rows, ld] fp32 (first n_feat columns used; NaN = missing)
 *   nodes [n_nodes] 16-byte records {fp32 threshold (leaf: value), feature | default_left << 31, left, right}
 *         (children are global node ids, left < 0 marks a leaf);  tree_root [n_trees]
 *   class_ptr [n_class+1], class_trees [n_trees]: the trees of every output group in tree order (CSR)
 *   margins [rows, n_class] fp32 (optional) = base_score + the group's leaves summed in tree order in fp32;
 *   pred [rows] int64 (optional) = first maximal margin. */
int rmp_forest_predict(const float* X, int rows, int ld, int n_feat, const void* nodes, int n_nodes,
                       const int32_t* tree_root, int n_trees, const int32_t* class_ptr, const int32_t* class_trees,
                       int n_class, float base_score, float* margins, int64_t* pred, void* stream);

/* ================================================================== secondary path: rnampnn (a-10 .. a-16)
 * The reference keeps the padded layout end to end: rows are dense residues b*N+i, edges are (row, slot k),
 * edge_index [B,N,K] int64 holds the neighbour's index INSIDE its structure or -1.  The Linear stacks of this
 * path (raw_project + RNABert FFN, edge embedding, ResMPNN message / edge MLPs, RawFFN, Readout) run on
 * rmp_row_mlp; the entry points below are what is not a Linear.
 */

/* a-10.  Replaces ResFeature._get_res_graph, rnampnn/model/feature.py:205-256.
 *   coords [B,N,A,3] fp32, mask [B,N]; centroid = mean of the A atoms (summed in torch-CPU's order 0,4,5,6,1,2,3
 *   for A = 7), D = sqrt(((dx^2+dy^2)+dz^2)+1e-6) (IEEE), pairs with a padded end and the diagonal = 1e6;
 *   ascending (D, index); then columns past N, "self", slots k > L_b-1 and padded rows -> -1.
 *   edge_index [B,N,K] int64 (K = num_neighbours, may exceed N); D_neighbors [B,N,K] fp32 or NULL. */
int rmp_res_graph(const float* coords, const float* mask, int B, int N, int A, int K,
                  int64_t* edge_index, float* D_neighbors, void* stream);

/* Per-edge bookkeeping derived from any edge_index (the module API accepts a caller-made one):
 *   ctr_row/nbr_row [B*N*K] int32 dense rows of the centre and of the neighbour (-1 gathers row b*N+0 like the
 *   reference's safe_edge_index, rnampnn/model/mpnn.py:173-174); edge_keep = (edge_index != -1) * mask_i
 *   (feature.py:564-569); edge_w = (edge_index != -1) / max(#valid slots, 1) (mpnn.py:191, :215-219). */
int rmp_res_edge_rows(const int64_t* edge_index, const float* mask, int B, int N, int K, int32_t* ctr_row,
                      int32_t* nbr_row, float* edge_keep, float* edge_w, void* stream);

/* a-11, a-12.  Replaces _inside_dists/_inside_angles/_inside_dihedrals and _cross_dists/_cross_angles/
 * _cross_dihedrals (+ _gather_neighbours), rnampnn/model/feature.py:258-517.
 *   raw      [B*N, nd(nd-1)/2 + na-2 + nh-3]          (padded rows: distances 1e6, cosines 0); NULL skips it
 *   edge_raw [B*N*K, nd^2 + (na-1)^2 + (nh-2)^2]      (invalid edges: distances 1e6, cosines 0); NULL skips it */
int rmp_res_raw_features(const float* coords, const float* mask, const int64_t* edge_index, int B, int N,
                         int A, int K, int n_inside_dist, int n_inside_angle, int n_inside_dihedral,
                         int n_cross_dist, int n_cross_angle, int n_cross_dihedral, float* raw,
                         float* edge_raw, void* stream);

/* a-13.  Replaces GraphNormalization.forward, rnampnn/model/functional.py:18-48, over 128 channels.
 *   out = ((x - mean_b)/sqrt(var_b + 1e-6) * scale + shift) * mask,  x = a (+ b when b != NULL)
 *   mean/var per structure and channel over the valid rows; the reference's variance also counts (0 - mean)^2
 *   for each of the n_total - L_b padded rows it holds (n_total = N, or 4500 inside RNABert).
 *   zero_b != 0 re-zeroes b after it is read (b doubles as the segmented-sum buffer of the message kernel).
 *   stats_ws: caller workspace of B*257 doubles. */
int rmp_graph_norm(const float* a, float* b, const float* mask, int B, int N, int n_total,
                   const float* scale, const float* shift, int zero_b, double* stats_ws, float* out,
                   void* out_bf16 /* optional bf16 copy of out for the tensor-core gathers, or NULL */, void* stream);

/* a-15.  Replaces nn.MultiheadAttention(query=key=value=x, key_padding_mask=~mask) inside RNABert.forward,
 * rnampnn/model/functional.py:161-168, between in_proj and out_proj (both run on rmp_row_mlp).
 *   qkv [B*N, 3E] = x in_proj_weight^T + in_proj_bias (q | k | v), E % heads == 0, head dim in {8,16,32}
 *   out [B*N, E] = softmax(q k^T / sqrt(hd)) v per head over the VALID keys of the structure; padded query rows 0. */
int rmp_res_attention(const float* qkv, const float* mask, int B, int N, int E, int heads, float* out, void* stream);
/* The same on the tcgen05 tensor cores (bf16 arm): qkv is the in-projection's bf16 side copy [B*N, 3E], E = 128, head dim 16
 * or 32, an even number of heads.  S = Q K^T and P V per head as UMMA tiles over the valid key tiles, softmax between TMEM
 * and registers (P goes back to tensor memory as the A operand of P V); out fp32 [B*N, E], padded query rows -> 0. */
int rmp_res_attention_tc(const void* qkv_bf16, const float* mask, int B, int N, int E, int heads, float* out, void* stream);


/* x[r,:] *= w[r]  (row masks) */
int rmp_scale_rows(float* x, const float* w, int64_t rows, int width, void* stream);

/* a-16.  rmp_readout_head with logits *= row_mask before softmax / argmax (rnampnn/model/functional.py:86-88). */
int rmp_readout_head_masked(const float* h, int k_in, const float* W, const float* b, const float* row_mask,
                            float temperature, uint64_t seed, int rows, float* logits, float* probs,
                            int64_t* samples, void* stream);

/* Tensor-core arm of rmp_row_mlp (tcgen05 / TMEM / TMA): same descriptor, with the weights passed separately as
 * bf16 in nn.Linear layout, W_bf16[s] = [n_out[s]][k_in[s]] (desc->Wt is ignored).  Operands are rounded to bf16,
 * accumulation is fp32, GELU is the tanh form; 2e-2 logit budget like RMP_PREC_BF16.
 * Restrictions: every segment width a multiple of 128; k_in % 64 == 0, n_out % 128 == 0, both <= 512; bias required;
 * epilogues RMP_EPI_STORE / RMP_EPI_SEGSUM / RMP_EPI_RESID_ADD; rows given on the host (rows_dev == NULL).
 * Replaces the Linear/GELU stacks of rnampnn/model/mpnn.py:131-152 (message / edge MLPs) and
 * rnampnn/model/functional.py:62-74, :105-127, :179-187 (Readout hidden layer, RNABert FFN + projections, RawFFN). */
int rmp_row_mlp_bf16(const rmp_mlp_desc* desc, const void* const* W_bf16, void* stream);

/* Same kernel with a choice of 16-bit operand format.  RMP_OPERAND_FP16 (weights passed as IEEE fp16) keeps 11
 * significant bits instead of 8; conversions saturate to +-65504, so finite inputs never produce inf / NaN.  The
 * rnampnn engine uses it for the two stacks that consume raw geometry and feed a GraphNorm directly (embedding FFN,
 * RawFFN), where bf16 rounding is amplified beyond the 2e-2 budget on badly scaled inputs. */
#define RMP_OPERAND_BF16 0
#define RMP_OPERAND_FP16 1
int rmp_row_mlp_tc(const rmp_mlp_desc* desc, const void* const* W_16, int operand_format, void* stream);
/* fp32-parity arm of rmp_row_mlp on tcgen05: fp32 inputs / weights / activations carried as two bf16 terms
 * (x = x_hi + x_lo, |error| <= 2^-18 |x|), three MMAs per product (hi*hi + lo*hi + hi*lo), fp32 accumulate in TMEM,
 * bias / exact erf GELU / epilogues in fp32: keeps RMP_PREC_FP32's 1e-4 logit parity (measured 1.5e-5 at the C2 shapes).
 * W_split[s] = bf16 [2][n_out[s]][k_in[s]] (high terms, then low terms of the nn.Linear weight; desc->Wt is ignored).
 * add0 / add1 (optional, n_out[0] = 128): fp32 rows of 128 values, add_stride floats apart, gathered by add*_idx (or row r when NULL), added to stage 0's
 * pre-activation -- the per-node projections of W_0's centre / neighbour column blocks, which replace the
 * [h_V[ctr] | h_V[nbr]] input columns of the message MLP (rdesign/model/mpnn.py:31-33).
 * Restrictions: fp32 segments, widths multiples of 128; k_in % 64 == 0 up to 384; n_out 128 or 256; bias required;
 * every epilogue of rmp_row_mlp (NORMALIZE / RESID_LN / SEGSUM stage whole 128-wide rows in shared memory: coalesced
 * rows out, one red per segment and column); rows given on the host. */
int rmp_row_mlp_x3(const rmp_mlp_desc* desc, const void* const* W_split, const float* add0, const int32_t* add0_idx,
                   const float* add1, const int32_t* add1_idx, int add_stride, void* stream);
/* tuning aid: int64 [500] device buffer receiving (tag, clock64) pairs of CTA 0's first epilogue thread of every
 * following rmp_row_mlp_tc launch (scripts/rowmlp_tc_timeline.py); NULL disables it */
void rmp_set_row_tc_timeline(void* buf);
void rmp_set_row_x3_timeline(void* buf);    /* the same for rmp_row_mlp_x3 */

/* ------------------------------------------------------------------ bf16 tensor-core layer path
 * (tcgen05 / TMEM / TMA).  Declared here, implemented in layer_tc.cu.
 */

/* Fused edge-message kernel: for every edge e (rows of h_E, centre ctr[e], neighbour nbr[e])
 *   m = GELU(W_S-1 .. GELU(W_1 GELU(W_0e h_E[e] + Pc[ctr[e]] + Pn[nbr[e]]) + b_1) .. + b_S-1)
 *   agg[ctr[e], :] += alpha * m
 * Replaces rdesign/model/rdesign.py:85 (gather + cat) and rdesign/model/mpnn.py:33-34 (message MLP,
 * scatter_sum / scale) with W_0 = [W_0e | W_0c | W_0n] split by column block and
 * Pc = h_V W_0c^T + b_0, Pn = h_V W_0n^T computed per node by the caller (rmp_row_mlp).
 *   h_E_bf16 [n_edges,128] bf16; W_bf16 [num_stages*128,128] bf16 = rows of W_0e, W_1, .. (nn.Linear layout)
 *   bias[s] fp32 [128] for s >= 1 (bias[0] unused: folded into Pc); Pc [Nv,128] fp32, Pn [Nv,128] bf16
 *   agg [Nv,128] fp32 must be zero on entry (edges of a node may span two tiles);
 *   debug: NULL, or [num_stages][n_edges][128] fp32 receiving every stage's activations (tests).
 * Edges of one centre must be contiguous (they are in the valid-only flattening): row_ptr [Nv+1] are the
 * CSR offsets of the centres (ctr[e] == v  <=>  row_ptr[v] <= e < row_ptr[v+1]). */
int rmp_edge_mlp_bf16(const void* h_E_bf16, const void* W_bf16, int num_stages, const float* const* bias,
                      const float* Pc, const void* Pn_bf16, const int32_t* ctr, const int32_t* nbr,
                      const int32_t* row_ptr, int n_edges, float alpha, float* agg, float* debug, void* stream);

/* Same, with an optional device buffer `timeline` (int64 [8000]) that receives (tag, clock64) pairs of
 * CTA 0's first epilogue thread -- a tuning aid used by scripts/edge_timeline.py; NULL disables it. */
int rmp_edge_mlp_bf16_traced(const void* h_E_bf16, const void* W_bf16, int num_stages, const float* const* bias,
                             const float* Pc, const void* Pn_bf16, const int32_t* ctr, const int32_t* nbr,
                             const int32_t* row_ptr, int n_edges, float alpha, float* agg, float* debug,
                             long long* timeline, void* stream);
/* Training forward of the same kernel (rdesign/model/mpnn.py:15-19,33-34 with Dropout after every GELU): the three message
 * stages and the neighbour sums in ONE pass over h_E, with what the backward needs written out on the way:
 *   z_save[s]   bf16 [n_edges,128]  pre-activation of stage s (s < num_stages)
 *   a_save[s]   bf16 [n_edges,128]  activation of stage s < num_stages-1 in OPERAND FORM 2 GELU(z) mask: the true
 *                                    Dropout(GELU(z)) is a_save * keep_scale / 2, keep_scale = 256 / (256 - round(256 p))
 *   agg         += alpha * sum over the node's edges of Dropout(GELU(z_last))
 * The keep mask of stage s is that of rmp_act_fwd_bf16 / rmp_act_bwd_bf16 for (seed [+ *seed_dev], stream_id0 + s). */
int rmp_edge_mlp_train_fwd_bf16(const void* h_E_bf16, const void* W_bf16, int num_stages, const float* const* bias, const float* Pc,
                                const void* Pn_bf16, const int32_t* ctr, const int32_t* nbr, const int32_t* row_ptr, int n_edges,
                                float alpha, float* agg, void* const* z_save, void* const* a_save, float p_drop, uint64_t seed,
                                const uint64_t* seed_dev, uint32_t stream_id0, void* stream);

/* ------------------------------------------------------------------ training path (section 8f rank 1, configs[3])
 * Per-edge backward of MPNNLayer's message MLP (rdesign/model/mpnn.py:15-19, :33-34), decomposed into HBM-bound passes
 * over bf16 [rows,128] tensors and tcgen05 GEMMs; the Linear stages themselves run on rmp_row_mlp_tc (forward with the
 * weight, backward-data with the transposed weight).  Dropout masks are never stored: both directions regenerate them
 * from a counter-based hash of (seed + *seed_dev, element group, stream_id); seed_dev / step_dev are optional device
 * words so that a CUDA-graph replay of the step sees a new seed / step number without new kernel arguments.
 *
 * a = Dropout_p(GELU_erf(z)) (nn.GELU + nn.Dropout, mpnn.py:17-18); z, a bf16 [rows,width], width % 8 == 0 */
int rmp_act_fwd_bf16(const void* z_bf16, void* a_bf16, int64_t rows, int width, float p_drop, uint64_t seed,
                     const uint64_t* seed_dev, uint32_t stream_id, void* stream);
/* dz = da * mask/(1-p) * GELU'(z) for 128-wide rows; da is either a bf16 [rows,128] tensor or, for the last message
 * stage, gathered on the fly as dagg[ctr[r],:] * alpha (the backward of scatter_sum / scale, mpnn.py:34).
 * db [128] fp32 (may be NULL) accumulates the column sums of dz = the bias gradient. */
int rmp_act_bwd_bf16(const void* z_bf16, const void* da_bf16, const float* dagg, const int32_t* ctr, float alpha,
                     void* dz_bf16, float* db, int rows, float p_drop, uint64_t seed, const uint64_t* seed_dev,
                     uint32_t stream_id, void* stream);
/* out[idx[r],:] += x[r,:]   x bf16 [rows,128], out fp32: backward of the h_V[E_idx] gathers (rdesign.py:85) */
int rmp_scatter_add_rows(const void* x_bf16, const int32_t* idx, float* out, int rows, void* stream);
/* out[v,:] = sum of rows row_ptr[v] .. row_ptr[v+1]-1 of x (bf16 [rows,128]) in row order, fp32 [n_seg,128], OVERWRITTEN: the
 * scatter_sum by centre node (functional.scatter_sum with src_idx = E_idx[0], rdesign/model/mpnn.py:34) and its transposes
 * in the backward, without atomics (the edges of a node are contiguous). */
int rmp_segment_sum_rows(const void* x_bf16, const int32_t* row_ptr, float* out, int n_seg, void* stream);
/* The same with an indirection: out[v,:] = sum over r in [row_ptr[v], row_ptr[v+1]) of x[perm[r],:].  With perm = the edges
 * sorted by NEIGHBOUR node and row_ptr the matching offsets (the transposed CSR of the kNN graph) this is the scatter of
 * d Pn = sum of dZ0 rows by E_idx[1] (backward of the h_V[E_idx] gather, rdesign/model/rdesign.py:85), atomics-free. */
int rmp_segment_sum_rows_perm(const void* x_bf16, const int32_t* perm, const int32_t* row_ptr, float* out, int n_seg, void* stream);
/* dW[n,k] += sum_r dZ[r,n] X[r,k]   dZ, X bf16 [rows,128], dW fp32 [128,128]: weight gradient of a 128->128 Linear
 * (tcgen05, both operands read MN-major from their row-major tiles, per-CTA TMEM accumulator, one red.global pass) */
int rmp_gemm_dw_bf16(const void* dZ_bf16, const void* X_bf16, float* dW, int rows, void* stream);
/* One 128 -> 128 Linear stage over bf16 rows as a lean HBM-bound tcgen05 GEMM with fused epilogues:
 *   Z = A W^T (+ bias) (+ Pc[ctr[r]] + Pn[nbr[r]]);   A bf16 [rows,128], W bf16 [128,128] (nn.Linear layout)
 *   z_bf16 <- Z;  a_bf16 <- Dropout_p(GELU(Z)) (same mask as rmp_act_fwd/bwd_bf16 for (seed, stream_id));
 *   acc32 += Z (fp32, gradient accumulation).  Any subset of the three outputs; Pc/Pn bf16 [Nv,128] optional.
 * The dropout probability is quantised to round(256 p)/256 (one 64-bit hash per 8 elements, 8-bit lanes). */
int rmp_gemm_rows_bf16(const void* A_bf16, const void* W_bf16, const float* bias, const void* Pc_bf16, const void* Pn_bf16,
                       const int32_t* ctr, const int32_t* nbr, void* z_bf16, void* a_bf16, float* acc32, int rows,
                       int w_ld /* elements per row of the matrix W is a 128x128 window of; 0 = 128 */,
                       int w_transposed /* Z = A W instead of A W^T: backward-data without a transposed copy */,
                       float p_drop, uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id, void* stream);
/* nn.LayerNorm over 128 channels with the residual add in front of it (rdesign/model/mpnn.py:35 norm1(h_V + dh), :37
 * norm2(h_V + dh)), training path:
 *   forward   x = a + b (b fp32 or bf16, may be NULL: x = a, x_out untouched);  y = gamma (x - mean) rstd + beta;  mean, rstd [rows]
 *   backward  dx (fp32, and optionally a bf16 copy) from dy;  dgamma[128] += sum dy xhat;  dbeta[128] += sum dy */
int rmp_add_layernorm_fwd(const float* a, const void* b, int b_is_bf16, const float* gamma, const float* beta, float eps, int rows,
                          float* x_out, float* y, float* mean, float* rstd, void* stream);
int rmp_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int rows, float* dx,
                      void* dx_bf16, float* dgamma, float* dbeta, void* stream);
/* Column sums over the rows of a [rows,width] matrix, ADDED to the outputs (bias / LayerNorm parameter gradients of the per-node
 * modules, rdesign/model/mpnn.py:11-12,20-28):  s0[c] += sum_r a[r][c];  with x (fp32), mean, rstd [rows]:
 * s1[c] += sum_r a[r][c] (x[r][c] - mean[r]) rstd[r].   a fp32 or bf16; width = 4 * (a divisor of 256). */
int rmp_colsum(const void* a, int a_is_bf16, const float* x, const float* mean, const float* rstd, int rows, int width, float* s0,
               float* s1, void* stream);
/* out32 = sum_l A_l W_l^T (or A_l W_l with w_transposed), l < k_blocks:   A bf16 [k_blocks][rows][128], W bf16 [k_blocks][128][128],
 * out32 fp32 [rows,128], OVERWRITTEN.  One pass over the K-concatenated operands (W_l streamed with A_l through the TMA ring):
 * the backward of the edge embedding needs d h_E = sum over the layers of dZ0_l W1e_l (rdesign/model/rdesign.py:85 feeds the
 * same h_E to every layer); accumulating it layer by layer re-reads and re-writes the fp32 [Ne,128] buffer nine times. */
int rmp_gemm_rows_kcat_bf16(const void* A_bf16, const void* W_bf16, int k_blocks, float* out32, int rows, int w_transposed,
                            void* stream);
/* Reference Normalize (rdesign/model/functional.py:79-96) over 128 channels for the training path:
 *   forward   y_bf16 = gain * (x - mu) / (sqrt(var_unbiased + eps) + eps) + bias            x fp32 [rows,128]
 *   backward  dx_bf16 from dy fp32; dgain[128] += sum dy * xhat, dbias[128] += sum dy, dbsum[128] += sum dx */
int rmp_normalize_fwd_bf16(const float* x, const float* gain, const float* bias, void* y_bf16, int rows, float eps, void* stream);
int rmp_normalize_bwd_bf16(const float* x, const float* dy, const float* gain, void* dx_bf16, float* dgain, float* dbias,
                           float* dbsum, int rows, float eps, void* stream);
/* torch.optim.Adam step (no weight decay, no amsgrad) over flat fp32 buffers of n elements (n % 4 == 0);
 * grad is multiplied by grad_scale first (1 / world size after the all-reduce; rdesign.py:90-93) */
int rmp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float grad_scale,
                  float lr, float beta1, float beta2, float eps, int step, const int* step_dev, void* stream);

/* Fused node-update kernel on 128-node tiles (tcgen05, weights streamed by TMA):
 *   x1 = prologue_ln ? LayerNorm1(h_in + agg) : h_in      (agg rows are re-zeroed after being read)
 *   then num_stages chained Linear stages, stage s: K[s] -> N[s] (N in {128,256}), W_bf16[s] = nn.Linear
 *   weight [N][K] as bf16, bias[s] fp32, kind[s]:
 *     0  bias + GELU (tanh form)                          -> operand of the next stage
 *     1  bias + residual x1 + LayerNorm2 (eps 1e-5)       -> h_out [n_rows,128] (and next operand)
 *     2  bias (N = 256, last stage only)                  -> p_out [n_rows,128] fp32 = columns 0..127 (Pc),
 *                                                            pn_out [n_rows,128] bf16 = columns 128..255 (Pn)
 * Replaces rdesign/model/mpnn.py:35-37 (norm1, dense FFN, norm2); a trailing kind-2 stage emits the
 * next layer's per-node projections [Pc | Pn] consumed by rmp_edge_mlp_bf16.
 * x1_scratch [n_rows,128] fp32 is a caller-provided workspace for the residual. */
int rmp_node_mlp_bf16(const float* h_in, float* agg, int prologue_ln, const float* ln1_g, const float* ln1_b,
                      int num_stages, const void* const* W_bf16, const int32_t* K, const int32_t* N,
                      const int32_t* kind, const float* const* bias, const float* ln2_g, const float* ln2_b,
                      float* x1_scratch, float* h_out, float* p_out, void* pn_out_bf16, int n_rows, void* stream);

/* Fused edge featurizer of the bf16 arm: raw edge features (a-4, a-5) -> Linear(115->128) on tcgen05 ->
 * reference Normalize -> h_E as bf16.  Replaces rdesign/model/feature.py:193-206, :215-220, :224 (edge half).
 *   X [B*N,6,3] fp32, frames [B*N,9] (from rmp_rdesign_raw_features), node_src/ctr/nbr as in rmp_edge_list
 *   W_perm_bf16 [128,128] bf16: edge_embedding.weight with its input columns permuted to the kernel's
 *   feature order [rbf 96 | quaternion 4 | direction 15 | 13 zero columns]; bias/gain/beta fp32 [128]
 *   h_E_bf16 [n_edges,128] bf16 */
int rmp_edge_embed_bf16(const float* X, const float* frames, const int32_t* node_src, const int32_t* ctr,
                        const int32_t* nbr, int n_edges, const float* rbf_mu, const void* W_perm_bf16,
                        const float* bias, const float* gain, const float* beta, void* h_E_bf16, void* stream);

/* tuning aid: int64 [800] device buffer that receives a clock64 trace of the node kernel while the environment
 * variable RMP_NODE_TIMELINE is set (scripts/node_timeline.py); NULL disables it */
void rmp_set_node_timeline(void* buf);
/* tuning aid (scripts/edge_variants.py): epilogue variant of rmp_edge_mlp_bf16 (0 = GELU, 1 = activations held as 2 GELU,
 * bit-identical results; shipped default 1) and producer / issuer flags (1 = first edge tile requested right behind W1e,
 * 2 = run-ahead stage-0 GEMMs issued only behind a GEMM the epilogue waits for; shipped default 3).  The environment variable
 * RMP_EDGE_TUNING="<variant>,<flags>" sets the same pair before the first launch. */
void rmp_set_edge_tuning(int variant, int flags);

/* fp32 -> bf16 row conversion (h_E storage for the bf16 path), n elements */
int rmp_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RMPNN_B200_H */
