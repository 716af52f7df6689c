/*
 * lcr_b200.h -- C ABI of liblcr_b200.so: the B200 (sm_100a) implementation of the LCR-Net
 * inference hot path.  This is the drop-in boundary: plain pointers and sizes only, no torch
 * types.  Every entry point cites the reference interface it replaces (paths relative to the
 * upstream nubot-nudt/LCR-Net tree).
 *
 * Conventions
 *  - All data pointers are DEVICE pointers unless the name ends in `_host`.
 *  - All calls are asynchronous on `stream` (a cudaStream_t passed as void*).
 *  - Return value: 0 on success, <0 on error (LCR_ERR_*); lcr_last_error() describes it.
 *  - "Stack mode" batches (reference: experiments/lcrnet/data.py:10-74): several clouds are
 *    concatenated along dim 0 and described by int64 `lengths[B]`.  A *stack* is the unit the
 *    reference feeds to one model.forward() (1 cloud for the descriptor path, 2 for a pair);
 *    GroupNorm statistics are per stack.  Kernels that normalise take `stack_off[S+1]` so that
 *    many stacks run in one launch with per-stack statistics.
 *  - Scratch memory is caller-provided: call the matching *_ws_bytes() first.
 *  - Index tables: int64 at the reference-compatible boundary, int32 internally
 *    (`idx_is64` selects).  Padding value of a neighbour table = total number of support rows.
 */
#ifndef LCR_B200_H_
#define LCR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCR_OK 0
#define LCR_ERR_INVALID (-1)   /* bad argument (reference: TORCH_CHECK in common/torch_helper.h:6-35) */
#define LCR_ERR_WORKSPACE (-2) /* workspace too small */
#define LCR_ERR_CUDA (-3)      /* CUDA runtime error */
#define LCR_ERR_OVERFLOW (-4)  /* a capacity bound was exceeded on the device */

const char* lcr_last_error(void);
int lcr_abi_version(void);
/* number of CUDA kernels this library has launched in this process (for benchmark reports) */
int64_t lcr_launch_count(void);
/* Per-kernel-group device timing for benchmark reports: between begin and end every kernel group
 * is bracketed by CUDA events on its launching stream; get(i) returns its name, elapsed ms and
 * the algorithmic flops / bytes it was credited with. */
void lcr_profile_begin(void);
int lcr_profile_end(void);
int lcr_profile_get(int i, char* name, int name_cap, double* ms, double* flops, double* bytes);

/* ------------------------------------------------------------------------------------------
 * a1. Voxel-grid subsampling.
 * Replaces utils.ext.grid_subsampling (utils/extensions/pybind.cpp:14-18;
 * cpu/grid_subsampling/grid_subsampling.cpp:5-62, grid_subsampling_cpu.cpp:3-75) and the legacy
 * twin cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106.
 *
 * points[n_total,3] f32, lengths[batch] i64 (device).  Output: out_points (capacity n_total rows),
 * out_lengths[batch] i64 (device), *out_total (device, i64) = sum(out_lengths).
 * order_mode 1 = reference order (iteration order of the reference's std::unordered_map,
 * reproduced exactly), 0 = first-seen order (faster; same set of centroids, bit-identical values).
 * Centroids are bit-exact in both modes (sequential fp32 sums in input order).
 * *out_status (device i32, optional) is set to LCR_ERR_OVERFLOW if a voxel key exceeds 44 bits.
 * ---------------------------------------------------------------------------------------- */
size_t lcr_grid_subsample_ws_bytes(int64_t n_total, int batch);
int lcr_grid_subsample(const float* points, int64_t n_total, const int64_t* lengths, int batch,
                       float voxel_size, int order_mode, float* out_points, int64_t* out_lengths,
                       int64_t* out_total, int32_t* out_status, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a2. Radius neighbour search (stack mode).
 * Replaces utils.ext.radius_neighbors (pybind.cpp:9-13; cpu/radius_neighbors/radius_neighbors.cpp:5-68,
 * radius_neighbors_cpu.cpp:3-91) + the python column cut of ops/radius_search.py:25-26, and the
 * legacy cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:125-333.
 *
 * For every query, all supports of the same batch element with d2 < radius^2 (fp32, no FMA,
 * strict), ascending d2, exact ties by ascending support index; the first `width` are written
 * to out_idx[nq_total, width] (global stack indices; pad = ns_total).  out_counts[nq_total] (i32,
 * optional) receives the untruncated neighbour count, *out_max_count (device i32) their maximum
 * (the reference table width).  `width` must be <= LCR_RADIUS_MAX_WIDTH.  out_idx == NULL runs the
 * counting pass only.  *out_status (device i32, optional) is set to LCR_ERR_OVERFLOW if a cloud spans
 * more than 16384 grid cells per axis or a query has more than LCR_RADIUS_MAX_WIDTH neighbours.
 * ---------------------------------------------------------------------------------------- */
#define LCR_RADIUS_MAX_WIDTH 8192
size_t lcr_radius_neighbors_ws_bytes(int64_t nq_total, int64_t ns_total, int batch);
int lcr_radius_neighbors(const float* q_points, int64_t nq_total, const float* s_points, int64_t ns_total,
                         const int64_t* q_lengths, const int64_t* s_lengths, int batch, float radius,
                         int width, void* out_idx, int idx_is64, int32_t* out_counts,
                         int32_t* out_max_count, int32_t* out_status, void* ws, size_t ws_bytes, void* stream);
/* As above; reuse_grid is a bit set.  Bit 0 (1): skip the support-grid build and reuse the one a previous call left
 * in `ws` (same ws pointer, supports, support lengths, batch and radius): the pyramid asks three tables per support
 * level.  Bit 1 (2): nearest-only -- width must be 1; writes column 0 of the table (the closest support inside the
 * radius, ties by index, pad = ns_total) without collecting and sorting the other hits: what the decoder's
 * nearest_upsample reads from the up-sampling tables (backbone4.py:333-373, modules/ops/index_select... [:, 0]). */
int lcr_radius_neighbors_ex(const float* q_points, int64_t nq_total, const float* s_points, int64_t ns_total,
                            const int64_t* q_lengths, const int64_t* s_lengths, int batch, float radius, int width,
                            void* out_idx, int idx_is64, int32_t* out_counts, int32_t* out_max_count,
                            int32_t* out_status, void* ws, size_t ws_bytes, int reuse_grid, void* stream);

/* ------------------------------------------------------------------------------------------
 * a4. KPConv forward.  Replaces KPConv.forward (experiments/lcrnet/modules/kpconv/kpconv.py:79-122).
 * s_feats[n_support, c_in], q_points[m_query,3], s_points[n_support,3], idx int32 [m_query, ld_idx]
 * (first H columns used, pad = n_support), kernel_points[15,3], weights[15, c_in, c_out],
 * bias[c_out] or NULL -> out[m_query, c_out].  s_flags[n_support] (u8, optional): 1 where the
 * feature row sum is > 0 (the reference's neighbour_num counts only those rows, :113-116);
 * NULL counts every valid neighbour.  c_in in {1, 32, 64, 128, 256}.  weights_nk (optional): the same
 * weights transposed to [c_out, 15 * c_in]; when given, the contraction runs on the tcgen05 3xTF32 GEMM
 * (weights_nk_lo != NULL: the pair is the pre-split hi / lo halves from lcr_tf32_split).
 * kernel_points_host (optional): HOST copy of kernel_points; when given (and c_in > 1) the gather runs in
 * its sparse form with the kernel points as constant-bank kernel arguments, otherwise the dense form
 * reads them from device memory.
 * ---------------------------------------------------------------------------------------- */
size_t lcr_kpconv_ws_bytes(int64_t m_query, int c_in);
/* workspace with room for the packed (x, y, z, row flag) float4 copy of the support points: selects the fast
 * gather's single 16-byte point load (a workspace of lcr_kpconv_ws_bytes still works, on the 12-byte path) */
size_t lcr_kpconv_ws_bytes2(int64_t m_query, int64_t n_support, int c_in);
/* Tuning knob (A/B measurements, tests): gather variant used when kernel_points_host is given.
 * 0 = exact dense loop (IEEE sqrt / divide influences, also used without host kernel points), 1 = fast dense
 * loop (rsqrt influences, constant-bank kernel points), 2 = sparse influence lists, 3 = auto (default: 6 for
 * c_in = 32, 1 otherwise -- the measured winners on B200), 4 = non-zero-influence mask dispatch, 5 = packed
 * fp32 FMA (FFMA2) loop, 6 = warp-level tensor path (mma.sync m16n8k8, 3xTF32), 7 = fast dense loop that skips
 * all-zero float4 groups of kernel points. */
void lcr_set_gather_mode(int mode);
int lcr_kpconv(const float* s_feats, const uint8_t* s_flags, int64_t n_support, const float* q_points,
               int64_t m_query, const float* s_points, const int32_t* idx, int ld_idx, int H,
               const float* kernel_points, const float* kernel_points_host, float sigma, const float* weights,
               const float* weights_nk, const float* weights_nk_lo, const float* bias, int c_in, int c_out,
               float* out, void* ws, size_t ws_bytes, void* stream);
/* lcr_kpconv with the GroupNorm statistics of its output fused into the GEMM epilogue (the KPConv of a
 * ConvBlock / ResidualBlock is always followed by a GroupNorm: modules/kpconv/modules.py:104-146,186-196).
 * gn_partial: f32 scratch of lcr_gn_blocks_ws_bytes(m_query, c_out) bytes; stack_off i64 [n_stacks+1] row
 * offsets of the stacks among the query rows; every stack must have >= 32 rows.  Needs c_in > 1 and
 * weights_nk.  Follow with lcr_group_norm_finalize_blocks to get the [n_stacks, groups, 2] statistics. */
int lcr_kpconv_gn(const float* s_feats, const uint8_t* s_flags, int64_t n_support, const float* q_points,
                  int64_t m_query, const float* s_points, const int32_t* idx, int ld_idx, int H,
                  const float* kernel_points, const float* kernel_points_host, float sigma, const float* weights,
                  const float* weights_nk, const float* weights_nk_lo, const float* bias, int c_in, int c_out,
                  float* out, void* ws, size_t ws_bytes, float* gn_partial, const int64_t* stack_off, int n_stacks,
                  void* stream);
/* flags[r] = (sum_c x[r, c] > 0) */
int lcr_row_flags(const float* x, int64_t rows, int channels, uint8_t* flags, void* stream);

/* ------------------------------------------------------------------------------------------
 * a5. Linear, GroupNorm, LeakyReLU, residual add, neighbour max-pool.  Replace nn.Linear +
 * GroupNorm wrapper + UnaryBlock / ResidualBlock glue (modules/kpconv/modules.py:33-225) and
 * maxpool (modules/kpconv/functional.py:54-67).
 * lcr_linear: out[n_rows, c_out] = x[n_rows, c_in] . weight_t[c_in, c_out] + bias  (weight_t is the
 *   TRANSPOSE of nn.Linear.weight, prepared once at load time).
 * lcr_group_norm_stats: per (stack, group) mean and 1/sqrt(var+eps) over (channels/groups) x (all
 *   rows of the stack) (biased variance), stats_out f32 [n_stacks, groups, 2].
 * lcr_group_norm_apply: y = act(gn(x) + other) with other = nothing (x2 NULL), a raw tensor
 *   (x2, stats2 NULL) or gn(x2; stats2, gamma2, beta2); act = LeakyReLU(slope) if leaky.
 *   row_flags (optional) receives (sum_c y[r, c] > 0) for the following KPConv.
 * ---------------------------------------------------------------------------------------- */
int lcr_linear(const float* x, int64_t n_rows, int c_in, const float* weight_t, int c_out, const float* bias,
               float* out, void* stream);
size_t lcr_group_norm_ws_bytes(int64_t max_stack_rows, int n_stacks, int groups);
int lcr_group_norm_stats(const float* x, int64_t rows, int channels, int groups, const int64_t* stack_off,
                         int n_stacks, int64_t max_stack_rows, float eps, float* stats_out, void* ws,
                         size_t ws_bytes, void* stream);
/* GroupNorm statistics from the per-32-row-block column partials a fused producer wrote
 * (lcr_linear_tc_gn, lcr_kpconv_gn): same result layout as lcr_group_norm_stats, fixed summation order. */
size_t lcr_gn_blocks_ws_bytes(int64_t rows, int channels);
int lcr_group_norm_finalize_blocks(const float* gn_partial, int64_t rows, int channels, int groups,
                                   const int64_t* stack_off, int n_stacks, float eps, float* stats_out, void* stream);
int lcr_group_norm_apply(const float* x, const float* stats, const float* gamma, const float* beta,
                         const float* x2, const float* stats2, const float* gamma2, const float* beta2,
                         int64_t rows, int channels, int groups, const int64_t* stack_off, int n_stacks,
                         int leaky, float slope, float* y, uint8_t* row_flags, void* stream);
int lcr_maxpool(const float* x, int64_t n_support, const int32_t* idx, int ld_idx, int H, int64_t m_query,
                int channels, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a8. NetVLAD global-descriptor head (eval mode), batched over scans.  Replaces
 * F.normalize + NetVLADLoupe2.forward + GatingContext + F.normalize
 * (modules/netvlad/NetVlad.py:49-87,189-201; model_family/LCRNet_GlobalDescrition.py:34-38).
 * feats[rows, 1024] with scan_off[n_scans+1] row offsets -> out[n_scans, 256] (unit L2 norm).
 * bn1 / bn2 / gating_bn: weight, bias, running_mean, running_var concatenated (4 x C floats).
 * ---------------------------------------------------------------------------------------- */
size_t lcr_netvlad_ws_bytes(int64_t rows, int n_scans);
int lcr_netvlad(const float* feats, int64_t rows, const int64_t* scan_off, int n_scans,
                const float* cluster_weights, const float* cluster_weights2, const float* hidden1_weights,
                const float* bn1, const float* bn2, const float* gating_weights, const float* gating_bn, float* out,
                void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a15. Exact squared-L2 top-k descriptor retrieval.  Replaces the faiss IndexIVFFlat(nlist=1)
 * loop of experiments/loop_detection/eval_loop_detection_overlap_dataset.py:183-214 and
 * experiments/inference/infer_loop_detection_find_top1.py:79-104.
 * queries[n_queries, 256], db[n_db, 256]; valid_counts[n_queries] (i32, optional): query i only
 * searches db rows [0, valid_counts[i]).  out_d2 / out_idx [n_queries, k] ascending (d2, index);
 * missing entries are (+inf, -1).  k <= 64.
 * ---------------------------------------------------------------------------------------- */
int lcr_l2_topk(const float* queries, int64_t n_queries, const float* db, int64_t n_db, int dim, int k,
                const int32_t* valid_counts, float* out_d2, int64_t* out_idx, void* stream);

/* out = act(rowscale[m] * (x . weight_t) + bias): lcr_linear with leading dimensions, an optional
 * per-row scale and act in {0: none, 1: ReLU} (FFN of vanilla_transformer.py:22-28, vote MLP). */
int lcr_linear_ex(const float* x, int64_t n_rows, int c_in, int ld_x, const float* weight_t, int c_out,
                  const float* bias, const float* rowscale, int act, float* out, int ld_out, void* stream);

/* Tensor-core variant of lcr_linear_ex (tcgen05.mma kind::tf32 with 3xTF32 operand splitting, fp32-class
 * accuracy): `weight` is in nn.Linear layout [c_out, c_in] (no transpose).  Requires c_in % 32 == 0,
 * c_out % 4 == 0. */
int lcr_linear_tc(const float* x, int64_t n_rows, int c_in, int ld_x, const float* weight, const float* weight_lo,
                  int c_out, int ld_w, const float* bias, const float* rowscale, int act, float* out, int ld_out,
                  void* stream);
/* lcr_linear_tc (no row scale, no activation, dense output) + the GroupNorm statistics of the output fused
 * into the epilogue (UnaryBlock: Linear -> GroupNorm, modules/kpconv/modules.py:53-83); see lcr_kpconv_gn. */
int lcr_linear_tc_gn(const float* x, int64_t n_rows, int c_in, int ld_x, const float* weight, const float* weight_lo,
                     int c_out, int ld_w, const float* bias, float* out, int ld_out, float* gn_partial,
                     const int64_t* stack_off, int n_stacks, void* stream);
/* Operand split of the 3xTF32 scheme for a weight tensor, done once: hi = rna_tf32(w), lo = rna_tf32(w - hi).
 * Passing (hi, lo) as (weight, weight_lo) to lcr_linear_tc / (weights_nk, weights_nk_lo) to lcr_kpconv lets
 * the GEMM copy both halves of its weight tiles instead of splitting them in every CTA; weight_lo = NULL
 * splits on the fly. */
int lcr_tf32_split(const float* w, int64_t count, float* hi, float* lo, void* stream);

/* ------------------------------------------------------------------------------------------
 * a7. 3D-RoFormer pieces (experiments/lcrnet/modules/thdroformer/*).
 * lcr_layer_norm: y = act(LayerNorm(x + residual) * gamma + beta), residual optional, relu flag
 *   (post-LN residual of rpetransformer.py:139-142 / vanilla_transformer.py:22-28,98-101; LN+ReLU of
 *   modules/vote/vote.py:124-128).
 * lcr_rope: in-place rotary embedding of the first 128 columns (4 heads x 32) of x with per-point
 *   angles theta[rows, 64] (RotaryPositionalEmbedding, rpetransformer.py:41-54).
 * lcr_attention: out = softmax(q k^T / sqrt(32)) v per head, batched over problems given by row
 *   offsets (problem p: queries q_off[p]..q_off[p+1], keys/values k_off[p]..k_off[p+1])
 *   (dynamic_attention with k=None, rpetransformer.py:19-24; MultiHeadAttention,
 *   vanilla_transformer.py:58-72).  q/k/v/out hold heads*32 columns at their leading dimensions.
 * ---------------------------------------------------------------------------------------- */
int lcr_layer_norm(const float* x, const float* residual, const float* gamma, const float* beta, int64_t rows,
                   int channels, float eps, int relu, float* y, void* stream);
int lcr_rope(float* x, int ld, const float* theta, int64_t rows, void* stream);
int lcr_attention(const float* q, int ld_q, const float* k, int ld_k, const float* v, int ld_v,
                  const int64_t* q_off, const int64_t* k_off, int n_problems, int64_t max_q_rows, int heads,
                  int head_dim, float* out, int ld_out, double flops_hint, void* stream);

/* ------------------------------------------------------------------------------------------
 * a9. Vote layer tail, greedy NMS, node centres (modules/vote/vote.py:13-70,166-172;
 * backbone4.py:140-176).
 * lcr_vote_shift: out = points + offsets[:, :3] * min(1, max_range / |offset|).
 * lcr_nms_greedy: per cloud, sequential greedy suppression: point i is kept iff its
 *   nn.PairwiseDistance (eps 1e-6) to ALL previously kept points is > radius; the first point is
 *   always kept.  keep[n] (u8), counts[n_clouds], kept_idx: per cloud, the kept point indices
 *   compacted at the cloud's row offset.
 * lcr_neighbor_mean: out[m] = mean of points[idx[m, h]] over valid entries (idx < n_points).
 * ---------------------------------------------------------------------------------------- */
/* Tensor-core attention: same contract as lcr_attention, computed with tcgen05.mma kind::tf32 (3xTF32
 * operand splitting, fp32-class accuracy), TMEM accumulators and TMA (cp.async.bulk) operand fetch.
 * Requires 16-byte aligned rows. */
int lcr_attention_tc(const float* q, int ld_q, const float* k, int ld_k, const float* v, int ld_v,
                     const int64_t* q_off, const int64_t* k_off, int n_problems, int64_t max_q_rows, int heads,
                     int head_dim, float* out, int ld_out, double flops_hint, void* stream);
/* lcr_attention_tma: the same operator with every q / k / v tile fetched by ONE TMA tensor copy (CUtensorMap with
 * the 128-byte swizzle) -- the default; q_rows / k_rows = row counts of the q and the k / v operands. */
int lcr_attention_tma(const float* q, int ld_q, int64_t q_rows, const float* k, int ld_k, const float* v, int ld_v,
                      int64_t k_rows, const int64_t* q_off, const int64_t* k_off, int n_problems, int64_t max_q_rows,
                      int heads, int head_dim, float* out, int ld_out, double flops_hint, void* stream);

int lcr_vote_shift(const float* points, const float* offsets, int ld_offsets, float max_range, int64_t n, float* out,
                   void* stream);
int lcr_nms_greedy(const float* points, const int64_t* cloud_off, int n_clouds, int64_t max_cloud_rows, float radius,
                   uint8_t* keep, int32_t* counts, int32_t* kept_idx, void* stream);
int lcr_neighbor_mean(const float* points, int64_t n_points, const int32_t* idx, int ld_idx, int H, int64_t m_rows,
                      float* out, void* stream);
/* decoder gather: out[i] = concat(coarse[up_idx[i, 0]] or zeros, fine[i])
 * (nearest_upsample, modules/kpconv/functional.py:6-22; KPDecoder, backbone4.py:355-368) */
int lcr_upsample_concat(const float* coarse, int64_t n_coarse, int c_coarse, const int32_t* up_idx, int ld_up,
                        const float* fine, int c_fine, int64_t n_fine, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a10. Point-to-node partition (modules/ops/pointcloud_partition.py:61-107): nearest node per point
 * (matmul-form squared distance, first minimum), per-node k nearest OWNED points ascending
 * (ties: ascending point index), pad index = n_points, knn_mask 1 where valid.
 * ---------------------------------------------------------------------------------------- */
size_t lcr_point_to_node_ws_bytes(int64_t n_points, int64_t n_nodes);
int lcr_point_to_node(const float* points, int64_t n_points, const float* nodes, int64_t n_nodes, int k,
                      int32_t* point_to_node, uint8_t* node_mask, void* knn_idx, int idx_is64, uint8_t* knn_mask,
                      int32_t* out_status, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a11. Log-domain Sinkhorn with a learnable dustbin (modules/sinkhorn/learnable_sinkhorn.py:13-66).
 * scores[batch, rows, cols], masks u8 (1 = valid, NULL = all valid), alpha device scalar ->
 * out[batch, rows+1, cols+1].
 * a12. Coarse correspondences (modules/geotransformer/superpoint_matching.py:129-160): from the
 * (rows+1) x (cols+1) log scores, row-major list of (i, j, exp score); *out_count on the device;
 * capacity rows + cols.
 * ---------------------------------------------------------------------------------------- */
int lcr_sinkhorn(const float* scores, int batch, int rows, int cols, const uint8_t* row_mask, const uint8_t* col_mask,
                 const float* alpha, int iters, float* out, void* stream);
/* Debug / tuning: out4 = {LOG iterations, LIN iterations, discarded LIN iterations, absorptions} summed over all
 * problems since the last reset (sinkhorn.cu); synchronises the device. */
int lcr_sinkhorn_stats(int64_t* out4, int reset);
/* Node-level Sinkhorn kernel: 0 single CTA per problem (plan in L2), 1 (default) the smallest thread-block cluster
 * whose shared-memory row slabs hold the plan, 4 / 8 force that cluster size.  Point-level (128 x 128) problems always
 * use the register-resident kernel. */
void lcr_set_sinkhorn_cluster(int mode);
/* Point-level kernel: 1 (default) stops the iteration as soon as every scaling reproduces, bit for bit, its value of
 * two iterations before -- the sequence is then periodic and the result after `iters` iterations is known exactly
 * (tests compare both settings bitwise); 0 always runs the full count. */
void lcr_set_sinkhorn_early_exit(int on);
size_t lcr_coarse_matching_ws_bytes(int rows, int cols);
int lcr_coarse_matching(const float* log_scores, int rows, int cols, int32_t* out_i, int32_t* out_j,
                        float* out_scores, int32_t* out_count, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a13. Dense matching and local-to-global registration (model_family/LCRNet.py:218-262;
 * modules/geotransformer/local_global_registration.py:49-246; modules/registration/procrustes.py:6-73).
 * lcr_patch_scores: out[p] = F_a[knn_a[node_a[p]]] . F_b[knn_b[node_b[p]]]^T / sqrt(128), [128 x 128].
 * lcr_fine_correspondences: from the Sinkhorn output [n_pairs, 129, 129]: (i, j) kept iff row- or
 *   column-top-1 beating the dustbin, both points valid; row-major per pair; pair_off = exclusive
 *   scan of the per-pair counts (pair_off[n_pairs] = total); outputs have capacity n_pairs * 256 (only arg-max
 *   entries can be kept: <= 256 per pair); ws = lcr_fine_correspondences_ws_bytes(n_pairs) of scratch.
 * lcr_corr_points: the 3-D points of the correspondences.
 * lcr_local_global_registration: per-pair weighted Procrustes (pairs with >= min_corr
 *   correspondences), hypothesis with most inliers (< radius) over all correspondences, then
 *   `steps` rounds of inlier-reweighted global Procrustes -> out_T[16] (row-major 4x4, src -> ref).
 * ---------------------------------------------------------------------------------------- */
int lcr_patch_scores(const float* feats_a, int64_t n_a, const int32_t* knn_a, const int32_t* node_a,
                     const float* feats_b, int64_t n_b, const int32_t* knn_b, const int32_t* node_b, int n_pairs,
                     int k, int channels, float* out, void* stream);
size_t lcr_fine_correspondences_ws_bytes(int n_pairs);
int lcr_fine_correspondences(const float* log_scores, int n_pairs, const uint8_t* knn_mask_a, const int32_t* node_a,
                             const uint8_t* knn_mask_b, const int32_t* node_b, int32_t* pair_cnt, int32_t* pair_off,
                             int32_t* out_pair, int32_t* out_i, int32_t* out_j, float* out_scores, void* ws,
                             size_t ws_bytes, void* stream);
int lcr_corr_points(const int32_t* c_pair, const int32_t* c_i, const int32_t* c_j, const int32_t* n_corr,
                    int64_t capacity, const float* pts_a, const int32_t* knn_a, const int32_t* node_a,
                    const float* pts_b, const int32_t* knn_b, const int32_t* node_b, float* ref, float* src,
                    void* stream);
size_t lcr_lgr_ws_bytes(int n_pairs, int64_t capacity);
int lcr_local_global_registration(const float* ref, const float* src, const float* scores, const int32_t* pair_off,
                                  int n_pairs, int64_t capacity, float radius, int min_corr, int steps, float* out_T,
                                  void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched forms of a10 / a12 / a13 and the node score product of LCRNet.py:196-199: every pair of a
 * chunk in ONE set of launches.  Clouds are stacked (ref_0, src_0, ref_1, src_1, ...); `pts_off` /
 * `node_off` are int64 device arrays of row offsets per cloud (n_clouds + 1 entries).
 *   lcr_point_to_node_batched  pointcloud_partition.py:61-107 per cloud: owner = GLOBAL node row;
 *       knn_global int32 [M, k] = global point rows (pad = n_points), knn_local int64 [M, k] = rows
 *       local to the node's cloud (pad = points of that cloud: the reference's table; may be NULL)
 *   lcr_node_scores            out[p, i, j] = <f[ref node i], f[src node j]> / sqrt(C) padded with
 *       zeros to [n_pairs, m_max, n_max], plus the padded node masks (rows / columns of the Sinkhorn)
 *   lcr_coarse_matching_batched  superpoint_matching.py:129-160 on [n_pairs, rows+1, cols+1] log
 *       scores; outputs [n_pairs, rows + cols], counts [n_pairs]
 *   lcr_gather_coarse          compacts them into one patch list (global + local node indices,
 *       patch -> pair) given patch_off int32 [n_pairs + 1] (exclusive scan of the counts)
 *   lcr_lgr_batched            local_global_registration.py:140-202 for all scan pairs: patches t own
 *       correspondences [pair_off[t], pair_off[t+1]), scan pair s owns patches [patch_off[s],
 *       patch_off[s+1]); out_T [n_scan_pairs, 4, 4]
 * ---------------------------------------------------------------------------------------- */
int lcr_point_to_node_batched(const float* points, int64_t n_points, const int64_t* pts_off, const float* nodes,
                              int64_t n_nodes, const int64_t* node_off, int n_clouds, int64_t max_cloud_points,
                              int64_t max_cloud_nodes, int k, int32_t* point_to_node, uint8_t* node_mask,
                              int32_t* knn_global, int64_t* knn_local, uint8_t* knn_mask, int32_t* out_status, void* ws,
                              size_t ws_bytes, void* stream);
int lcr_node_scores(const float* feats, int channels, const int64_t* node_off, const uint8_t* node_mask, int n_pairs,
                    int m_max, int n_max, float* out, uint8_t* row_mask, uint8_t* col_mask, void* stream);
int lcr_coarse_matching_batched(const float* log_scores, int n_pairs, int rows, int cols, int32_t* out_i, int32_t* out_j,
                                float* out_scores, int32_t* out_count, void* ws, size_t ws_bytes, void* stream);
int lcr_gather_coarse(const int32_t* out_i, const int32_t* out_j, const float* out_scores, int capacity,
                      const int32_t* patch_off, const int64_t* node_off, int n_pairs, int32_t* ci_global,
                      int32_t* cj_global, int32_t* ci_local, int32_t* cj_local, float* scores, int32_t* patch_pair,
                      void* stream);
size_t lcr_lgr_batched_ws_bytes(int n_patches, int n_scan_pairs, int64_t capacity);
int lcr_lgr_batched(const float* ref, const float* src, const float* scores, const int32_t* c_pair,
                    const int32_t* pair_off, int n_patches, const int32_t* patch_pair, const int32_t* patch_off,
                    int n_scan_pairs, int64_t capacity, float radius, int min_corr, int steps, float* out_T, void* ws,
                    size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * RANSAC rigid registration from correspondences (utils/utils/open3d.py:145-173:
 * open3d registration_ransac_based_on_correspondence, point-to-point estimation, ransac_n = 3,
 * distance_threshold = 0.05, num_iterations = 10000).  src / ref float[n,3] device (correspondence t =
 * (src[t], ref[t])); out_T float[16] row-major (maps src -> ref), out_best int32[2] = {winning
 * hypothesis, its inlier count}.  Sample k of hypothesis h = hash(seed, h, k) % n (reproducible; the
 * oracle evaluates the same samples).  open3d itself is third party: parity unpinned at that boundary.
 * ---------------------------------------------------------------------------------------- */
size_t lcr_ransac_ws_bytes(int num_iterations);
int lcr_ransac_correspondences(const float* src, const float* ref, int64_t n, float distance_threshold, int ransac_n,
                               int num_iterations, uint64_t seed, float* out_T, int32_t* out_best, void* ws,
                               size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LCR_B200_H_ */
