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

#ifdef __cplusplus
}
#endif
#endif /* LCR_B200_H_ */
