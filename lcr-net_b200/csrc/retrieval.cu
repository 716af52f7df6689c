// retrieval.cu -- a15: brute-force exact squared-L2 top-k over a descriptor database.
// Reference: the faiss IndexIVFFlat(nlist=1) train/add/search loop of
// experiments/loop_detection/eval_loop_detection_overlap_dataset.py:183-214 and
// experiments/inference/infer_loop_detection_find_top1.py:79-104 (exact L2 top-k; the reference
// rebuilds the index per query so that query i only sees rows [0, i-100): `valid_counts`).
//
// One CTA = 32 queries against the whole database: the 32x256 query tile stays in shared
// memory, database rows stream through in 64-row tiles (coalesced float4 loads), each thread
// accumulates a 2x4 block of d2 = sum (q-d)^2 in fp32 (direct differences: no cancellation, so
// near-duplicate descriptors order correctly).  Every warp owns 4 queries and keeps their
// running top-k as sorted (d2, idx) lists in shared memory; a candidate is inserted only if it
// beats the current k-th entry, ties broken by ascending index.
#include "common.cuh"

namespace {
constexpr int D = 256;       // descriptor size
constexpr int QT = 32;       // queries per CTA
constexpr int DT = 64;       // database rows per tile
constexpr int KMAX = 64;     // largest supported k
constexpr int DS = 32;       // K-slab (floats) staged per step

__global__ void __launch_bounds__(256)
l2_topk_kernel(const float* __restrict__ q, int nq, const float* __restrict__ db, int ndb,
               const int32_t* __restrict__ valid_counts, int k, float* __restrict__ out_d2,
               int64_t* __restrict__ out_idx) {
  __shared__ __align__(16) float s_q[QT][DS + 1];
  __shared__ __align__(16) float s_d[DT][DS + 1];
  __shared__ float s_dist[QT][DT + 1];
  __shared__ float l_d[QT][KMAX];
  __shared__ int l_i[QT][KMAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * QT;
  for (int i = tid; i < QT * KMAX; i += 256) {
    l_d[i / KMAX][i % KMAX] = INFINITY;
    l_i[i / KMAX][i % KMAX] = -1;
  }
  // the largest database prefix any query of this CTA may see
  int limit = ndb;
  if (valid_counts) {
    limit = 0;
    for (int i = 0; i < QT && q0 + i < nq; i++) limit = max(limit, min(valid_counts[q0 + i], ndb));
  }
  const int tq = tid >> 4;        // 16 query pairs: queries 2*tq, 2*tq+1
  const int td = tid & 15;        // 16 db quads:   rows td, td+16, td+32, td+48
  __syncthreads();
  for (int d0 = 0; d0 < limit; d0 += DT) {
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    for (int c0 = 0; c0 < D; c0 += DS) {
      {  // stage query slab (32 x 32) and database slab (64 x 32): 8 float4 per row
        const int r = tid >> 3, c4 = tid & 7;
        const int gq = q0 + r;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 v = gq < nq ? *reinterpret_cast<const float4*>(q + (size_t)gq * D + c0 + c4 * 4) : z;
        s_q[r][c4 * 4 + 0] = v.x; s_q[r][c4 * 4 + 1] = v.y; s_q[r][c4 * 4 + 2] = v.z; s_q[r][c4 * 4 + 3] = v.w;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int rr = r + 32 * h, gd = d0 + rr;
          const float4 w = gd < limit ? *reinterpret_cast<const float4*>(db + (size_t)gd * D + c0 + c4 * 4) : z;
          s_d[rr][c4 * 4 + 0] = w.x; s_d[rr][c4 * 4 + 1] = w.y; s_d[rr][c4 * 4 + 2] = w.z; s_d[rr][c4 * 4 + 3] = w.w;
        }
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < DS; c++) {
        const float qa = s_q[2 * tq][c], qb = s_q[2 * tq + 1][c];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float dv = s_d[td + 16 * j][c];
          const float ea = qa - dv, eb = qb - dv;
          acc[0][j] = fmaf(ea, ea, acc[0][j]);
          acc[1][j] = fmaf(eb, eb, acc[1][j]);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) s_dist[2 * tq + i][td + 16 * j] = acc[i][j];
    __syncthreads();
    // selection: warp w owns queries 4w..4w+3
    for (int qi = warp * 4; qi < warp * 4 + 4; qi++) {
      const int gq = q0 + qi;
      if (gq >= nq) break;
      const int vc = valid_counts ? min(valid_counts[gq], ndb) : ndb;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int col = lane + 32 * half, gd = d0 + col;
        const float dist = s_dist[qi][col];
        // strict (d2, idx) order: a later index never displaces an equal distance
        bool cand = gd < vc && dist < l_d[qi][k - 1];
        unsigned m = __ballot_sync(0xffffffffu, cand);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float cd = __shfl_sync(0xffffffffu, dist, src);
          const int ci = d0 + src + 32 * half;
          if (!(cd < l_d[qi][k - 1])) continue;  // the list may have tightened since the ballot
          // position = number of entries ordered before the candidate (entries with equal
          // distance have smaller indices because the database is scanned in ascending order)
          int before = 0;
          for (int e = lane; e < k; e += 32) before += (l_d[qi][e] <= cd);
          before = lcr_warp_sum(before);
          // shift [before, k-1) right by one
          float mv_d[2];
          int mv_i[2];
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const int e = lane + 32 * t;
            const bool mv = e > before && e < k;
            mv_d[t] = mv ? l_d[qi][e - 1] : 0.f;
            mv_i[t] = mv ? l_i[qi][e - 1] : 0;
          }
          __syncwarp();
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const int e = lane + 32 * t;
            if (e > before && e < k) {
              l_d[qi][e] = mv_d[t];
              l_i[qi][e] = mv_i[t];
            }
          }
          if (lane == 0) {
            l_d[qi][before] = cd;
            l_i[qi][before] = ci;
          }
          __syncwarp();
        }
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < QT * k; i += 256) {
    const int qi = i / k, e = i % k, gq = q0 + qi;
    if (gq < nq) {
      out_d2[(size_t)gq * k + e] = l_d[qi][e];
      out_idx[(size_t)gq * k + e] = (int64_t)l_i[qi][e];
    }
  }
}
}  // namespace

extern "C" int lcr_l2_topk(const float* queries, int64_t n_queries, const float* db, int64_t n_db, int dim, int k,
                           const int32_t* valid_counts, float* out_d2, int64_t* out_idx, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(dim == D, "l2_topk: descriptor dimension must be 256");
  LCR_REQUIRE(k >= 1 && k <= KMAX, "l2_topk: k must be in [1, 64]");
  LCR_REQUIRE(n_queries >= 0 && n_db >= 0 && n_queries < (1ll << 31) && n_db < (1ll << 31), "l2_topk: sizes");
  if (n_queries == 0) return LCR_OK;
  LcrProfScope prof("l2_topk", 3.0 * n_queries * (double)n_db * D, 4.0 * D * (double)(n_queries + n_db) + 12.0 * n_queries * k,
                    stream);
  l2_topk_kernel<<<(unsigned)((n_queries + QT - 1) / QT), 256, 0, stream>>>(queries, (int)n_queries, db, (int)n_db,
                                                                          valid_counts, k, out_d2, out_idx);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
