// matching.cu -- a10-a13 of the registration path:
//   point -> node partition            modules/ops/pointcloud_partition.py:61-107, pairwise_distance.py:4-31
//   log-domain Sinkhorn with dustbin   modules/sinkhorn/learnable_sinkhorn.py:13-66
//   coarse (node) correspondences      modules/geotransformer/superpoint_matching.py:129-160
//   patch score matrices               model_family/LCRNet.py:231-249
//   fine correspondences               modules/geotransformer/local_global_registration.py:49-92
//   local-to-global registration       local_global_registration.py:140-202, registration/procrustes.py:6-73
// SFU / latency-bound small problems: one CTA per problem, everything stays on the device.
#include <string.h>

#include "common.cuh"
#include "kabsch.cuh"

namespace {

constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

// ================================================================== point -> node partition
// owner[j] = argmin_i clamp(|n_i|^2 - 2 n_i.p_j + |p_j|^2, 1e-12)   (first minimum), d2 kept for the sort
__global__ void __launch_bounds__(256)
owner_kernel(const float* __restrict__ pts, int N, const float* __restrict__ nodes, int M,
             int32_t* __restrict__ owner, float* __restrict__ owner_d2, uint32_t* __restrict__ node_count) {
  extern __shared__ float4 s_nodes[];  // x, y, z, |n|^2
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float x = nodes[3 * i], y = nodes[3 * i + 1], z = nodes[3 * i + 2];
    s_nodes[i] = make_float4(x, y, z, x * x + y * y + z * z);
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const float px = pts[3 * j], py = pts[3 * j + 1], pz = pts[3 * j + 2];
  const float p2 = px * px + py * py + pz * pz;
  float best = INFINITY;
  int bi = 0;
  for (int i = 0; i < M; i++) {
    const float4 n = s_nodes[i];
    const float xy = n.x * px + n.y * py + n.z * pz;
    const float d = fmaxf(n.w - 2.f * xy + p2, 1e-12f);
    if (d < best) {
      best = d;
      bi = i;
    }
  }
  owner[j] = bi;
  owner_d2[j] = best;
  atomicAdd(&node_count[bi], 1u);
}

__global__ void node_scan_kernel(const uint32_t* __restrict__ node_count, int M, uint32_t* __restrict__ node_start,
                                 uint8_t* __restrict__ node_mask) {
  // single CTA, M is a few hundred to a few thousand
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < M; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < M ? node_count[i] : 0u;
    // block inclusive scan via warp scans
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < (blockDim.x >> 5) ? wsum[lane] : 0u, wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t ex = incl - v + wsum[warp] + carry;
    if (i < M) {
      node_start[i] = ex;
      node_mask[i] = v > 0;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = ex + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) node_start[M] = carry;
}

__global__ void node_scatter_kernel(const int32_t* __restrict__ owner, const float* __restrict__ owner_d2, int N,
                                    const uint32_t* __restrict__ node_start, uint32_t* __restrict__ node_cursor,
                                    unsigned long long* __restrict__ keys) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const int o = owner[j];
  const uint32_t pos = node_start[o] + atomicAdd(&node_cursor[o], 1u);
  keys[pos] = ((unsigned long long)__float_as_uint(owner_d2[j]) << 32) | (unsigned)j;
}

// one CTA per node: sort its owned points by (d2, index), emit the first K
constexpr int kNodeCap = 4096;
__global__ void __launch_bounds__(256)
node_topk_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ node_start, int N, int K,
                 int idx_is64, void* __restrict__ knn_idx, uint8_t* __restrict__ knn_mask, int* __restrict__ err) {
  __shared__ unsigned long long s[kNodeCap];
  const int node = blockIdx.x;
  const uint32_t st = node_start[node];
  int cnt = (int)(node_start[node + 1] - st);
  if (cnt > kNodeCap) {
    if (threadIdx.x == 0) *err = LCR_ERR_OVERFLOW;
    cnt = kNodeCap;
  }
  int n = 32;
  while (n < cnt) n <<= 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = i < cnt ? keys[st + i] : kEmptyKey;
  __syncthreads();
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], b = s[ixj];
          if ((a > b) == ((i & k) == 0)) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  for (int t = threadIdx.x; t < K; t += blockDim.x) {
    const bool ok = t < cnt;
    const long long v = ok ? (long long)(unsigned)(s[t] & 0xFFFFFFFFull) : (long long)N;
    if (idx_is64) ((int64_t*)knn_idx)[(size_t)node * K + t] = v;
    else ((int32_t*)knn_idx)[(size_t)node * K + t] = (int32_t)v;
    knn_mask[(size_t)node * K + t] = ok;
  }
}

// ================================================================== coarse correspondences
// log_scores [(M+1) x (N+1)]: (i,j) kept iff it is the column-argmax and beats the dustbin row, or
// the row-argmax and beats the dustbin column (exp domain, strict >); listed row-major.
__global__ void __launch_bounds__(1024)
coarse_match_kernel(const float* __restrict__ ls, int M, int N, int32_t* __restrict__ out_i,
                    int32_t* __restrict__ out_j, float* __restrict__ out_s, int32_t* __restrict__ out_n,
                    int32_t* __restrict__ row_best, int32_t* __restrict__ col_best, int32_t* __restrict__ row_cnt) {
  const int R = M + 1, C = N + 1;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  {  // one CTA per problem of a batch (capacity M + N outputs per problem)
    const size_t b = blockIdx.x;
    ls += b * R * C;
    out_i += b * (M + N);
    out_j += b * (M + N);
    out_s += b * (M + N);
    out_n += b;
    row_best += b * R;
    row_cnt += b * R;
    col_best += b * C;
  }
  // exp() is monotone, so arg-maxima are taken on exp(ls) exactly like the reference (ties: first)
  for (int i = warp; i < R; i += nw) {
    float best = -INFINITY;
    int bj = 0x7fffffff;
    for (int j = lane; j < C; j += 32) {
      const float e = expf(ls[(size_t)i * C + j]);
      if (e > best) { best = e; bj = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
    }
    if (lane == 0) row_best[i] = bj;
  }
  for (int j = tid; j < C; j += nt) {
    float best = -INFINITY;
    int bi = 0;
    for (int i = 0; i < R; i++) {
      const float e = expf(ls[(size_t)i * C + j]);
      if (e > best) { best = e; bi = i; }
    }
    col_best[j] = bi;
  }
  __syncthreads();
  auto kept = [&](int i, int j) -> bool {
    const float e = expf(ls[(size_t)i * C + j]);
    const bool by_col = col_best[j] == i && e > expf(ls[(size_t)M * C + j]);
    const bool by_row = row_best[i] == j && e > expf(ls[(size_t)i * C + N]);
    return by_col || by_row;
  };
  for (int i = warp; i < M; i += nw) {
    int c = 0;
    for (int j = lane; j < N; j += 32) c += kept(i, j);
    c = lcr_warp_sum(c);
    if (lane == 0) row_cnt[i] = c;
  }
  __syncthreads();
  if (tid == 0) {  // M is a few hundred: serial exclusive scan
    int acc = 0;
    for (int i = 0; i < M; i++) {
      const int c = row_cnt[i];
      row_cnt[i] = acc;
      acc += c;
    }
    *out_n = acc;
  }
  __syncthreads();
  for (int i = warp; i < M; i += nw) {
    int base = row_cnt[i];
    for (int j0 = 0; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const bool k = j < N && kept(i, j);
      const unsigned m = __ballot_sync(0xffffffffu, k);
      if (k) {
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        out_i[pos] = i;
        out_j[pos] = j;
        out_s[pos] = expf(ls[(size_t)i * C + j]);
      }
      base += __popc(m);
    }
  }
}

// ================================================================== patch score matrices
// out[p, a, b] = <fa[knn_a[p, a]], fb[knn_b[p, b]]> / sqrt(C)  (pad index -> zero row); K = 128, C = 128.
// One CTA per patch pair, 8 x 8 register tile per thread, operands transposed into shared memory ([channel][point])
// in two 64-channel stages (66 KB: two CTAs per SM, so one CTA's gather overlaps the other's FMAs).  Staging map:
// a warp covers 16 points x 2 float4 per step, lane = (point & 15, float4 & 1): the transposed store
// sa[4 c4 + k][r] then lands on bank 16 (c4 & 1) + (r & 15) -- conflict free (the first version had the 32 lanes
// on the 32 float4 of ONE point: 16-way conflicts on every store, 2/3 of the kernel time).
constexpr int PK = 128, PC = 128, PCH = 64;
__global__ void __launch_bounds__(256, 2)
patch_scores_kernel(const float* __restrict__ fa, int na, const int32_t* __restrict__ knn_a,
                    const int32_t* __restrict__ node_a, const float* __restrict__ fb, int nb,
                    const int32_t* __restrict__ knn_b, const int32_t* __restrict__ node_b, float div,
                    float* __restrict__ out) {
  extern __shared__ float sp[];  // A^T [PCH][PK+4], B^T [PCH][PK+4]
  float(*sa)[PK + 4] = reinterpret_cast<float(*)[PK + 4]>(sp);
  float(*sb)[PK + 4] = reinterpret_cast<float(*)[PK + 4]>(sp + PCH * (PK + 4));
  const int p = blockIdx.x, tid = threadIdx.x;
  const int r = (tid >> 5) * 16 + (tid & 15), c4l = (tid >> 4) & 1;   // staging: this thread's point and float4 parity
  const int ia = knn_a[(size_t)node_a[p] * PK + r], ib = knn_b[(size_t)node_b[p] * PK + r];
  const float4* ra = ia < na ? reinterpret_cast<const float4*>(fa + (size_t)ia * PC) : nullptr;
  const float4* rb = ib < nb ? reinterpret_cast<const float4*>(fb + (size_t)ib * PC) : nullptr;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
#pragma unroll 1
  for (int half = 0; half < PC / PCH; half++) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 va[8], vb[8];
#pragma unroll
    for (int it = 0; it < 8; it++) {           // all 16 loads in flight before the first store
      const int c4 = half * (PCH / 4) + 2 * it + c4l;
      va[it] = ra ? ra[c4] : z;
      vb[it] = rb ? rb[c4] : z;
    }
    if (half) __syncthreads();                 // the previous stage has been consumed
#pragma unroll
    for (int it = 0; it < 8; it++) {
      const int c = 4 * (2 * it + c4l);
      sa[c + 0][r] = va[it].x; sa[c + 1][r] = va[it].y; sa[c + 2][r] = va[it].z; sa[c + 3][r] = va[it].w;
      sb[c + 0][r] = vb[it].x; sb[c + 1][r] = vb[it].y; sb[c + 2][r] = vb[it].z; sb[c + 3][r] = vb[it].w;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < PCH; c++) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&sa[c][ty * 4]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&sa[c][64 + ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&sb[c][tx * 4]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&sb[c][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* o = out + (size_t)p * PK * PK;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int row = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int cidx = h * 64 + tx * 4;
      *reinterpret_cast<float4*>(o + (size_t)row * PK + cidx) =
          make_float4(acc[i][h * 4 + 0] / div, acc[i][h * 4 + 1] / div, acc[i][h * 4 + 2] / div, acc[i][h * 4 + 3] / div);
    }
  }
}

// ------------------------------------------------------------------ the same product on the warp-level tensor path
// 128 x 128 x 128 per patch pair with mma.sync.m16n8k8 TF32 and the 3xTF32 operand split of the GEMMs (hi = upper 19
// bits, lo = x - hi; lo.hi + hi.lo + hi.hi, fp32 accumulate: fp32-class accuracy).  Both gathered operands stay in their
// natural [point][channel] layout in shared memory (row.col fragments of A . B^T read rows of both); the row stride of
// 68 floats puts the eight rows x four columns of a fragment load on 32 distinct banks.  8 warps, each a 64 x 32
// output block; two 64-channel stages, two CTAs per SM.  (A 128-row tcgen05 tile per pair would need the gather to
// write swizzled hi / lo operand tiles; the legacy tensor path already makes the kernel MMA-bound instead of
// FFMA-bound: 2.2 -> see DESIGN.md.)
constexpr int PSS = PCH + 4;
__device__ __forceinline__ void ps_mma(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ps_split(float x, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

__global__ void __launch_bounds__(256, 2)
patch_scores_mma_kernel(const float* __restrict__ fa, int na, const int32_t* __restrict__ knn_a,
                        const int32_t* __restrict__ node_a, const float* __restrict__ fb, int nb,
                        const int32_t* __restrict__ knn_b, const int32_t* __restrict__ node_b, float div,
                        float* __restrict__ out) {
  extern __shared__ __align__(16) float sp[];  // A [PK][PSS], B [PK][PSS]
  float(*sa)[PSS] = reinterpret_cast<float(*)[PSS]>(sp);
  float(*sb)[PSS] = reinterpret_cast<float(*)[PSS]>(sp + PK * PSS);
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int row0 = (warp & 1) * 64, col0 = (warp >> 1) * 32;
  const int32_t* ka = knn_a + (size_t)node_a[p] * PK;
  const int32_t* kb = knn_b + (size_t)node_b[p] * PK;
  // staging: 16 threads per 64-channel half row, 8 rows per thread and operand
  const int sr = tid >> 4, sc = (tid & 15) * 4;
  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[i][j][e] = 0.f;
#pragma unroll 1
  for (int half = 0; half < PC / PCH; half++) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 va[8], vb[8];
#pragma unroll
    for (int it = 0; it < 8; it++) {           // all 16 loads in flight before the first store
      const int ia = ka[sr + 16 * it], ib = kb[sr + 16 * it];     // (re-read per stage: L1 hits, 16 registers less)
      va[it] = ia < na ? *reinterpret_cast<const float4*>(fa + (size_t)ia * PC + half * PCH + sc) : z;
      vb[it] = ib < nb ? *reinterpret_cast<const float4*>(fb + (size_t)ib * PC + half * PCH + sc) : z;
    }
    if (half) __syncthreads();                 // the previous stage has been consumed
#pragma unroll
    for (int it = 0; it < 8; it++) {
      *reinterpret_cast<float4*>(&sa[sr + 16 * it][sc]) = va[it];
      *reinterpret_cast<float4*>(&sb[sr + 16 * it][sc]) = vb[it];
    }
    __syncthreads();
#pragma unroll 2
    for (int ks = 0; ks < PCH / 8; ks++) {
      const int k0 = ks * 8 + t;
      unsigned bh[4][2], bl[4][2];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float* r = &sb[col0 + 8 * j + g][k0];
        ps_split(r[0], bh[j][0], bl[j][0]);
        ps_split(r[4], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float* r0 = &sa[row0 + 16 * i + g][k0];
        const float* r1 = r0 + 8 * PSS;
        unsigned ah[4], al[4];
        ps_split(r0[0], ah[0], al[0]);
        ps_split(r1[0], ah[1], al[1]);
        ps_split(r0[4], ah[2], al[2]);
        ps_split(r1[4], ah[3], al[3]);
        // three passes over the four n-tiles: consecutive MMAs never share an accumulator
#pragma unroll
        for (int j = 0; j < 4; j++) ps_mma(acc[i][j], al, bh[j][0], bh[j][1]);
#pragma unroll
        for (int j = 0; j < 4; j++) ps_mma(acc[i][j], ah, bl[j][0], bl[j][1]);
#pragma unroll
        for (int j = 0; j < 4; j++) ps_mma(acc[i][j], ah, bh[j][0], bh[j][1]);
      }
    }
  }
  float* o = out + (size_t)p * PK * PK;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = row0 + 16 * i + g, c = col0 + 8 * j + 2 * t;
      *reinterpret_cast<float2*>(o + (size_t)r * PK + c) = make_float2(acc[i][j][0] / div, acc[i][j][1] / div);
      *reinterpret_cast<float2*>(o + (size_t)(r + 8) * PK + c) = make_float2(acc[i][j][2] / div, acc[i][j][3] / div);
    }
}

// ================================================================== fine correspondences
// per pair p: log scores [129 x 129]; (i,j), i,j < 128, kept iff (row-argmax and > dustbin col) or
// (col-argmax and > dustbin row) in the exp domain, and both points valid (LCRNet.py:239-249).
// Only the row / column arg-max entries can be kept: at most 256 candidates per pair.  Pass 1 (one CTA per pair)
// stages E = exp(S) in shared memory ONCE (66.5 KB, 3 CTAs per SM; the first version re-read S from L2 seven times:
// row arg-max, column arg-max, the 128 x 128 `kept` scan, all of it again in the emit pass), finds the arg-maxima,
// marks the kept candidates in a 128 x 128 bitmap and writes them in row-major order into the pair's 256-entry slot
// of a scratch buffer; after the scan of the counts, pass 2 copies the slots to their final offsets.
constexpr int FR = PK + 1;
constexpr size_t kFineSmem = sizeof(float) * FR * FR + sizeof(int) * (2 * FR + PK + 4 * PK) + 64;
__global__ void __launch_bounds__(256, 3)
fine_corr_find_kernel(const float* __restrict__ ls, const uint8_t* __restrict__ mask_a,
                      const int32_t* __restrict__ node_a, const uint8_t* __restrict__ mask_b,
                      const int32_t* __restrict__ node_b, int32_t* __restrict__ pair_cnt,
                      int32_t* __restrict__ slot_ij, float* __restrict__ slot_s) {
  extern __shared__ __align__(16) float fsm[];
  float* E = fsm;                                          // [129][129]
  int* row_best = reinterpret_cast<int*>(E + FR * FR);     // [129]
  int* col_best = row_best + FR;                           // [129]
  int* row_off = col_best + FR;                            // [128]
  unsigned* bm = reinterpret_cast<unsigned*>(row_off + PK);   // [128][4]
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* S = ls + (size_t)p * FR * FR;
  const uint8_t* ma = mask_a + (size_t)node_a[p] * PK;
  const uint8_t* mb = mask_b + (size_t)node_b[p] * PK;
  for (int e = tid; e < FR * FR; e += 256) E[e] = expf(S[e]);
  for (int e = tid; e < 4 * PK; e += 256) bm[e] = 0u;
  __syncthreads();
  for (int i = warp; i < FR; i += 8) {                     // row arg-max, lowest index among equals
    float best = -INFINITY;
    int bj = 0x7fffffff;
    for (int j = lane; j < FR; j += 32) {
      const float e = E[i * FR + j];
      if (e > best) { best = e; bj = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
    }
    if (lane == 0) row_best[i] = bj;
  }
  if (tid < FR) {                                          // column arg-max, lowest index among equals
    float b0 = -INFINITY, b1 = -INFINITY;
    int i0 = 0, i1 = 0;
    for (int i = 0; i + 1 < FR; i += 2) {                  // two interleaved chains (even / odd rows)
      const float e0 = E[i * FR + tid], e1 = E[(i + 1) * FR + tid];
      if (e0 > b0) { b0 = e0; i0 = i; }
      if (e1 > b1) { b1 = e1; i1 = i + 1; }
    }
    const float el = E[(FR - 1) * FR + tid];               // FR is odd: the last row
    if (el > b0) { b0 = el; i0 = FR - 1; }
    col_best[tid] = (b1 > b0 || (b1 == b0 && i1 < i0)) ? i1 : i0;
  }
  __syncthreads();
  if (tid < PK) {
    const int i = tid, j = row_best[i];
    if (j < PK && ma[i] && mb[j] && E[i * FR + j] > E[i * FR + PK]) atomicOr(&bm[i * 4 + (j >> 5)], 1u << (j & 31));
  } else {
    const int j = tid - PK, i = col_best[j];
    if (i < PK && ma[i] && mb[j] && E[i * FR + j] > E[PK * FR + j]) atomicOr(&bm[i * 4 + (j >> 5)], 1u << (j & 31));
  }
  __syncthreads();
  if (tid < PK) {                                          // exclusive scan of the row counts (4 warps)
    const int c = __popc(bm[tid * 4]) + __popc(bm[tid * 4 + 1]) + __popc(bm[tid * 4 + 2]) + __popc(bm[tid * 4 + 3]);
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    row_off[tid] = incl - c;
    if (lane == 31) row_best[warp] = incl;                 // (row_best is dead by now: warp totals)
  }
  __syncthreads();
  if (tid < PK) {
    int base = row_off[tid];
    for (int w = 0; w < warp; w++) base += row_best[w];
    int32_t* dij = slot_ij + (size_t)p * 256;
    float* ds = slot_s + (size_t)p * 256;
#pragma unroll
    for (int w = 0; w < 4; w++) {
      unsigned bits = bm[tid * 4 + w];
      while (bits) {
        const int j = 32 * w + __ffs(bits) - 1;
        bits &= bits - 1;
        dij[base] = (tid << 8) | j;
        ds[base] = E[tid * FR + j];
        base++;
      }
    }
    if (tid == PK - 1) pair_cnt[p] = base;
  }
}

__global__ void fine_corr_compact_kernel(const int32_t* __restrict__ slot_ij, const float* __restrict__ slot_s,
                                         const int32_t* __restrict__ pair_cnt, const int32_t* __restrict__ pair_off,
                                         int n_pairs, int32_t* __restrict__ out_pair, int32_t* __restrict__ out_i,
                                         int32_t* __restrict__ out_j, float* __restrict__ out_s) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int p = (int)(idx >> 8), t = (int)(idx & 255);
  if (p >= n_pairs || t >= pair_cnt[p]) return;
  const int ij = slot_ij[idx];
  const int pos = pair_off[p] + t;
  out_pair[pos] = p;
  out_i[pos] = ij >> 8;
  out_j[pos] = ij & 255;
  out_s[pos] = slot_s[idx];
}

__global__ void exclusive_scan_i32_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out) {
  // single CTA of 1024 threads, chunked
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = lane < (blockDim.x >> 5) ? wsum[lane] : 0, wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const int ex = incl - v + wsum[warp] + carry;
    if (i < n) out[i] = ex;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = ex + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry;
}

// gather correspondence points: ref = pts_a[knn_a[node_a[pair], i]], src = pts_b[knn_b[node_b[pair], j]]
__global__ void corr_points_kernel(const int32_t* __restrict__ c_pair, const int32_t* __restrict__ c_i,
                                   const int32_t* __restrict__ c_j, const int32_t* __restrict__ n_corr,
                                   const float* __restrict__ pts_a, const int32_t* __restrict__ knn_a,
                                   const int32_t* __restrict__ node_a, const float* __restrict__ pts_b,
                                   const int32_t* __restrict__ knn_b, const int32_t* __restrict__ node_b,
                                   float* __restrict__ ref, float* __restrict__ src) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *n_corr) return;
  const int p = c_pair[t];
  const int ia = knn_a[(size_t)node_a[p] * PK + c_i[t]], ib = knn_b[(size_t)node_b[p] * PK + c_j[t]];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    ref[3 * t + d] = pts_a[3 * (size_t)ia + d];
    src[3 * t + d] = pts_b[3 * (size_t)ib + d];
  }
}

// ================================================================== weighted Procrustes
// (kabsch_rotation: 3x3 SVD by one-sided Jacobi in double precision, kabsch.cuh)
// One CTA per segment [seg_off[s], seg_off[s+1]) of the correspondence arrays (or the whole range
// when seg_off == NULL).  weights = w[t] (>= 0 enforced), normalised by (sum + 1e-5).
// Segments shorter than min_count are skipped (valid[s] = 0).
__global__ void __launch_bounds__(256)
procrustes_kernel(const float* __restrict__ src, const float* __restrict__ ref, const float* __restrict__ w,
                  const int32_t* __restrict__ seg_off, const int32_t* __restrict__ n_total, int min_count,
                  float* __restrict__ T_out /*[S,16]*/, int32_t* __restrict__ valid) {
  __shared__ double red[8][13];
  __shared__ double tot[13];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t0 = seg_off ? seg_off[s] : 0, t1 = seg_off ? seg_off[s + 1] : *n_total;
  if (t1 - t0 < min_count) {
    if (tid == 0 && valid) valid[s] = 0;
    return;
  }
  auto block_sum = [&](double* vals, int n) {
    for (int k = 0; k < n; k++) {
      double x = vals[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) red[warp][k] = x;
    }
    __syncthreads();
    if (tid < n) {
      double x = 0;
      for (int ww = 0; ww < 8; ww++) x += red[ww][tid];
      tot[tid] = x;
    }
    __syncthreads();
  };
  // pass 1: sum of weights, weighted sums of src and ref
  double v[13];
  for (int k = 0; k < 13; k++) v[k] = 0;
  for (int t = t0 + tid; t < t1; t += 256) {
    const double ww = fmax((double)w[t], 0.0);
    v[0] += ww;
    for (int d = 0; d < 3; d++) {
      v[1 + d] += ww * (double)src[3 * t + d];
      v[4 + d] += ww * (double)ref[3 * t + d];
    }
  }
  block_sum(v, 7);
  const double wsum = tot[0] + 1e-5;
  double sc[3], rc[3];
  for (int d = 0; d < 3; d++) {
    sc[d] = tot[1 + d] / wsum;
    rc[d] = tot[4 + d] / wsum;
  }
  __syncthreads();
  // pass 2: H = sum w/wsum * (src - sc)(ref - rc)^T
  for (int k = 0; k < 13; k++) v[k] = 0;
  for (int t = t0 + tid; t < t1; t += 256) {
    const double ww = fmax((double)w[t], 0.0) / wsum;
    double a[3], b[3];
    for (int d = 0; d < 3; d++) {
      a[d] = (double)src[3 * t + d] - sc[d];
      b[d] = (double)ref[3 * t + d] - rc[d];
    }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) v[3 * i + j] += ww * a[i] * b[j];
  }
  block_sum(v, 9);
  if (tid == 0) {
    double H[3][3], R[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) H[i][j] = tot[3 * i + j];
    kabsch_rotation(H, R);
    float* T = T_out + (size_t)s * 16;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) T[4 * i + j] = (float)R[i][j];
      T[4 * i + 3] = (float)(rc[i] - (R[i][0] * sc[0] + R[i][1] * sc[1] + R[i][2] * sc[2]));
    }
    T[12] = T[13] = T[14] = 0.f;
    T[15] = 1.f;
    if (valid) valid[s] = 1;
  }
}

__device__ __forceinline__ bool is_inlier(const float* T, const float* src, const float* ref, int t, float radius) {
  const float x = src[3 * t], y = src[3 * t + 1], z = src[3 * t + 2];
  const float ax = T[0] * x + T[1] * y + T[2] * z + T[3], ay = T[4] * x + T[5] * y + T[6] * z + T[7],
              az = T[8] * x + T[9] * y + T[10] * z + T[11];
  const float dx = ref[3 * t] - ax, dy = ref[3 * t + 1] - ay, dz = ref[3 * t + 2] - az;
  return sqrtf(dx * dx + dy * dy + dz * dz) < radius;
}

// inlier count of every local transform over ALL correspondences; one CTA per segment
// (batched: hypothesis s belongs to scan pair seg_pair[s], whose correspondences are [corr_off[sp], corr_off[sp+1]))
__global__ void __launch_bounds__(256)
inlier_count_kernel(const float* __restrict__ T, const int32_t* __restrict__ valid, const float* __restrict__ src,
                    const float* __restrict__ ref, const int32_t* __restrict__ n_total, float radius,
                    int32_t* __restrict__ counts, const int32_t* __restrict__ seg_pair,
                    const int32_t* __restrict__ corr_off) {
  __shared__ int s_c[8];
  const int s = blockIdx.x;
  if (!valid[s]) {
    if (threadIdx.x == 0) counts[s] = -1;
    return;
  }
  const float* Ts = T + (size_t)s * 16;
  int c = 0;
  const int sp = seg_pair ? seg_pair[s] : 0;
  const int t_lo = corr_off ? corr_off[sp] : 0, n = corr_off ? corr_off[sp + 1] : *n_total;
  for (int t = t_lo + threadIdx.x; t < n; t += 256) c += is_inlier(Ts, src, ref, t, radius);
  c = lcr_warp_sum(c);
  if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; w++) tot += s_c[w];
    counts[s] = tot;
  }
}

// best = first argmax of counts (>= 0); copies T[best] to T_cur, or flags "no local transform"
// (batched: CTA b picks among the hypotheses [seg_off[b], seg_off[b+1]) of scan pair b)
__global__ void pick_best_kernel(const int32_t* __restrict__ counts, int S, const float* __restrict__ T,
                                 float* __restrict__ T_cur, int32_t* __restrict__ have_local,
                                 const int32_t* __restrict__ seg_off) {
  __shared__ int s_best[32], s_idx[32];
  int best = -1, bi = 0x7fffffff;
  const int s_lo = seg_off ? seg_off[blockIdx.x] : 0, s_hi = seg_off ? seg_off[blockIdx.x + 1] : S;
  T_cur += (size_t)blockIdx.x * 16;
  have_local += blockIdx.x;
  for (int s = s_lo + threadIdx.x; s < s_hi; s += blockDim.x)
    if (counts[s] > best) { best = counts[s]; bi = s; }
  for (int o = 16; o > 0; o >>= 1) {
    const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_best[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++)
      if (s_best[w] > best || (s_best[w] == best && s_idx[w] < bi)) { best = s_best[w]; bi = s_idx[w]; }
    *have_local = best >= 0;
    if (best >= 0)
      for (int k = 0; k < 16; k++) T_cur[k] = T[(size_t)bi * 16 + k];
  }
}

// w_cur[t] = score[t] * inlier(T), T = (*select != 0) ? T_a : T_b   (select == NULL -> T_a)
// (batched: correspondence t belongs to scan pair seg_pair[c_pair[t]]; T_a, T_b, select are indexed by scan pair)
__global__ void reweight_kernel(const float* __restrict__ T_a, const float* __restrict__ T_b,
                                const int32_t* __restrict__ select, const float* __restrict__ src,
                                const float* __restrict__ ref, const float* __restrict__ score,
                                const int32_t* __restrict__ n_total, float radius, float* __restrict__ w_cur,
                                const int32_t* __restrict__ c_pair, const int32_t* __restrict__ seg_pair) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *n_total) return;
  const int sp = (c_pair && seg_pair) ? seg_pair[c_pair[t]] : 0;
  const float* T = (!select || select[sp] != 0) ? T_a + (size_t)sp * 16 : T_b + (size_t)sp * 16;
  w_cur[t] = is_inlier(T, src, ref, t, radius) ? score[t] : 0.f;
}


// ================================================================== batched variants (all pairs of a chunk per launch)
// owner of every point among the nodes of ITS cloud: grid (ceil(max cloud points / 256), clouds); owner indices
// are GLOBAL node rows
__global__ void __launch_bounds__(256)
owner_batched_kernel(const float* __restrict__ pts, const int64_t* __restrict__ pts_off,
                     const float* __restrict__ nodes, const int64_t* __restrict__ node_off,
                     int32_t* __restrict__ owner, float* __restrict__ owner_d2, uint32_t* __restrict__ node_count) {
  extern __shared__ float4 s_nodes[];  // x, y, z, |n|^2
  const int b = blockIdx.y;
  const int n0 = (int)pts_off[b], n1 = (int)pts_off[b + 1], m0 = (int)node_off[b], M = (int)node_off[b + 1] - m0;
  if ((int)(blockIdx.x * blockDim.x) >= n1 - n0) return;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float x = nodes[3 * (size_t)(m0 + i)], y = nodes[3 * (size_t)(m0 + i) + 1], z = nodes[3 * (size_t)(m0 + i) + 2];
    s_nodes[i] = make_float4(x, y, z, x * x + y * y + z * z);
  }
  __syncthreads();
  const int j = n0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n1 || M <= 0) return;
  const float px = pts[3 * (size_t)j], py = pts[3 * (size_t)j + 1], pz = pts[3 * (size_t)j + 2];
  const float p2 = px * px + py * py + pz * pz;
  float best = INFINITY;
  int bi = 0;
  for (int i = 0; i < M; i++) {
    const float4 n = s_nodes[i];
    const float xy = n.x * px + n.y * py + n.z * pz;
    const float d = fmaxf(n.w - 2.f * xy + p2, 1e-12f);
    if (d < best) {
      best = d;
      bi = i;
    }
  }
  owner[j] = m0 + bi;
  owner_d2[j] = best;
  atomicAdd(&node_count[m0 + bi], 1u);
}

// one CTA per node (global row): sort its owned points by (d2, index), emit the first K as GLOBAL point rows (int32,
// pad = n_total) and as rows LOCAL to the node's cloud (int64, pad = points of that cloud: the reference's table)
__global__ void __launch_bounds__(256)
node_topk_batched_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ node_start,
                         int n_total, int K, const int64_t* __restrict__ pts_off, const int64_t* __restrict__ node_off,
                         int n_clouds, int32_t* __restrict__ knn_global, int64_t* __restrict__ knn_local,
                         uint8_t* __restrict__ knn_mask, int* __restrict__ err) {
  __shared__ unsigned long long s[kNodeCap];
  const int node = blockIdx.x;
  const uint32_t st = node_start[node];
  int cnt = (int)(node_start[node + 1] - st);
  if (cnt > kNodeCap) {
    if (threadIdx.x == 0) *err = LCR_ERR_OVERFLOW;
    cnt = kNodeCap;
  }
  const int b = lcr_find_segment(node_off, n_clouds, (int64_t)node);
  const int64_t p0 = pts_off[b], np = pts_off[b + 1] - p0;
  int n = 32;
  while (n < cnt) n <<= 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = i < cnt ? keys[st + i] : kEmptyKey;
  __syncthreads();
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], c = s[ixj];
          if ((a > c) == ((i & k) == 0)) {
            s[i] = c;
            s[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  for (int t = threadIdx.x; t < K; t += blockDim.x) {
    const bool ok = t < cnt;
    const int64_t g = ok ? (int64_t)(unsigned)(s[t] & 0xFFFFFFFFull) : (int64_t)n_total;
    knn_global[(size_t)node * K + t] = (int32_t)g;
    if (knn_local) knn_local[(size_t)node * K + t] = ok ? g - p0 : np;
    knn_mask[(size_t)node * K + t] = ok;
  }
}

// node score matrices of all pairs: out[p, i, j] = <f[node_off[2p] + i], f[node_off[2p+1] + j]> / sqrt(C), zero
// outside the pair's m x n block (LCRNet.py:196-199), plus the padded row / column masks for the Sinkhorn
constexpr int NS_T = 64, NS_K = 16;
__global__ void __launch_bounds__(256)
node_scores_kernel(const float* __restrict__ f, int C, const int64_t* __restrict__ node_off,
                   const uint8_t* __restrict__ node_mask, int m_max, int n_max, float inv_div,
                   float* __restrict__ out, uint8_t* __restrict__ row_mask, uint8_t* __restrict__ col_mask) {
  __shared__ float sa[NS_K][NS_T + 4], sb[NS_K][NS_T + 4];
  const int p = blockIdx.z, i0 = blockIdx.y * NS_T, j0 = blockIdx.x * NS_T, tid = threadIdx.x;
  const int a0 = (int)node_off[2 * p], m = (int)node_off[2 * p + 1] - a0;
  const int b0 = (int)node_off[2 * p + 1], n = (int)node_off[2 * p + 2] - b0;
  if (blockIdx.x == 0)
    for (int i = i0 + tid; i < min(i0 + NS_T, m_max); i += 256) row_mask[(size_t)p * m_max + i] = i < m ? node_mask[a0 + i] : 0;
  if (blockIdx.y == 0)
    for (int j = j0 + tid; j < min(j0 + NS_T, n_max); j += 256) col_mask[(size_t)p * n_max + j] = j < n ? node_mask[b0 + j] : 0;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  if (i0 < m && j0 < n) {
    for (int c0 = 0; c0 < C; c0 += NS_K) {
      for (int e = tid; e < NS_T * (NS_K / 4); e += 256) {
        const int r = e / (NS_K / 4), c4 = e % (NS_K / 4);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 va = i0 + r < m ? *reinterpret_cast<const float4*>(f + (size_t)(a0 + i0 + r) * C + c0 + 4 * c4) : z;
        const float4 vb = j0 + r < n ? *reinterpret_cast<const float4*>(f + (size_t)(b0 + j0 + r) * C + c0 + 4 * c4) : z;
        sa[4 * c4][r] = va.x; sa[4 * c4 + 1][r] = va.y; sa[4 * c4 + 2][r] = va.z; sa[4 * c4 + 3][r] = va.w;
        sb[4 * c4][r] = vb.x; sb[4 * c4 + 1][r] = vb.y; sb[4 * c4 + 2][r] = vb.z; sb[4 * c4 + 3][r] = vb.w;
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < NS_K; c++) {
        const float4 a4 = *reinterpret_cast<const float4*>(&sa[c][4 * ty]);
        const float4 b4 = *reinterpret_cast<const float4*>(&sb[c][4 * tx]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = i0 + 4 * ty + i;
    if (r >= m_max) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int cidx = j0 + 4 * tx + j;
      if (cidx < n_max) out[((size_t)p * m_max + r) * n_max + cidx] = (r < m && cidx < n) ? acc[i][j] * inv_div : 0.f;
    }
  }
}

// node correspondences of all pairs, compacted to one list of GLOBAL node rows: patch t in [patch_off[p],
// patch_off[p+1]) is correspondence t - patch_off[p] of pair p
__global__ void gather_coarse_kernel(const int32_t* __restrict__ out_i, const int32_t* __restrict__ out_j,
                                     const float* __restrict__ out_s, int cap, const int32_t* __restrict__ patch_off,
                                     const int64_t* __restrict__ node_off, int32_t* __restrict__ ci_g,
                                     int32_t* __restrict__ cj_g, int32_t* __restrict__ ci_l, int32_t* __restrict__ cj_l,
                                     float* __restrict__ cs, int32_t* __restrict__ patch_pair) {
  const int p = blockIdx.x, t0 = patch_off[p], cnt = patch_off[p + 1] - t0;
  const int a0 = (int)node_off[2 * p], b0 = (int)node_off[2 * p + 1];
  for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
    const int i = out_i[(size_t)p * cap + t], j = out_j[(size_t)p * cap + t];
    ci_g[t0 + t] = a0 + i;
    cj_g[t0 + t] = b0 + j;
    ci_l[t0 + t] = i;
    cj_l[t0 + t] = j;
    cs[t0 + t] = out_s[(size_t)p * cap + t];
    patch_pair[t0 + t] = p;
  }
}

// corr_off[s] = pair_off[patch_off[s]]: first correspondence of scan pair s
__global__ void corr_offsets_kernel(const int32_t* __restrict__ pair_off, const int32_t* __restrict__ patch_off, int S,
                                    int32_t* __restrict__ corr_off) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s <= S) corr_off[s] = pair_off[patch_off[s]];
}

}  // namespace

// ================================================================== C ABI
extern "C" size_t lcr_point_to_node_ws_bytes(int64_t n_points, int64_t n_nodes) {
  return lcr_align_up(n_points * 4) * 2 + lcr_align_up(n_points * 8) + lcr_align_up((n_nodes + 1) * 4) * 3 + 1024;
}

extern "C" int lcr_point_to_node(const float* points, int64_t n_points, const float* nodes, int64_t n_nodes, int k,
                                 int32_t* point_to_node, uint8_t* node_mask, void* knn_idx, int idx_is64,
                                 uint8_t* knn_mask, int32_t* out_status, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n_points >= 1 && n_nodes >= 1 && n_points < (1ll << 31) && n_nodes <= 12000,
              "point_to_node: sizes (at most 12000 nodes)");
  LCR_REQUIRE(k >= 1 && k <= kNodeCap, "point_to_node: k");
  LCR_REQUIRE(ws && ws_bytes >= lcr_point_to_node_ws_bytes(n_points, n_nodes), "point_to_node: workspace too small");
  const int N = (int)n_points, M = (int)n_nodes;
  LcrArena a(ws, ws_bytes);
  float* d2 = a.take<float>(N);
  int32_t* owner_tmp = a.take<int32_t>(N);
  unsigned long long* keys = a.take<unsigned long long>(N);
  uint32_t* count = a.take<uint32_t>(M + 1);
  uint32_t* start = a.take<uint32_t>(M + 1);
  uint32_t* cursor = a.take<uint32_t>(M + 1);
  int* err = (int*)a.take<int>(1);
  int32_t* owner = point_to_node ? point_to_node : owner_tmp;
  LcrProfScope prof("point_to_node", 8.0 * N * M, 12.0 * (N + M) + 4.0 * N + 5.0 * M * k, stream);
  LCR_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(uint32_t) * (M + 1), stream));
  LCR_CUDA_TRY(cudaMemsetAsync(cursor, 0, sizeof(uint32_t) * (M + 1), stream));
  if (out_status) LCR_CUDA_TRY(cudaMemsetAsync(out_status, 0, sizeof(int32_t), stream));
  const size_t smem = sizeof(float4) * M;
  if (smem > 48 * 1024)
    LCR_CUDA_TRY(cudaFuncSetAttribute(owner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  owner_kernel<<<(N + 255) / 256, 256, smem, stream>>>(points, N, nodes, M, owner, d2, count);
  node_scan_kernel<<<1, 1024, 0, stream>>>(count, M, start, node_mask);
  node_scatter_kernel<<<(N + 255) / 256, 256, 0, stream>>>(owner, d2, N, start, cursor, keys);
  node_topk_kernel<<<M, 256, 0, stream>>>(keys, start, N, k, idx_is64, knn_idx, knn_mask, out_status ? out_status : err);
  LCR_LAUNCHED(4);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" size_t lcr_coarse_matching_ws_bytes(int rows, int cols) {
  return lcr_align_up((size_t)(rows + 1) * 4) * 2 + lcr_align_up((size_t)(cols + 1) * 4) + 256;
}

extern "C" int lcr_coarse_matching(const float* log_scores, int rows, int cols, int32_t* out_i, int32_t* out_j,
                                   float* out_scores, int32_t* out_count, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(rows >= 1 && cols >= 1, "coarse_matching: sizes");
  LCR_REQUIRE(ws && ws_bytes >= lcr_coarse_matching_ws_bytes(rows, cols), "coarse_matching: workspace too small");
  LcrArena a(ws, ws_bytes);
  int32_t* row_best = a.take<int32_t>(rows + 1);
  int32_t* row_cnt = a.take<int32_t>(rows + 1);
  int32_t* col_best = a.take<int32_t>(cols + 1);
  coarse_match_kernel<<<1, 1024, 0, stream>>>(log_scores, rows, cols, out_i, out_j, out_scores, out_count, row_best,
                                              col_best, row_cnt);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_patch_scores(const float* feats_a, int64_t n_a, const int32_t* knn_a, const int32_t* node_a,
                                const float* feats_b, int64_t n_b, const int32_t* knn_b, const int32_t* node_b,
                                int n_pairs, int k, int channels, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(k == PK && channels == PC, "patch_scores: specialised to 128 points x 128 channels");
  if (n_pairs == 0) return LCR_OK;
  LcrProfScope prof("patch_scores", 2.0 * n_pairs * PK * PK * PC, 4.0 * n_pairs * (2.0 * PK * PC + PK * PK), stream);
  static const bool simt = getenv("LCR_PATCH") && !strcmp(getenv("LCR_PATCH"), "simt");   // default: mma.sync 3xTF32
  if (simt) {
    const size_t smem = sizeof(float) * 2 * PCH * (PK + 4);
    LCR_CUDA_TRY(cudaFuncSetAttribute(patch_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patch_scores_kernel<<<n_pairs, 256, smem, stream>>>(feats_a, (int)n_a, knn_a, node_a, feats_b, (int)n_b, knn_b,
                                                        node_b, sqrtf((float)PC), out);
  } else {
    const size_t smem = sizeof(float) * 2 * PK * PSS;
    LCR_CUDA_TRY(cudaFuncSetAttribute(patch_scores_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patch_scores_mma_kernel<<<n_pairs, 256, smem, stream>>>(feats_a, (int)n_a, knn_a, node_a, feats_b, (int)n_b, knn_b,
                                                            node_b, sqrtf((float)PC), out);
  }
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

// Fine correspondences of all node pairs, row-major per pair (the reference's torch.nonzero order).
// pair_off[n_pairs+1] (device) receives the exclusive scan of the per-pair counts; outputs have
// capacity n_pairs * 256.
extern "C" size_t lcr_fine_correspondences_ws_bytes(int n_pairs) {
  return lcr_align_up((size_t)n_pairs * 256 * 4) * 2 + 256;
}

extern "C" int lcr_fine_correspondences(const float* log_scores, int n_pairs, const uint8_t* knn_mask_a,
                                        const int32_t* node_a, const uint8_t* knn_mask_b, const int32_t* node_b,
                                        int32_t* pair_cnt, int32_t* pair_off, int32_t* out_pair, int32_t* out_i,
                                        int32_t* out_j, float* out_scores, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_pairs == 0) {
    LCR_CUDA_TRY(cudaMemsetAsync(pair_off, 0, sizeof(int32_t), stream));
    return LCR_OK;
  }
  LCR_REQUIRE(ws && ws_bytes >= lcr_fine_correspondences_ws_bytes(n_pairs), "fine_correspondences: workspace too small");
  LcrArena a(ws, ws_bytes);
  int32_t* slot_ij = a.take<int32_t>((size_t)n_pairs * 256);
  float* slot_s = a.take<float>((size_t)n_pairs * 256);
  static LcrOncePerDevice once;
  const int dev = once.need();
  if (dev != -1) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(fine_corr_find_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kFineSmem));
    once.done(dev);
  }
  LcrProfScope prof("fine_correspondences", 0.0, 4.0 * n_pairs * 129.0 * 129.0, stream);
  fine_corr_find_kernel<<<n_pairs, 256, kFineSmem, stream>>>(log_scores, knn_mask_a, node_a, knn_mask_b, node_b,
                                                             pair_cnt, slot_ij, slot_s);
  exclusive_scan_i32_kernel<<<1, 1024, 0, stream>>>(pair_cnt, n_pairs, pair_off);
  fine_corr_compact_kernel<<<(unsigned)n_pairs, 256, 0, stream>>>(slot_ij, slot_s, pair_cnt, pair_off, n_pairs, out_pair,
                                                                  out_i, out_j, out_scores);
  LCR_LAUNCHED(3);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_corr_points(const int32_t* c_pair, const int32_t* c_i, const int32_t* c_j, const int32_t* n_corr,
                               int64_t capacity, const float* pts_a, const int32_t* knn_a, const int32_t* node_a,
                               const float* pts_b, const int32_t* knn_b, const int32_t* node_b, float* ref, float* src,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (capacity == 0) return LCR_OK;
  corr_points_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, stream>>>(c_pair, c_i, c_j, n_corr, pts_a, knn_a,
                                                                            node_a, pts_b, knn_b, node_b, ref, src);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" size_t lcr_lgr_ws_bytes(int n_pairs, int64_t capacity) {
  return lcr_align_up((size_t)(n_pairs + 1) * 64) + lcr_align_up((size_t)(n_pairs + 1) * 4) * 2 +
         lcr_align_up((size_t)capacity * 4) + 1024;
}

// Local-to-global registration on the device.  ref/src/scores [n_corr] grouped by pair
// (pair_off[n_pairs+1]); n_corr (device) = pair_off[n_pairs].  out_T [16] row-major 4x4.
extern "C" int lcr_local_global_registration(const float* ref, const float* src, const float* scores,
                                             const int32_t* pair_off, int n_pairs, int64_t capacity, float radius,
                                             int min_corr, int steps, float* out_T, void* ws, size_t ws_bytes,
                                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n_pairs >= 0 && capacity >= 0 && steps >= 1, "lgr: sizes");
  LCR_REQUIRE(ws && ws_bytes >= lcr_lgr_ws_bytes(n_pairs, capacity), "lgr: workspace too small");
  LcrArena a(ws, ws_bytes);
  float* T_local = a.take<float>((size_t)(n_pairs + 1) * 16);
  int32_t* valid = a.take<int32_t>(n_pairs + 1);
  int32_t* counts = a.take<int32_t>(n_pairs + 1);
  float* w_cur = a.take<float>(capacity > 0 ? capacity : 1);
  int32_t* have_local = (int32_t*)a.take<int32_t>(1);
  const int32_t* n_corr = pair_off + n_pairs;
  const unsigned gridC = (unsigned)((capacity + 255) / 256) > 0 ? (unsigned)((capacity + 255) / 256) : 1u;
  LcrProfScope prof("lgr", 0.0, 28.0 * capacity * (n_pairs + 2.0 * steps), stream);
  LCR_CUDA_TRY(cudaMemsetAsync(have_local, 0, sizeof(int32_t), stream));
  int launches = 0;
  if (n_pairs > 0) {
    procrustes_kernel<<<n_pairs, 256, 0, stream>>>(src, ref, scores, pair_off, n_corr, min_corr, T_local, valid);
    inlier_count_kernel<<<n_pairs, 256, 0, stream>>>(T_local, valid, src, ref, n_corr, radius, counts, nullptr, nullptr);
    pick_best_kernel<<<1, 1024, 0, stream>>>(counts, n_pairs, T_local, out_T, have_local, nullptr);
    launches += 3;
  }
  // degenerate branch (no patch with >= min_corr correspondences): start from all correspondences
  // (local_global_registration.py:183-188); computed unconditionally, selected on the device
  float* T_deg = T_local + (size_t)n_pairs * 16;
  procrustes_kernel<<<1, 256, 0, stream>>>(src, ref, scores, nullptr, n_corr, 0, T_deg, nullptr);
  // w = score * inlier(have_local ? T_best : T_deg)
  reweight_kernel<<<gridC, 256, 0, stream>>>(out_T, T_deg, have_local, src, ref, scores, n_corr, radius, w_cur, nullptr,
                                             nullptr);
  launches += 2;
  for (int it = 0; it < steps; it++) {
    procrustes_kernel<<<1, 256, 0, stream>>>(src, ref, w_cur, nullptr, n_corr, 0, out_T, nullptr);
    launches++;
    if (it + 1 < steps) {
      reweight_kernel<<<gridC, 256, 0, stream>>>(out_T, nullptr, nullptr, src, ref, scores, n_corr, radius, w_cur,
                                                 nullptr, nullptr);
      launches++;
    }
  }
  LCR_LAUNCHED(launches);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

// ================================================================== batched C ABI (all pairs of a chunk per call)
extern "C" int lcr_point_to_node_batched(const float* points, int64_t n_points, const int64_t* pts_off,
                                         const float* nodes, int64_t n_nodes, const int64_t* node_off, int n_clouds,
                                         int64_t max_cloud_points, int64_t max_cloud_nodes, int k,
                                         int32_t* point_to_node, uint8_t* node_mask, int32_t* knn_global,
                                         int64_t* knn_local, uint8_t* knn_mask, int32_t* out_status, void* ws,
                                         size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n_points >= 1 && n_nodes >= 1 && n_points < (1ll << 31) && n_nodes < (1ll << 31) && n_clouds >= 1,
              "point_to_node_batched: sizes");
  LCR_REQUIRE(max_cloud_nodes >= 1 && max_cloud_nodes <= 12000, "point_to_node_batched: at most 12000 nodes per cloud");
  LCR_REQUIRE(k >= 1 && k <= kNodeCap, "point_to_node_batched: k");
  LCR_REQUIRE(points && pts_off && nodes && node_off && node_mask && knn_global && knn_mask, "point_to_node_batched: null");
  LCR_REQUIRE(ws && ws_bytes >= lcr_point_to_node_ws_bytes(n_points, n_nodes), "point_to_node_batched: workspace too small");
  const int N = (int)n_points, M = (int)n_nodes;
  LcrArena a(ws, ws_bytes);
  float* d2 = a.take<float>(N);
  int32_t* owner_tmp = a.take<int32_t>(N);
  unsigned long long* keys = a.take<unsigned long long>(N);
  uint32_t* count = a.take<uint32_t>(M + 1);
  uint32_t* start = a.take<uint32_t>(M + 1);
  uint32_t* cursor = a.take<uint32_t>(M + 1);
  int* err = (int*)a.take<int>(1);
  int32_t* owner = point_to_node ? point_to_node : owner_tmp;
  LcrProfScope prof("point_to_node", 8.0 * N * (double)max_cloud_nodes, 12.0 * (N + M) + 4.0 * N + 13.0 * M * k, stream);
  LCR_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(uint32_t) * (M + 1), stream));
  LCR_CUDA_TRY(cudaMemsetAsync(cursor, 0, sizeof(uint32_t) * (M + 1), stream));
  if (out_status) LCR_CUDA_TRY(cudaMemsetAsync(out_status, 0, sizeof(int32_t), stream));
  const size_t smem = sizeof(float4) * (size_t)max_cloud_nodes;
  if (smem > 48 * 1024)
    LCR_CUDA_TRY(cudaFuncSetAttribute(owner_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((max_cloud_points + 255) / 256), (unsigned)n_clouds);
  owner_batched_kernel<<<grid, 256, smem, stream>>>(points, pts_off, nodes, node_off, owner, d2, count);
  node_scan_kernel<<<1, 1024, 0, stream>>>(count, M, start, node_mask);
  node_scatter_kernel<<<(N + 255) / 256, 256, 0, stream>>>(owner, d2, N, start, cursor, keys);
  node_topk_batched_kernel<<<M, 256, 0, stream>>>(keys, start, N, k, pts_off, node_off, n_clouds, knn_global, knn_local,
                                                  knn_mask, out_status ? out_status : err);
  LCR_LAUNCHED(4);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_node_scores(const float* feats, int channels, const int64_t* node_off, const uint8_t* node_mask,
                               int n_pairs, int m_max, int n_max, float* out, uint8_t* row_mask, uint8_t* col_mask,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n_pairs >= 1 && m_max >= 1 && n_max >= 1 && channels % NS_K == 0, "node_scores: sizes");
  LCR_REQUIRE(feats && node_off && node_mask && out && row_mask && col_mask, "node_scores: null");
  LcrProfScope prof("node_scores", 2.0 * n_pairs * (double)m_max * n_max * channels,
                    4.0 * n_pairs * ((double)(m_max + n_max) * channels + (double)m_max * n_max), stream);
  dim3 grid((unsigned)((n_max + NS_T - 1) / NS_T), (unsigned)((m_max + NS_T - 1) / NS_T), (unsigned)n_pairs);
  node_scores_kernel<<<grid, 256, 0, stream>>>(feats, channels, node_off, node_mask, m_max, n_max,
                                               1.f / sqrtf((float)channels), out, row_mask, col_mask);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_coarse_matching_batched(const float* log_scores, int n_pairs, int rows, int cols, int32_t* out_i,
                                           int32_t* out_j, float* out_scores, int32_t* out_count, void* ws,
                                           size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n_pairs >= 1 && rows >= 1 && cols >= 1, "coarse_matching_batched: sizes");
  const size_t per = (size_t)(rows + 1) * 2 + (size_t)(cols + 1);
  LCR_REQUIRE(ws && ws_bytes >= sizeof(int32_t) * per * n_pairs + 1024, "coarse_matching_batched: workspace too small");
  LcrArena a(ws, ws_bytes);
  int32_t* row_best = a.take<int32_t>((size_t)(rows + 1) * n_pairs);
  int32_t* row_cnt = a.take<int32_t>((size_t)(rows + 1) * n_pairs);
  int32_t* col_best = a.take<int32_t>((size_t)(cols + 1) * n_pairs);
  LcrProfScope prof("coarse_matching", 0.0, 12.0 * n_pairs * (double)(rows + 1) * (cols + 1), stream);
  coarse_match_kernel<<<n_pairs, 1024, 0, stream>>>(log_scores, rows, cols, out_i, out_j, out_scores, out_count, row_best,
                                                    col_best, row_cnt);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_gather_coarse(const int32_t* out_i, const int32_t* out_j, const float* out_scores, int capacity,
                                 const int32_t* patch_off, const int64_t* node_off, int n_pairs, int32_t* ci_global,
                                 int32_t* cj_global, int32_t* ci_local, int32_t* cj_local, float* scores,
                                 int32_t* patch_pair, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n_pairs >= 1 && capacity >= 1, "gather_coarse: sizes");
  gather_coarse_kernel<<<n_pairs, 256, 0, stream>>>(out_i, out_j, out_scores, capacity, patch_off, node_off, ci_global,
                                                    cj_global, ci_local, cj_local, scores, patch_pair);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" size_t lcr_lgr_batched_ws_bytes(int n_patches, int n_scan_pairs, int64_t capacity) {
  return lcr_align_up((size_t)(n_patches + 1) * 16 * 4) + 2 * lcr_align_up((size_t)(n_patches + 1) * 4) +
         lcr_align_up((size_t)(capacity > 0 ? capacity : 1) * 4) + 2 * lcr_align_up((size_t)(n_scan_pairs + 1) * 4) +
         lcr_align_up((size_t)(n_scan_pairs + 1) * 16 * 4) + 1024;
}

// Local-to-global registration (local_global_registration.py:140-202) of MANY scan pairs in one set of launches.
// ref / src / scores [capacity]: correspondences of all scan pairs, grouped by patch (node correspondence):
// patch t owns [pair_off[t], pair_off[t+1]); c_pair[c] = patch of correspondence c; scan pair s owns the patches
// [patch_off[s], patch_off[s+1]) (patch_pair[t] = s).  out_T [n_scan_pairs, 4, 4].
extern "C" int lcr_lgr_batched(const float* ref, const float* src, const float* scores, const int32_t* c_pair,
                               const int32_t* pair_off, int n_patches, const int32_t* patch_pair,
                               const int32_t* patch_off, int n_scan_pairs, int64_t capacity, float radius, int min_corr,
                               int steps, float* out_T, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n_patches >= 0 && n_scan_pairs >= 1 && capacity >= 0 && steps >= 1, "lgr_batched: sizes");
  LCR_REQUIRE(ws && ws_bytes >= lcr_lgr_batched_ws_bytes(n_patches, n_scan_pairs, capacity), "lgr_batched: workspace too small");
  LcrArena a(ws, ws_bytes);
  float* T_local = a.take<float>((size_t)(n_patches + 1) * 16);
  int32_t* valid = a.take<int32_t>(n_patches + 1);
  int32_t* counts = a.take<int32_t>(n_patches + 1);
  float* w_cur = a.take<float>(capacity > 0 ? capacity : 1);
  int32_t* have_local = a.take<int32_t>(n_scan_pairs + 1);
  int32_t* corr_off = a.take<int32_t>(n_scan_pairs + 1);
  float* T_deg = a.take<float>((size_t)(n_scan_pairs + 1) * 16);
  const int32_t* n_corr = pair_off + n_patches;
  const unsigned gridC = (unsigned)((capacity + 255) / 256) > 0 ? (unsigned)((capacity + 255) / 256) : 1u;
  const int S = n_scan_pairs;
  LcrProfScope prof("lgr", 0.0, 28.0 * capacity * (n_patches / (double)S + 2.0 * steps), stream);
  LCR_CUDA_TRY(cudaMemsetAsync(have_local, 0, sizeof(int32_t) * (S + 1), stream));
  int launches = 1;
  corr_offsets_kernel<<<(S + 256) / 256, 256, 0, stream>>>(pair_off, patch_off, S, corr_off);
  if (n_patches > 0) {
    procrustes_kernel<<<n_patches, 256, 0, stream>>>(src, ref, scores, pair_off, n_corr, min_corr, T_local, valid);
    inlier_count_kernel<<<n_patches, 256, 0, stream>>>(T_local, valid, src, ref, n_corr, radius, counts, patch_pair,
                                                       corr_off);
    launches += 2;
  }
  pick_best_kernel<<<S, 1024, 0, stream>>>(counts, n_patches, T_local, out_T, have_local, patch_off);
  // degenerate branch per scan pair (no patch with >= min_corr correspondences): all its correspondences
  procrustes_kernel<<<S, 256, 0, stream>>>(src, ref, scores, corr_off, n_corr, 0, T_deg, nullptr);
  reweight_kernel<<<gridC, 256, 0, stream>>>(out_T, T_deg, have_local, src, ref, scores, n_corr, radius, w_cur, c_pair,
                                             patch_pair);
  launches += 3;
  for (int it = 0; it < steps; it++) {
    procrustes_kernel<<<S, 256, 0, stream>>>(src, ref, w_cur, corr_off, n_corr, 0, out_T, nullptr);
    launches++;
    if (it + 1 < steps) {
      reweight_kernel<<<gridC, 256, 0, stream>>>(out_T, nullptr, nullptr, src, ref, scores, n_corr, radius, w_cur, c_pair,
                                                 patch_pair);
      launches++;
    }
  }
  LCR_LAUNCHED(launches);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
