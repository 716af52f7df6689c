// ransac.cu -- SURVEY 8(f) row 4: RANSAC rigid registration from putative correspondences, the alternative
// estimator of the reference's evaluation (utils/utils/open3d.py:145-173 wraps open3d's
// registration_ransac_based_on_correspondence with TransformationEstimationPointToPoint(False), ransac_n = 3,
// distance_threshold = 0.05, num_iterations hypotheses).  open3d is a third-party dependency that is neither in the
// reference tree nor pinned by it: PARITY UNPINNED at that boundary.  This is the published algorithm -- per
// hypothesis: sample ransac_n correspondences, least-squares rigid transform of the sample (Kabsch), fitness =
// number of correspondences with |T src - ref| < threshold, ties broken by the lower inlier RMSE (open3d's
// IsBetterRANSACThan), best hypothesis wins -- with a counter-based sampler so that every hypothesis is reproducible
// (oracle/ransac_oracle.py evaluates the same samples on the CPU).  All hypotheses run in parallel, one CTA each.
#include "common.cuh"
#include "kabsch.cuh"

namespace {

constexpr int kMaxSample = 8;

// sample index of draw k of hypothesis h: splitmix-style hash of (seed, h, k) reduced to [0, n)
__device__ __forceinline__ uint32_t ransac_draw(uint64_t seed, uint32_t h, uint32_t k, uint32_t n) {
  const uint64_t x = lcr_mix64(seed ^ (((uint64_t)h << 8) | (uint64_t)k) * 0x9E3779B97F4A7C15ull);
  return (uint32_t)((x >> 11) % (uint64_t)n);
}

// rigid transform (row-major 3x4 into T[12]) mapping the sampled src points onto the sampled ref points
__device__ void sample_transform(const float* __restrict__ src, const float* __restrict__ ref, const uint32_t* idx,
                                 int ns, float* T) {
  double sc[3] = {0, 0, 0}, rc[3] = {0, 0, 0};
  for (int s = 0; s < ns; s++)
    for (int d = 0; d < 3; d++) {
      sc[d] += (double)src[3 * (size_t)idx[s] + d];
      rc[d] += (double)ref[3 * (size_t)idx[s] + d];
    }
  for (int d = 0; d < 3; d++) {
    sc[d] /= ns;
    rc[d] /= ns;
  }
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, R[3][3];
  for (int s = 0; s < ns; s++) {
    double a[3], b[3];
    for (int d = 0; d < 3; d++) {
      a[d] = (double)src[3 * (size_t)idx[s] + d] - sc[d];
      b[d] = (double)ref[3 * (size_t)idx[s] + d] - rc[d];
    }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) H[i][j] += a[i] * b[j];
  }
  kabsch_rotation(H, R);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[4 * i + j] = (float)R[i][j];
    T[4 * i + 3] = (float)(rc[i] - (R[i][0] * sc[0] + R[i][1] * sc[1] + R[i][2] * sc[2]));
  }
}

__global__ void __launch_bounds__(128)
ransac_hypotheses_kernel(const float* __restrict__ src, const float* __restrict__ ref, int n, int ns, float thr,
                         uint64_t seed, float* __restrict__ T_all /*[H,12]*/, int32_t* __restrict__ count,
                         float* __restrict__ sqsum) {
  __shared__ float sT[12];
  __shared__ int s_c[4];
  __shared__ float s_e[4];
  const int h = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    uint32_t idx[kMaxSample];
    for (int k = 0; k < ns; k++) idx[k] = ransac_draw(seed, (uint32_t)h, (uint32_t)k, (uint32_t)n);
    sample_transform(src, ref, idx, ns, sT);
    for (int k = 0; k < 12; k++) T_all[(size_t)h * 12 + k] = sT[k];
  }
  __syncthreads();
  int c = 0;
  float e = 0.f;
  const float thr2 = thr * thr;
  for (int t = tid; t < n; t += 128) {
    const float x = src[3 * (size_t)t], y = src[3 * (size_t)t + 1], z = src[3 * (size_t)t + 2];
    const float dx = ref[3 * (size_t)t] - (sT[0] * x + sT[1] * y + sT[2] * z + sT[3]);
    const float dy = ref[3 * (size_t)t + 1] - (sT[4] * x + sT[5] * y + sT[6] * z + sT[7]);
    const float dz = ref[3 * (size_t)t + 2] - (sT[8] * x + sT[9] * y + sT[10] * z + sT[11]);
    const float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 < thr2) {
      c++;
      e += d2;
    }
  }
  c = lcr_warp_sum(c);
  e = lcr_warp_sum(e);
  if ((tid & 31) == 0) {
    s_c[tid >> 5] = c;
    s_e[tid >> 5] = e;
  }
  __syncthreads();
  if (tid == 0) {
    count[h] = s_c[0] + s_c[1] + s_c[2] + s_c[3];
    sqsum[h] = (s_e[0] + s_e[1]) + (s_e[2] + s_e[3]);
  }
}

// best = most inliers; ties: lower inlier RMSE (= lower squared sum at equal count), then the lower hypothesis index
__global__ void __launch_bounds__(1024)
ransac_select_kernel(const int32_t* __restrict__ count, const float* __restrict__ sqsum, int H,
                     const float* __restrict__ T_all, float* __restrict__ out_T, int32_t* __restrict__ out_best) {
  __shared__ int s_c[32], s_i[32];
  __shared__ float s_e[32];
  int bc = -1, bi = 0x7fffffff;
  float be = INFINITY;
  auto better = [](int c, float e, int i, int c2, float e2, int i2) {
    return c > c2 || (c == c2 && (e < e2 || (e == e2 && i < i2)));
  };
  for (int h = threadIdx.x; h < H; h += blockDim.x)
    if (better(count[h], sqsum[h], h, bc, be, bi)) { bc = count[h]; be = sqsum[h]; bi = h; }
  for (int o = 16; o > 0; o >>= 1) {
    const int oc = __shfl_xor_sync(0xffffffffu, bc, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const float oe = __shfl_xor_sync(0xffffffffu, be, o);
    if (better(oc, oe, oi, bc, be, bi)) { bc = oc; be = oe; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = bc; s_e[threadIdx.x >> 5] = be; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++)
      if (better(s_c[w], s_e[w], s_i[w], bc, be, bi)) { bc = s_c[w]; be = s_e[w]; bi = s_i[w]; }
    for (int i = 0; i < 12; i++) out_T[i] = T_all[(size_t)bi * 12 + i];
    out_T[12] = out_T[13] = out_T[14] = 0.f;
    out_T[15] = 1.f;
    out_best[0] = bi;
    out_best[1] = bc;
  }
}

}  // namespace

extern "C" size_t lcr_ransac_ws_bytes(int num_iterations) {
  return lcr_align_up((size_t)num_iterations * 12 * 4) + 2 * lcr_align_up((size_t)num_iterations * 4) + 256;
}

extern "C" int lcr_ransac_correspondences(const float* src, const float* ref, int64_t n, float distance_threshold,
                                          int ransac_n, int num_iterations, uint64_t seed, float* out_T,
                                          int32_t* out_best, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n >= 1 && n < (1ll << 31) && num_iterations >= 1 && ransac_n >= 3 && ransac_n <= kMaxSample,
              "ransac: sizes (3 <= ransac_n <= 8)");
  LCR_REQUIRE(src && ref && out_T && out_best && distance_threshold > 0.f, "ransac: null / threshold");
  LCR_REQUIRE(ws && ws_bytes >= lcr_ransac_ws_bytes(num_iterations), "ransac: workspace too small");
  LcrArena a(ws, ws_bytes);
  float* T_all = a.take<float>((size_t)num_iterations * 12);
  int32_t* count = a.take<int32_t>(num_iterations);
  float* sqsum = a.take<float>(num_iterations);
  LcrProfScope prof("ransac", 20.0 * n * num_iterations, 24.0 * n, stream);
  ransac_hypotheses_kernel<<<num_iterations, 128, 0, stream>>>(src, ref, (int)n, ransac_n, distance_threshold, seed, T_all,
                                                               count, sqsum);
  ransac_select_kernel<<<1, 1024, 0, stream>>>(count, sqsum, num_iterations, T_all, out_T, out_best);
  LCR_LAUNCHED(2);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
