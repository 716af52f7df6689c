// pair.cu -- a7/a9 and the decoder gather of the registration path:
//   LayerNorm(+residual,+ReLU)          vanilla_transformer.py:22-28, rpetransformer.py:139-142, vote.py:124-128
//   rotary position embedding           thdroformer/rpetransformer.py:41-54
//   multi-head softmax attention        rpetransformer.py:19-24, vanilla_transformer.py:58-72
//   vote offset clamp                   modules/vote/vote.py:166-172
//   greedy radius NMS                   modules/vote/vote.py:13-70
//   node centre = mean of neighbours    backbone4.py:159-176
//   nearest-upsample + concat           modules/kpconv/functional.py:6-22, backbone4.py:355-368
// All fp32.  The attention kernel is a flash-style tiled kernel batched over (problem, head): a
// problem is one side of one scan pair (queries and keys given by row offsets), so the 32 pairs
// x 2 sides x 4 heads of BASELINE config 3 run as one launch.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ LayerNorm
// y = act(LN(x + res) * gamma + beta); one warp per row, C in {128, 256, 512, 1024}
__global__ void __launch_bounds__(256)
layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
                  const float* __restrict__ beta, int64_t rows, int C, float eps, int relu, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  float v[32];  // C / 32 <= 32
  const int per = C / 32;
  float sum = 0.f;
  for (int i = 0; i < per; i++) {
    const int c = lane + 32 * i;
    float t = x[r * C + c];
    if (res) t += res[r * C + c];
    v[i] = t;
    sum += t;
  }
  const float mean = lcr_warp_sum(sum) / (float)C;
  float sq = 0.f;
  for (int i = 0; i < per; i++) {
    const float d = v[i] - mean;
    sq = fmaf(d, d, sq);
  }
  const float rstd = 1.0f / sqrtf(lcr_warp_sum(sq) / (float)C + eps);
  for (int i = 0; i < per; i++) {
    const int c = lane + 32 * i;
    float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
    if (relu) o = fmaxf(o, 0.f);
    y[r * C + c] = o;
  }
}

// ------------------------------------------------------------------ rotary embedding (in place)
// x [rows, ld] (first 128 columns = 4 heads x 32), theta [rows, 64]: pair (2j, 2j+1) of head h is
// rotated by theta[16h + j].
__global__ void rope_kernel(float* __restrict__ x, int ld, const float* __restrict__ theta, int64_t rows) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * 64) return;
  const int64_t r = t >> 6;
  const int j = (int)(t & 63);
  float s, c;
  sincosf(theta[r * 64 + j], &s, &c);
  float2* p = reinterpret_cast<float2*>(x + r * ld) + j;
  const float2 v = *p;
  *p = make_float2(v.x * c - v.y * s, v.y * c + v.x * s);
}

// ------------------------------------------------------------------ attention
// grid (q tiles, problems * heads); 256 threads; 64 queries x 64 keys per inner tile, d = 32.
constexpr int AQ = 64, AK = 64, AD = 32;

__global__ void __launch_bounds__(256)
attention_kernel(const float* __restrict__ q, int ld_q, const float* __restrict__ k, int ld_k,
                 const float* __restrict__ v, int ld_v, const int64_t* __restrict__ q_off,
                 const int64_t* __restrict__ k_off, int heads, float inv_scale_div, float* __restrict__ out,
                 int ld_o) {
  __shared__ float s_q[AD][AQ + 1];   // transposed: [c][query]
  __shared__ float s_k[AD][AK + 1];   // transposed: [c][key]
  __shared__ float s_v[AK][AD];       // [key][c]
  __shared__ float s_p[AQ][AK + 1];   // probabilities
  const int prob = blockIdx.y / heads, head = blockIdx.y % heads;
  const int64_t q0 = q_off[prob] + (int64_t)blockIdx.x * AQ, q1 = q_off[prob + 1];
  if (q0 >= q1) return;
  const int64_t k0 = k_off[prob], k1 = k_off[prob + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // score micro-tile: rows ty*4.., cols tx*4..
  for (int i = tid; i < AQ * AD; i += 256) {
    const int r = i / AD, c = i % AD;
    s_q[c][r] = q0 + r < q1 ? q[(q0 + r) * ld_q + head * AD + c] : 0.f;
  }
  float m_run[4], l_run[4], acc[4][2];   // output micro-tile: rows ty*4.., cols tx*2..
#pragma unroll
  for (int i = 0; i < 4; i++) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
    acc[i][0] = acc[i][1] = 0.f;
  }
  for (int64_t kb = k0; kb < k1; kb += AK) {
    __syncthreads();
    for (int i = tid; i < AK * AD; i += 256) {
      const int r = i / AD, c = i % AD;
      const bool ok = kb + r < k1;
      s_k[c][r] = ok ? k[(kb + r) * ld_k + head * AD + c] : 0.f;
      s_v[r][c] = ok ? v[(kb + r) * ld_v + head * AD + c] : 0.f;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) s[i][j] = 0.f;
#pragma unroll
    for (int c = 0; c < AD; c++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = s_q[c][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = s_k[c][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s[i][j] = fmaf(a[i], b[j], s[i][j]);
    }
    // online softmax: the 16 threads sharing ty (one half-warp) own the same 4 rows
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        s[i][j] = (kb + tx * 4 + j < k1) ? s[i][j] / inv_scale_div : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m_run[i], mx);
      const float corr = expf(m_run[i] - m_new);
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float p = expf(s[i][j] - m_new);
        s_p[ty * 4 + i][tx * 4 + j] = p;
        ps += p;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      l_run[i] = l_run[i] * corr + ps;
      m_run[i] = m_new;
      acc[i][0] *= corr;
      acc[i][1] *= corr;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < AK; kk++) {
      const float2 vv = *reinterpret_cast<const float2*>(&s_v[kk][tx * 2]);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float p = s_p[ty * 4 + i][kk];
        acc[i][0] = fmaf(p, vv.x, acc[i][0]);
        acc[i][1] = fmaf(p, vv.y, acc[i][1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int64_t r = q0 + ty * 4 + i;
    if (r < q1) {
      float* o = out + r * ld_o + head * AD + tx * 2;
      o[0] = acc[i][0] / l_run[i];
      o[1] = acc[i][1] / l_run[i];
    }
  }
}

// ------------------------------------------------------------------ vote shift
__global__ void vote_shift_kernel(const float* __restrict__ pts, const float* __restrict__ off, int ld_off,
                                  float max_range, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ox = off[i * ld_off], oy = off[i * ld_off + 1], oz = off[i * ld_off + 2];
  const float dis = sqrtf(ox * ox + oy * oy + oz * oz);
  const float alpha = dis > max_range ? max_range / dis : 1.0f;
  out[3 * i] = pts[3 * i] + ox * alpha;
  out[3 * i + 1] = pts[3 * i + 1] + oy * alpha;
  out[3 * i + 2] = pts[3 * i + 2] + oz * alpha;
}

// ------------------------------------------------------------------ greedy NMS: one CTA per cloud
constexpr int kNmsThreads = 256;
__global__ void __launch_bounds__(kNmsThreads)
nms_kernel(const float* __restrict__ pts, const int64_t* __restrict__ off, float radius, uint8_t* __restrict__ keep,
           int32_t* __restrict__ counts, int32_t* __restrict__ kept_idx /* per point slot: compacted kept ids */) {
  extern __shared__ float s_kept[];  // 3 floats per kept point, capacity = cloud size
  const int b = blockIdx.x;
  const int64_t p0 = off[b];
  const int n = (int)(off[b + 1] - p0);
  int cnt = 0;
  for (int i = 0; i < n; i++) {
    const float x = pts[3 * (p0 + i)], y = pts[3 * (p0 + i) + 1], z = pts[3 * (p0 + i) + 2];
    int ok = 1;
    // nn.PairwiseDistance(p=2): || x1 - x2 + eps ||_2 with eps = 1e-6; kept iff ALL distances > radius
    for (int j = threadIdx.x; j < cnt; j += kNmsThreads) {
      const float dx = (x - s_kept[3 * j]) + 1e-6f, dy = (y - s_kept[3 * j + 1]) + 1e-6f,
                  dz = (z - s_kept[3 * j + 2]) + 1e-6f;
      const float d = sqrtf(dx * dx + dy * dy + dz * dz);
      if (!(d > radius)) ok = 0;
    }
    ok = __syncthreads_and(ok);
    if (i == 0) ok = 1;  // the first point is always kept (vote.py:39)
    if (ok) {
      if (threadIdx.x == 0) {
        s_kept[3 * cnt] = x;
        s_kept[3 * cnt + 1] = y;
        s_kept[3 * cnt + 2] = z;
        kept_idx[p0 + cnt] = (int32_t)(p0 + i);
      }
      cnt++;
    }
    if (threadIdx.x == 0) keep[p0 + i] = (uint8_t)ok;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = cnt;
}

// ------------------------------------------------------------------ mean of valid neighbours
__global__ void neighbor_mean_kernel(const float* __restrict__ pts, const int32_t* __restrict__ idx, int ld_idx, int H,
                                     int N, int M, float* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  int c = 0;
  for (int h = 0; h < H; h++) {
    const int j = idx[(size_t)m * ld_idx + h];
    if (j < N) {
      sx += pts[3 * (size_t)j];
      sy += pts[3 * (size_t)j + 1];
      sz += pts[3 * (size_t)j + 2];
      c++;
    }
  }
  const float d = (float)c;
  out[3 * m] = sx / d;
  out[3 * m + 1] = sy / d;
  out[3 * m + 2] = sz / d;
}

// ------------------------------------------------------------------ nearest-upsample + concat
// out[i, :C1] = coarse[up[i, 0]] (zeros if pad), out[i, C1:] = fine[i]
__global__ void upsample_concat_kernel(const float* __restrict__ coarse, int n_coarse, int C1,
                                       const int32_t* __restrict__ up, int ld_up, const float* __restrict__ fine,
                                       int C2, int64_t n_fine, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_fine) return;
  const int j = up[i * ld_up];
  float* o = out + i * (C1 + C2);
  for (int c = lane * 4; c < C1; c += 128)
    *reinterpret_cast<float4*>(o + c) =
        j < n_coarse ? *reinterpret_cast<const float4*>(coarse + (size_t)j * C1 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = lane * 4; c < C2; c += 128)
    *reinterpret_cast<float4*>(o + C1 + c) = *reinterpret_cast<const float4*>(fine + i * C2 + c);
}

}  // namespace

extern "C" int lcr_layer_norm(const float* x, const float* residual, const float* gamma, const float* beta,
                              int64_t rows, int channels, float eps, int relu, float* y, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(channels % 32 == 0 && channels <= 1024, "layer_norm: channels must be a multiple of 32, <= 1024");
  if (rows == 0) return LCR_OK;
  LcrProfScope prof("layer_norm", 8.0 * rows * channels, 4.0 * rows * channels * (residual ? 3.0 : 2.0), stream);
  layer_norm_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, stream>>>(x, residual, gamma, beta, rows, channels,
                                                                            eps, relu, y);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_rope(float* x, int ld, const float* theta, int64_t rows, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(ld >= 128 && ld % 2 == 0, "rope: rows must hold 4 heads x 32 channels");
  if (rows == 0) return LCR_OK;
  LcrProfScope prof("rope", 6.0 * rows * 64, 4.0 * rows * (128 * 2 + 64), stream);
  rope_kernel<<<(unsigned)((rows * 64 + 255) / 256), 256, 0, stream>>>(x, ld, theta, rows);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_attention(const float* q, int ld_q, const float* k, int ld_k, const float* v, int ld_v,
                             const int64_t* q_off, const int64_t* k_off, int n_problems, int64_t max_q_rows,
                             int heads, int head_dim, float* out, int ld_out, double flops_hint, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(head_dim == AD, "attention: head_dim must be 32");
  LCR_REQUIRE(n_problems >= 1 && heads >= 1 && max_q_rows >= 0, "attention: bad sizes");
  if (max_q_rows == 0) return LCR_OK;
  LcrProfScope prof("attention", flops_hint, 0.0, stream);
  dim3 grid((unsigned)((max_q_rows + AQ - 1) / AQ), (unsigned)(n_problems * heads));
  attention_kernel<<<grid, 256, 0, stream>>>(q, ld_q, k, ld_k, v, ld_v, q_off, k_off, heads, sqrtf((float)head_dim),
                                             out, ld_out);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_vote_shift(const float* points, const float* offsets, int ld_offsets, float max_range, int64_t n,
                              float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n == 0) return LCR_OK;
  vote_shift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(points, offsets, ld_offsets, max_range, n, out);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_nms_greedy(const float* points, const int64_t* cloud_off, int n_clouds, int64_t max_cloud_rows,
                              float radius, uint8_t* keep, int32_t* counts, int32_t* kept_idx, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n_clouds >= 1, "nms: n_clouds");
  const size_t smem = (size_t)max_cloud_rows * 3 * sizeof(float);
  LCR_REQUIRE(smem <= 200 * 1024, "nms: cloud too large for the shared-memory kept list (17066 points)");
  if (smem > 48 * 1024)
    LCR_CUDA_TRY(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LcrProfScope prof("nms_greedy", 0.0, 12.0 * max_cloud_rows * n_clouds, stream);
  nms_kernel<<<n_clouds, kNmsThreads, smem, stream>>>(points, cloud_off, radius, keep, counts, kept_idx);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_neighbor_mean(const float* points, int64_t n_points, const int32_t* idx, int ld_idx, int H,
                                 int64_t m_rows, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (m_rows == 0) return LCR_OK;
  neighbor_mean_kernel<<<(unsigned)((m_rows + 127) / 128), 128, 0, stream>>>(points, idx, ld_idx, H, (int)n_points,
                                                                            (int)m_rows, out);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_upsample_concat(const float* coarse, int64_t n_coarse, int c_coarse, const int32_t* up_idx,
                                   int ld_up, const float* fine, int c_fine, int64_t n_fine, float* out,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(c_coarse % 4 == 0 && c_fine % 4 == 0, "upsample_concat: channel counts must be multiples of 4");
  if (n_fine == 0) return LCR_OK;
  LcrProfScope prof("upsample_concat", 0.0, 8.0 * n_fine * (c_coarse + c_fine), stream);
  upsample_concat_kernel<<<(unsigned)((n_fine * 32 + 255) / 256), 256, 0, stream>>>(
      coarse, (int)n_coarse, c_coarse, up_idx, ld_up, fine, c_fine, n_fine, out);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
