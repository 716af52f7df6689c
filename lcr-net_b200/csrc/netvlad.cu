// netvlad.cu -- a8: NetVLAD (LOUPE) global-descriptor head, eval mode, batched over scans.
// Reference: experiments/lcrnet/modules/netvlad/NetVlad.py:49-87 (NetVLADLoupe2.forward),
// :189-201 (GatingContext), model_family/LCRNet_GlobalDescrition.py:34-38 (normalize before/after).
//
//   xn   = x / max(|x|, 1e-12)                                   (row-wise, 1024 channels)
//   a    = softmax(BN1(xn @ cluster_weights))                    [rows, 64]
//   V    = a^T xn - (sum_n a) * cluster_weights2                 per scan, stored [1024, 64]
//   V    = V / max(|V[:,k]|, 1e-6) ; v = flatten(V) / max(|V|, 1e-6)      (feature-major flatten)
//   h    = BN2(v @ hidden1_weights[65536, 256])
//   out  = normalize(h * sigmoid(BNg(h @ gating_weights)))
// The 67 MB hidden1_weights matrix is streamed ONCE per group of 16 scans (split-K over 512
// CTAs, deterministic two-pass reduction), so the head is HBM-bound on that read.
#include "common.cuh"

int lcr_gemm_f32(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                 const float* rowscale, const float* bias, cudaStream_t stream);

namespace {
constexpr int F = 1024;   // feature size
constexpr int KC = 64;    // clusters
constexpr int OD = 256;   // output dim
constexpr float BN_EPS = 1e-5f;

struct BnParams {
  const float *w, *b, *mean, *var;
};
__device__ __forceinline__ float bn_eval(float x, const BnParams& p, int c) {
  return (x - p.mean[c]) * (1.0f / sqrtf(p.var[c] + BN_EPS)) * p.w[c] + p.b[c];
}

__global__ void __launch_bounds__(256) row_normalize_kernel(const float* __restrict__ x, int64_t rows,
                                                            float* __restrict__ xn) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(x + r * F);
  float4 v[F / 128];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < F / 128; i++) {
    v[i] = src[lane + 32 * i];
    ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  ss = lcr_warp_sum(ss);
  const float d = fmaxf(sqrtf(ss), 1e-12f);
  float4* dst = reinterpret_cast<float4*>(xn + r * F);
#pragma unroll
  for (int i = 0; i < F / 128; i++)
    dst[lane + 32 * i] = make_float4(v[i].x / d, v[i].y / d, v[i].z / d, v[i].w / d);
}

// in place: a[r, :] = softmax(BN1(a[r, :])) over the 64 clusters
__global__ void __launch_bounds__(256) bn_softmax_kernel(float* __restrict__ a, int64_t rows, BnParams bn) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  float v0 = bn_eval(a[r * KC + lane], bn, lane), v1 = bn_eval(a[r * KC + lane + 32], bn, lane + 32);
  const float mx = lcr_warp_max(fmaxf(v0, v1));
  v0 = expf(v0 - mx);
  v1 = expf(v1 - mx);
  const float sum = lcr_warp_sum(v0 + v1);
  a[r * KC + lane] = v0 / sum;
  a[r * KC + lane + 32] = v1 / sum;
}

// V[s][c][k] = sum_n xn[n,c] a[n,k] - (sum_n a[n,k]) * cw2[c,k];  grid (F/64, S), 256 threads, 4x4 micro-tiles
__global__ void __launch_bounds__(256)
vlad_kernel(const float* __restrict__ xn, const float* __restrict__ a, const int64_t* __restrict__ scan_off,
            const float* __restrict__ cw2, float* __restrict__ vlad) {
  __shared__ __align__(16) float s_x[16][64];
  __shared__ __align__(16) float s_a[16][64];
  const int s = blockIdx.y, c0 = blockIdx.x * 64;
  const int64_t n0 = scan_off[s], n1 = scan_off[s + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx -> k, ty -> c
  float acc[4][4];
  float asum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  for (int64_t nb = n0; nb < n1; nb += 16) {
    {
      const int rr = threadIdx.x >> 4, q = threadIdx.x & 15;  // 16 rows x 16 float4
      const int64_t n = nb + rr;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&s_x[rr][q * 4]) = n < n1 ? *reinterpret_cast<const float4*>(xn + n * F + c0 + q * 4) : z;
      *reinterpret_cast<float4*>(&s_a[rr][q * 4]) = n < n1 ? *reinterpret_cast<const float4*>(a + n * KC + q * 4) : z;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) {
      const float4 xv = *reinterpret_cast<const float4*>(&s_x[r][ty * 4]);
      const float4 av = *reinterpret_cast<const float4*>(&s_a[r][tx * 4]);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, as[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(xs[i], as[j], acc[i][j]);
#pragma unroll
      for (int j = 0; j < 4; j++) asum[j] += as[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = c0 + ty * 4 + i;
    float4 o;
    float* op = &o.x;
#pragma unroll
    for (int j = 0; j < 4; j++) op[j] = acc[i][j] - asum[j] * cw2[c * KC + tx * 4 + j];
    *reinterpret_cast<float4*>(vlad + ((size_t)s * F + c) * KC + tx * 4) = o;
  }
}

// intra-normalisation over the 1024 features of each cluster, then global L2 normalisation.
__global__ void __launch_bounds__(256) vlad_normalize_kernel(float* __restrict__ vlad) {
  __shared__ float s_part[4][KC];
  __shared__ float s_norm[KC];
  __shared__ float s_red[8];
  float* V = vlad + (size_t)blockIdx.x * F * KC;
  const int k = threadIdx.x & 63, cl = threadIdx.x >> 6;
  float ss = 0.f;
  for (int c = cl; c < F; c += 4) {
    const float v = V[c * KC + k];
    ss = fmaf(v, v, ss);
  }
  s_part[cl][k] = ss;
  __syncthreads();
  if (threadIdx.x < KC)
    s_norm[threadIdx.x] = fmaxf(sqrtf((s_part[0][threadIdx.x] + s_part[1][threadIdx.x]) +
                                      (s_part[2][threadIdx.x] + s_part[3][threadIdx.x])), 1e-6f);
  __syncthreads();
  const float dk = s_norm[k];
  float gs = 0.f;
  for (int c = cl; c < F; c += 4) {
    const float v = V[c * KC + k] / dk;
    V[c * KC + k] = v;
    gs = fmaf(v, v, gs);
  }
  gs = lcr_warp_sum(gs);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = gs;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; w++) tot += s_red[w];
  const float dg = fmaxf(sqrtf(tot), 1e-6f);
  for (int c = cl; c < F; c += 4) V[c * KC + k] /= dg;
}

// split-K skinny GEMM: partial[chunk][s][o] = sum_{k in chunk} v[s][k] * W[k][o]
constexpr int HK = 128;   // K rows per CTA
constexpr int HS = 16;    // scans per CTA
__global__ void __launch_bounds__(OD)
hidden_partial_kernel(const float* __restrict__ v, const float* __restrict__ W, int S, float* __restrict__ partial) {
  __shared__ float s_v[HS][HK];
  const int chunk = blockIdx.x, s0 = blockIdx.y * HS;
  const int ns = min(HS, S - s0);
  for (int i = threadIdx.x; i < HS * HK; i += OD) {
    const int s = i / HK, kk = i % HK;
    s_v[s][kk] = s < ns ? v[(size_t)(s0 + s) * (F * KC) + (size_t)chunk * HK + kk] : 0.f;
  }
  __syncthreads();
  float acc[HS];
#pragma unroll
  for (int s = 0; s < HS; s++) acc[s] = 0.f;
  const float* Wp = W + (size_t)chunk * HK * OD + threadIdx.x;
#pragma unroll 8
  for (int kk = 0; kk < HK; kk++) {
    const float w = Wp[(size_t)kk * OD];
#pragma unroll
    for (int s = 0; s < HS; s++) acc[s] = fmaf(s_v[s][kk], w, acc[s]);
  }
  const int n_chunks = gridDim.x;
  for (int s = 0; s < ns; s++) partial[((size_t)(s0 + s) * n_chunks + chunk) * OD + threadIdx.x] = acc[s];
}

// one CTA per scan: fixed-order reduction of the split-K partials, BN2, context gating, normalise
__global__ void __launch_bounds__(OD)
head_tail_kernel(const float* __restrict__ partial, int n_chunks, BnParams bn2, const float* __restrict__ gw,
                 BnParams bng, float* __restrict__ out) {
  __shared__ float s_h[OD];
  __shared__ float s_red[OD / 32];
  const int s = blockIdx.x, o = threadIdx.x;
  const float* p = partial + (size_t)s * n_chunks * OD + o;
  float h = 0.f;
  for (int c = 0; c < n_chunks; c++) h += p[(size_t)c * OD];
  h = bn_eval(h, bn2, o);
  s_h[o] = h;
  __syncthreads();
  float g = 0.f;
  for (int i = 0; i < OD; i++) g = fmaf(s_h[i], gw[i * OD + o], g);
  g = bn_eval(g, bng, o);
  g = 1.f / (1.f + expf(-g));
  const float y = h * g;
  float ss = lcr_warp_sum(y * y);
  if ((o & 31) == 0) s_red[o >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < OD / 32; w++) tot += s_red[w];
  out[(size_t)s * OD + o] = y / fmaxf(sqrtf(tot), 1e-12f);
}
}  // namespace

extern "C" size_t lcr_netvlad_ws_bytes(int64_t rows, int n_scans) {
  const size_t n_chunks = (size_t)F * KC / HK;
  return lcr_align_up((size_t)rows * F * 4) + lcr_align_up((size_t)rows * KC * 4) +
         lcr_align_up((size_t)n_scans * F * KC * 4) + lcr_align_up((size_t)n_scans * n_chunks * OD * 4) + 1024;
}

extern "C" int lcr_netvlad(const float* feats, int64_t rows, const int64_t* scan_off, int n_scans,
                           const float* cluster_weights, const float* cluster_weights2, const float* hidden1_weights,
                           const float* bn1 /*w,b,mean,var: 4 x 64*/, const float* bn2 /*4 x 256*/,
                           const float* gating_weights, const float* gating_bn /*4 x 256*/, float* out /*[S,256]*/,
                           void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(n_scans >= 1 && rows >= 0 && rows < (1ll << 31), "netvlad: sizes");
  LCR_REQUIRE(ws && ws_bytes >= lcr_netvlad_ws_bytes(rows, n_scans), "netvlad: workspace too small");
  LcrArena ar(ws, ws_bytes);
  float* xn = ar.take<float>((size_t)rows * F);
  float* act = ar.take<float>((size_t)rows * KC);
  float* vlad = ar.take<float>((size_t)n_scans * F * KC);
  const int n_chunks = F * KC / HK;
  float* partial = ar.take<float>((size_t)n_scans * n_chunks * OD);
  const BnParams p1{bn1, bn1 + KC, bn1 + 2 * KC, bn1 + 3 * KC};
  const BnParams p2{bn2, bn2 + OD, bn2 + 2 * OD, bn2 + 3 * OD};
  const BnParams pg{gating_bn, gating_bn + OD, gating_bn + 2 * OD, gating_bn + 3 * OD};
  LcrProfScope prof_all("netvlad_total", 0.0, 0.0, stream);
  if (rows > 0) {
    const unsigned gw = (unsigned)((rows * 32 + 255) / 256);
    row_normalize_kernel<<<gw, 256, 0, stream>>>(feats, rows, xn);
    int rc = lcr_gemm_f32(xn, F, cluster_weights, KC, act, KC, (int)rows, KC, F, nullptr, nullptr, stream);
    if (rc != LCR_OK) return rc;
    bn_softmax_kernel<<<gw, 256, 0, stream>>>(act, rows, p1);
  }
  vlad_kernel<<<dim3(F / 64, n_scans), 256, 0, stream>>>(xn, act, scan_off, cluster_weights2, vlad);
  vlad_normalize_kernel<<<n_scans, 256, 0, stream>>>(vlad);
  {
  LcrProfScope prof("netvlad_hidden", 2.0 * n_scans * (double)F * KC * OD,
                    4.0 * (double)F * KC * OD * ((n_scans + HS - 1) / HS) + 4.0 * (double)n_scans * F * KC, stream);
  hidden_partial_kernel<<<dim3(n_chunks, (n_scans + HS - 1) / HS), OD, 0, stream>>>(vlad, hidden1_weights, n_scans,
                                                                                   partial);
  }
  head_tail_kernel<<<n_scans, OD, 0, stream>>>(partial, n_chunks, p2, gating_weights, pg, out);
  LCR_LAUNCHED(4 + (rows > 0 ? 2 : 0));  // + the GEMM, which counts itself
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
