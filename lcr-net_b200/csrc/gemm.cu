// gemm.cu -- fp32 SIMT GEMM used by the KPConv contraction, the unary (Linear) layers and the
// NetVLAD projections:  C[M,N] = rowscale[m] * (A[M,K] . B[K,N]) + bias[n].
//
// Why SIMT fp32 and not tcgen05: the encoder output feeds a descriptor that must match the
// reference (fp32 torch) within 1e-4 relative through 11 blocks with data-dependent
// normalisers; tensor cores offer no fp32 MMA (TF32 has a 10-bit mantissa).  The tile kernel
// is the classic register-blocked design: 256 threads, BK = 16, A staged transposed in shared
// memory so both operands are read with conflict-free LDS.128, global->register prefetch of the
// next K-slab overlapped with the FMA block of the current one.
#include "common.cuh"

namespace {

constexpr int BK = 16;
constexpr int APAD = 4;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
            int M, int N, int K, const float* __restrict__ rowscale, const float* __restrict__ bias, int relu) {
  constexpr int CT = BN / TN;        // threads along N
  constexpr int RT = BM / TM;        // threads along M
  static_assert(CT * RT == 256, "256 threads per CTA");
  constexpr int A_F4 = BM * BK / 4 / 256;                         // float4 loads of A per thread
  constexpr int B_F4 = (BK * BN / 4 + 255) / 256;                 // float4 loads of B per thread
  constexpr int RC = TM / 4, CC = TN / 4;                         // 4-wide chunks per thread
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid % CT, ty = tid / CT;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  float4 ra[A_F4], rb[B_F4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_F4; i++) {
      const int f = tid + i * 256;
      const int row = f >> 2, kq = f & 3;
      const int gm = m0 + row, gk = k0 + kq * 4;
      ra[i] = (gm < M && gk < K) ? *reinterpret_cast<const float4*>(A + (size_t)gm * lda + gk)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_F4; i++) {
      const int f = tid + i * 256;
      const int kr = f / (BN / 4), nq = f % (BN / 4);
      const int gk = k0 + kr, gn = n0 + nq * 4;
      rb[i] = (kr < BK && gk < K && gn < N) ? *reinterpret_cast<const float4*>(B + (size_t)gk * ldb + gn)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_F4; i++) {
      const int f = tid + i * 256;
      const int row = f >> 2, kq = f & 3;
      As[buf][kq * 4 + 0][row] = ra[i].x;
      As[buf][kq * 4 + 1][row] = ra[i].y;
      As[buf][kq * 4 + 2][row] = ra[i].z;
      As[buf][kq * 4 + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_F4; i++) {
      const int f = tid + i * 256;
      const int kr = f / (BN / 4), nq = f % (BN / 4);
      if (kr < BK) *reinterpret_cast<float4*>(&Bs[buf][kr][nq * 4]) = rb[i];
    }
  };

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[TM], b[TN];
#pragma unroll
      for (int r = 0; r < RC; r++) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][r * (RT * 4) + ty * 4]);
        a[r * 4 + 0] = v.x; a[r * 4 + 1] = v.y; a[r * 4 + 2] = v.z; a[r * 4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < CC; c++) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][c * (CT * 4) + tx * 4]);
        b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int r = 0; r < RC; r++)
#pragma unroll
    for (int ii = 0; ii < 4; ii++) {
      const int gm = m0 + r * (RT * 4) + ty * 4 + ii;
      if (gm >= M) continue;
      const float rs = rowscale ? rowscale[gm] : 1.f;
#pragma unroll
      for (int c = 0; c < CC; c++) {
        const int gn = n0 + c * (CT * 4) + tx * 4;
        if (gn >= N) continue;
        float4 v;
        v.x = acc[r * 4 + ii][c * 4 + 0] * rs;
        v.y = acc[r * 4 + ii][c * 4 + 1] * rs;
        v.z = acc[r * 4 + ii][c * 4 + 2] * rs;
        v.w = acc[r * 4 + ii][c * 4 + 3] * rs;
        if (bias) {
          const float4 bv = *reinterpret_cast<const float4*>(bias + gn);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        }
        if (relu) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        *reinterpret_cast<float4*>(C + (size_t)gm * ldc + gn) = v;
      }
    }
}

template <int BM, int BN, int TM, int TN>
void launch(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
            const float* rowscale, const float* bias, int relu, cudaStream_t stream) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_kernel<BM, BN, TM, TN><<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, M, N, K, rowscale, bias, relu);
}

}  // namespace

// Internal entry (also used by encoder.cu / netvlad.cu).
int lcr_gemm_f32_act(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                     const float* rowscale, const float* bias, int relu, cudaStream_t stream) {
  LCR_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad shape");
  LCR_REQUIRE((N % 4) == 0 && (K % 4) == 0 && (lda % 4) == 0 && (ldb % 4) == 0 && (ldc % 4) == 0,
              "gemm: N, K and leading dimensions must be multiples of 4");
  LCR_REQUIRE((((uintptr_t)A | (uintptr_t)B | (uintptr_t)C | (uintptr_t)bias) & 15) == 0,
              "gemm: pointers must be 16-byte aligned");
  if (M == 0) return LCR_OK;
  LcrProfScope prof("gemm_f32", 2.0 * M * N * K, 4.0 * ((double)M * K + (double)K * N + (double)M * N), stream);
  // tile choice: keep >= ~1 wave of CTAs on 148 SMs when the problem allows it
  const long ctas_128 = (long)((M + 127) / 128) * ((N + 127) / 128);
  if (N <= 32) {
    launch<256, 32, 8, 4>(A, lda, B, ldb, C, ldc, M, N, K, rowscale, bias, relu, stream);
  } else if (N <= 64 || ctas_128 < LCR_SM_COUNT) {
    if ((long)((M + 127) / 128) * ((N + 63) / 64) < LCR_SM_COUNT)
      launch<64, 64, 4, 4>(A, lda, B, ldb, C, ldc, M, N, K, rowscale, bias, relu, stream);
    else
      launch<128, 64, 8, 4>(A, lda, B, ldb, C, ldc, M, N, K, rowscale, bias, relu, stream);
  } else {
    launch<128, 128, 8, 8>(A, lda, B, ldb, C, ldc, M, N, K, rowscale, bias, relu, stream);
  }
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

int lcr_gemm_f32(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                 const float* rowscale, const float* bias, cudaStream_t stream) {
  return lcr_gemm_f32_act(A, lda, B, ldb, C, ldc, M, N, K, rowscale, bias, 0, stream);
}

// out = act(x . weight_t + bias), act: 0 none, 1 ReLU; optional per-row scale applied before the bias
extern "C" int lcr_linear_ex(const float* x, int64_t n_rows, int c_in, int ld_x, const float* weight_t, int c_out,
                             const float* bias, const float* rowscale, int act, float* out, int ld_out, void* stream) {
  LCR_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "linear: n_rows out of range");
  return lcr_gemm_f32_act(x, ld_x, weight_t, c_out, out, ld_out, (int)n_rows, c_out, c_in, rowscale, bias, act,
                          (cudaStream_t)stream);
}

extern "C" int lcr_linear(const float* x, int64_t n_rows, int c_in, const float* weight_t, int c_out,
                          const float* bias, float* out, void* stream) {
  LCR_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "linear: n_rows out of range");
  return lcr_gemm_f32(x, c_in, weight_t, c_out, out, c_out, (int)n_rows, c_out, c_in, nullptr, bias,
                      (cudaStream_t)stream);
}
