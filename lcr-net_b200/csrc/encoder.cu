// encoder.cu -- a4/a5/a6: KPConv, GroupNorm (+LeakyReLU, residual), max-pool for the KPConv
// encoder (reference: experiments/lcrnet/modules/kpconv/kpconv.py:79-122,
// modules/kpconv/modules.py:33-225, modules/kpconv/functional.py:54-67, backbone4.py:11-89).
//
// Data layout in HBM: features row-major f32 [rows, C]; points f32 [rows, 3]; neighbour tables
// int32 [M, ld_idx] with pad value = number of support rows (the int64 tables of the reference
// boundary are narrowed once on entry).  Several "stacks" (the unit of one reference forward:
// one scan, or the two scans of a pair) are concatenated along rows; GroupNorm statistics are
// per stack (stack_off[S+1] row offsets), everything else is oblivious to stack boundaries
// because neighbour indices never cross clouds.
//
// KPConv is split into
//   (A) kpconv_gather_kernel: irregular gather, one warp per query point.  Lanes compute the 15
//       kernel-point influences of 32 neighbours at a time into shared memory, then every lane
//       owns C/32 channels and accumulates wf[k][c] = sum_h w[k,h] * feat[idx[h], c] in registers
//       with coalesced 128 B feature-row loads.  Output wf[M, 15*C] and 1/neighbour_num.
//   (B) the fp32 GEMM of gemm.cu: [M, 15C] x [15C, O] with the 1/num row scale and bias fused.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

int lcr_gemm_f32(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                 const float* rowscale, const float* bias, cudaStream_t stream);
int lcr_gemm_tf32x3(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N,
                    int K, const float* rowscale, const float* bias, int relu, cudaStream_t stream);
int lcr_gemm_tf32x3_gn(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M,
                       int N, int K, const float* rowscale, const float* bias, int relu, float* gn_partial,
                       const int64_t* stack_off, int n_stacks, cudaStream_t stream);

namespace {

constexpr int KP = 15;  // kernel points (config_model.py:35)

// ------------------------------------------------------------------ KPConv gather
// Kernel points by value: they land in the constant bank and feed the FADDs as immediate operands.
struct KpArg { float v[KP * 3]; };

// FAST: kernel points from the constant bank (kp) instead of shared memory (kpts), and the influence
// 1 - d/sigma with d = d2 * rsqrt(d2) and a multiply by 1/sigma instead of the IEEE sqrt + divide
// sequences (<= 3e-7 absolute from the reference's weights; parity bar 1e-4): the influence pass is
// ~2.5x shorter, which matters most for the narrow layers where it outweighs the accumulation.
// GS ("group skip", FAST only): the influence pass also leaves a 4-bit mask per neighbour saying which of the four
// float4 groups of kernel points hold a non-zero influence (~1.4 of 4); the accumulation loads and multiplies
// only those groups.  The dense loop is bound by L1 wavefronts (4 broadcast LDS.128 + the feature row per
// neighbour) and issue slots (15 FFMA per channel): both drop.
template <int CPL, bool FAST, bool GS = false>  // channels per lane: C = 32 * CPL
__global__ void __launch_bounds__(128)
kpconv_gather_kernel(const float* __restrict__ s_feats, const float* __restrict__ q_pts,
                     const float* __restrict__ s_pts, const int32_t* __restrict__ idx, int ld_idx, int H,
                     const float* __restrict__ kpts, const KpArg kp, float sigma,
                     const uint8_t* __restrict__ flags, const float4* __restrict__ s_pts4, int M, int N,
                     float* __restrict__ wf, float* __restrict__ rowscale) {
  constexpr int C = 32 * CPL;
  // influences of 32 neighbours, packed as float4 groups of kernel points: [k/4][neighbour] so that
  // lanes write conflict-free 16-B vectors and the accumulation loop reads 4 broadcast LDS.128
  __shared__ __align__(16) float4 s_w[4][4][32];
  __shared__ int s_j[4][32];
  __shared__ float s_kp[KP * 3];
  __shared__ unsigned s_gm[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (!FAST) {
    if (threadIdx.x < KP * 3) s_kp[threadIdx.x] = kpts[threadIdx.x];
    __syncthreads();
  }
  const int m = blockIdx.x * 4 + warp;
  if (m >= M) return;
  const float qx = q_pts[3 * m], qy = q_pts[3 * m + 1], qz = q_pts[3 * m + 2];
  const float inv_sigma = 1.f / sigma;
  float acc[KP][CPL];
#pragma unroll
  for (int k = 0; k < KP; k++)
#pragma unroll
    for (int i = 0; i < CPL; i++) acc[k][i] = 0.f;
  int cnt = 0;
  const int32_t* row = idx + (size_t)m * ld_idx;
  for (int h0 = 0; h0 < H; h0 += 32) {
    const int h = h0 + lane;
    const int j = h < H ? row[h] : N;
    const bool valid = j < N;
    // valid neighbours are compacted to the front (rows are sorted with the pads last, but holes are
    // tolerated) so the accumulation loop below is a counted loop the compiler can software-pipeline
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int slot = __popc(vmask & ((1u << lane) - 1u));
    const int nv = __popc(vmask);
    if (valid) {
      // FAST: one 16-byte load of (x, y, z, row flag) per neighbour instead of three scattered 4-byte loads
      // plus a scattered byte load: the kernel is bound by L1 wavefronts (ncu: l1tex data-pipe 72-88 %), and a
      // scattered load costs one wavefront per lane whatever its width
      float4 sp;
      if (FAST) {
        sp = s_pts4[j];
      } else {
        sp = make_float4(s_pts[3 * (size_t)j], s_pts[3 * (size_t)j + 1], s_pts[3 * (size_t)j + 2],
                         flags ? (float)flags[j] : 1.f);
      }
      const float rx = sp.x - qx, ry = sp.y - qy, rz = sp.z - qz;
      float w[16];
#pragma unroll
      for (int k = 0; k < KP; k++) {
        if (FAST) {
          const float dx = rx - kp.v[3 * k], dy = ry - kp.v[3 * k + 1], dz = rz - kp.v[3 * k + 2];
          const float d2 = dx * dx + dy * dy + dz * dz;
          w[k] = fmaxf(fmaf(-d2 * rsqrtf(fmaxf(d2, 1e-30f)), inv_sigma, 1.f), 0.f);
        } else {
          const float dx = rx - s_kp[3 * k], dy = ry - s_kp[3 * k + 1], dz = rz - s_kp[3 * k + 2];
          const float d2 = dx * dx + dy * dy + dz * dz;
          w[k] = fmaxf(1.f - __fdiv_rn(sqrtf(d2), sigma), 0.f);
        }
      }
      w[15] = 0.f;
#pragma unroll
      for (int g = 0; g < 4; g++) s_w[warp][g][slot] = make_float4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
      s_j[warp][slot] = j;
      if (GS) {
        unsigned gm = 0;
#pragma unroll
        for (int g = 0; g < 4; g++)
          gm |= ((w[4 * g] > 0.f) | (w[4 * g + 1] > 0.f) | (w[4 * g + 2] > 0.f) | (w[4 * g + 3] > 0.f) ? 1u : 0u) << g;
        s_gm[warp][slot] = gm;
      }
      cnt += sp.w != 0.f;
    }
    __syncwarp();
    constexpr int UN = CPL >= 4 ? 2 : 4;  // neighbours whose feature rows are in flight together
    int hh = 0;
    for (; hh + UN <= nv; hh += UN) {
      float fv[UN][CPL];
#pragma unroll
      for (int u = 0; u < UN; u++) {
        const float* f = s_feats + (size_t)s_j[warp][hh + u] * C + lane;
#pragma unroll
        for (int i = 0; i < CPL; i++) fv[u][i] = f[32 * i];
      }
      if (GS) {
#pragma unroll
        for (int u = 0; u < UN; u++) {
          const unsigned gm = s_gm[warp][hh + u];
#pragma unroll
          for (int g = 0; g < 4; g++) {
            if (!(gm & (1u << g))) continue;            // warp-uniform
            const float4 t = s_w[warp][g][hh + u];
            const float wv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              if (4 * g + kk >= KP) continue;
#pragma unroll
              for (int i = 0; i < CPL; i++) acc[4 * g + kk][i] = fmaf(wv[kk], fv[u][i], acc[4 * g + kk][i]);
            }
          }
        }
        continue;
      }
#pragma unroll
      for (int u = 0; u < UN; u++) {
        float w[16];
#pragma unroll
        for (int g = 0; g < 4; g++) *reinterpret_cast<float4*>(w + 4 * g) = s_w[warp][g][hh + u];
#pragma unroll
        for (int k = 0; k < KP; k++) {
          // a neighbour lies inside the influence radius of only a few of the 15 kernel points; the
          // weight is warp-uniform, so for wide channel slices skipping the zero ones pays
          if (CPL >= 4 && w[k] == 0.f) continue;
#pragma unroll
          for (int i = 0; i < CPL; i++) acc[k][i] = fmaf(w[k], fv[u][i], acc[k][i]);
        }
      }
    }
    for (; hh < nv; hh++) {
      const float* f = s_feats + (size_t)s_j[warp][hh] * C + lane;
      float fv[CPL];
#pragma unroll
      for (int i = 0; i < CPL; i++) fv[i] = f[32 * i];
      float w[16];
#pragma unroll
      for (int g = 0; g < 4; g++) *reinterpret_cast<float4*>(w + 4 * g) = s_w[warp][g][hh];
#pragma unroll
      for (int k = 0; k < KP; k++) {
        if (CPL >= 4 && w[k] == 0.f) continue;
#pragma unroll
        for (int i = 0; i < CPL; i++) acc[k][i] = fmaf(w[k], fv[i], acc[k][i]);
      }
    }
    __syncwarp();
  }
  cnt = lcr_warp_sum(cnt);
  float* o = wf + (size_t)m * (KP * C) + lane;
#pragma unroll
  for (int k = 0; k < KP; k++)
#pragma unroll
    for (int i = 0; i < CPL; i++) o[k * C + 32 * i] = acc[k][i];
  if (lane == 0) rowscale[m] = 1.f / (float)max(cnt, 1);
}

// (x, y, z, flag) per support point for the fast gather
__global__ void pack_points_kernel(const float* __restrict__ pts, const uint8_t* __restrict__ flags, int N,
                                   float4* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N) out[j] = make_float4(pts[3 * j], pts[3 * j + 1], pts[3 * j + 2], flags ? (float)flags[j] : 1.f);
}

// ------------------------------------------------------------------ KPConv gather, packed-FMA form
// The dense loop above is bound by instruction issue (ncu: issue slots ~70 % busy, FMA pipe 40-48 %): per
// neighbour and channel it issues 15 FFMA.  Blackwell has a packed fp32 FMA (fma.rn.f32x2, SASS FFMA2: two
// independent IEEE fp32 FMAs per lane and instruction, same results as two FFMA), so the 15 kernel points
// are accumulated as 8 PAIRS: a = (w[2p], w[2p+1]) -- adjacent registers of the influence vector as it comes
// out of shared memory --, b = (f, f), c = (acc[2p], acc[2p+1]); the 16th weight is a zero pad.  8 FFMA2 + 1
// MOV per neighbour and channel instead of 15 FFMA.  For wide channel slices (C >= 128) a pair whose two
// influences are both zero is skipped (a neighbour is reached by ~1.6 of the 15 kernel points).
template <int CPL>
__global__ void __launch_bounds__(128)
kpconv_gather_packed_kernel(const float* __restrict__ s_feats, const float* __restrict__ q_pts,
                            const float* __restrict__ s_pts, const int32_t* __restrict__ idx, int ld_idx, int H,
                            const KpArg kp, float sigma, const uint8_t* __restrict__ flags, int M, int N,
                            float* __restrict__ wf, float* __restrict__ rowscale) {
  constexpr int C = 32 * CPL;
  constexpr int KPP = (KP + 1) / 2;   // kernel-point pairs
  __shared__ __align__(16) float4 s_w[4][4][32];
  __shared__ int s_j[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * 4 + warp;
  if (m >= M) return;
  const float qx = q_pts[3 * m], qy = q_pts[3 * m + 1], qz = q_pts[3 * m + 2];
  const float inv_sigma = 1.f / sigma;
  float2 acc[KPP][CPL];
#pragma unroll
  for (int k = 0; k < KPP; k++)
#pragma unroll
    for (int i = 0; i < CPL; i++) acc[k][i] = make_float2(0.f, 0.f);
  int cnt = 0;
  const int32_t* row = idx + (size_t)m * ld_idx;
  for (int h0 = 0; h0 < H; h0 += 32) {
    const int h = h0 + lane;
    const int j = h < H ? row[h] : N;
    const bool valid = j < N;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int slot = __popc(vmask & ((1u << lane) - 1u));
    const int nv = __popc(vmask);
    if (valid) {
      const float rx = s_pts[3 * (size_t)j] - qx, ry = s_pts[3 * (size_t)j + 1] - qy,
                  rz = s_pts[3 * (size_t)j + 2] - qz;
      float w[16];
#pragma unroll
      for (int k = 0; k < KP; k++) {
        const float dx = rx - kp.v[3 * k], dy = ry - kp.v[3 * k + 1], dz = rz - kp.v[3 * k + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        w[k] = fmaxf(fmaf(-d2 * rsqrtf(fmaxf(d2, 1e-30f)), inv_sigma, 1.f), 0.f);
      }
      w[15] = 0.f;
#pragma unroll
      for (int g = 0; g < 4; g++) s_w[warp][g][slot] = make_float4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
      s_j[warp][slot] = j;
      cnt += flags ? (int)flags[j] : 1;
    }
    __syncwarp();
    constexpr int UN = CPL >= 4 ? 2 : 4;  // neighbours whose feature rows are in flight together
    auto accumulate = [&](const float (&f)[CPL], int hh) {
      float2 wp[KPP];
#pragma unroll
      for (int g = 0; g < 4; g++) {
        const float4 t = s_w[warp][g][hh];
        wp[2 * g] = make_float2(t.x, t.y);
        wp[2 * g + 1] = make_float2(t.z, t.w);
      }
#pragma unroll
      for (int k = 0; k < KPP; k++) {
        if (CPL >= 4 && wp[k].x == 0.f && wp[k].y == 0.f) continue;   // warp-uniform
#pragma unroll
        for (int i = 0; i < CPL; i++) acc[k][i] = __ffma2_rn(wp[k], make_float2(f[i], f[i]), acc[k][i]);
      }
    };
    int hh = 0;
    for (; hh + UN <= nv; hh += UN) {
      float fv[UN][CPL];
#pragma unroll
      for (int u = 0; u < UN; u++) {
        const float* f = s_feats + (size_t)s_j[warp][hh + u] * C + lane;
#pragma unroll
        for (int i = 0; i < CPL; i++) fv[u][i] = f[32 * i];
      }
#pragma unroll
      for (int u = 0; u < UN; u++) accumulate(fv[u], hh + u);
    }
    for (; hh < nv; hh++) {
      const float* f = s_feats + (size_t)s_j[warp][hh] * C + lane;
      float fv[CPL];
#pragma unroll
      for (int i = 0; i < CPL; i++) fv[i] = f[32 * i];
      accumulate(fv, hh);
    }
    __syncwarp();
  }
  cnt = lcr_warp_sum(cnt);
  float* o = wf + (size_t)m * (KP * C) + lane;
#pragma unroll
  for (int k = 0; k < KPP; k++)
#pragma unroll
    for (int i = 0; i < CPL; i++) {
      o[(2 * k) * C + 32 * i] = acc[k][i].x;
      if (2 * k + 1 < KP) o[(2 * k + 1) * C + 32 * i] = acc[k][i].y;
    }
  if (lane == 0) rowscale[m] = 1.f / (float)max(cnt, 1);
}

// ------------------------------------------------------------------ KPConv gather on the warp-level tensor path
// wf[m][k][c] = sum_h w[m][h][k] * feat[idx[m][h]][c] is, per query, a [16 x H] x [H x C] product whose M = 15
// kernel points (+1 pad) fits mma.sync.m16n8k8 exactly (the 128-row tcgen05 tile would need a block-diagonal
// operand with 7/8 structural zeros).  fp32-class accuracy by the same 3xTF32 operand split as the GEMMs
// (hi = upper 19 bits, lo = x - hi; hi*hi + hi*lo + lo*hi, fp32 accumulate).  Measured on B200: the legacy
// mma.sync TF32 path sustains 278 TFLOP/s (scripts/micro/mma_sync_tf32.cu), 3.9x the FFMA rate, i.e. 1.3x after
// the 3-fold split -- but, more important here, it removes the shared-memory influence broadcast that made the
// FFMA loop L1-wavefront bound (ncu: l1tex data pipe 72-88 %): every influence is computed by exactly the lane
// whose A fragment needs it, and the B fragments come straight from 16-byte feature loads.
//   A fragment (weights): lane (g = lane / 4, t = lane % 4) holds k in {g, g + 8} x h in {8 s + t, 8 s + t + 4}
//   B fragments (features): n-tile j of a 32-channel set covers channels {4 g' + j}: the lane's float4 load of
//   channels 4 g .. 4 g + 3 of row h feeds the four n-tiles j = 0..3 (one 128-byte row = 8 lanes x 16 B)
//   D fragments: lane ends with channels 8 t .. 8 t + 7 of rows k = g and g + 8 -> two float4 stores each
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int CPL>
__global__ void __launch_bounds__(128)
kpconv_gather_mma_kernel(const float* __restrict__ s_feats, const float* __restrict__ q_pts,
                         const float4* __restrict__ s_pts4, const int32_t* __restrict__ idx, int ld_idx, int H,
                         const float* __restrict__ kpts, float sigma, int M, int N, float* __restrict__ wf,
                         float* __restrict__ rowscale) {
  constexpr int C = 32 * CPL;
  constexpr int NT = 4 * CPL;          // 8-channel n-tiles
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * 4 + warp;
  if (m >= M) return;
  const int g = lane >> 2, t = lane & 3;
  const float qx = q_pts[3 * m], qy = q_pts[3 * m + 1], qz = q_pts[3 * m + 2];
  const float inv_sigma = 1.f / sigma;
  // kernel points g and g + 8 (row 15 is the pad: its weights are forced to zero)
  const float k0x = kpts[3 * g], k0y = kpts[3 * g + 1], k0z = kpts[3 * g + 2];
  const int g1 = g + 8 < KP ? g + 8 : g;
  const float k1x = kpts[3 * g1], k1y = kpts[3 * g1 + 1], k1z = kpts[3 * g1 + 2];
  const float pad1 = g + 8 < KP ? 1.f : 0.f;
  float d[NT][4];
#pragma unroll
  for (int n = 0; n < NT; n++)
#pragma unroll
    for (int i = 0; i < 4; i++) d[n][i] = 0.f;
  int cnt = 0;
  const int32_t* row = idx + (size_t)m * ld_idx;
  auto influence = [&](float rx, float ry, float rz, float kx, float ky, float kz) {
    const float dx = rx - kx, dy = ry - ky, dz = rz - kz;
    const float d2 = dx * dx + dy * dy + dz * dz;
    return fmaxf(fmaf(-d2 * rsqrtf(fmaxf(d2, 1e-30f)), inv_sigma, 1.f), 0.f);
  };
  // (an explicitly software-pipelined variant -- indices two k-steps ahead, rows one ahead -- measured 20-30 %
  // slower: more registers, fewer resident warps; the hardware overlaps the k-steps of the 32 warps per SM)
  for (int h0 = 0; h0 < H; h0 += 32) {
    const int jl = h0 + lane < H ? row[h0 + lane] : N;       // 32 neighbour indices, one coalesced load
    const unsigned vmask = __ballot_sync(0xffffffffu, jl < N);
    if (vmask == 0) continue;
#pragma unroll
    for (int s = 0; s < 4; s++) {                             // k-steps of 8 neighbours
      if (((vmask >> (8 * s)) & 0xFFu) == 0) continue;        // warp-uniform
      const int ja = __shfl_sync(0xffffffffu, jl, 8 * s + t), jb = __shfl_sync(0xffffffffu, jl, 8 * s + t + 4);
      const bool va = ja < N, vb = jb < N;
      const float4 pa = va ? s_pts4[ja] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 pb = vb ? s_pts4[jb] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float* fa_p = s_feats + (size_t)(va ? ja : 0) * C + 4 * g;
      const float* fb_p = s_feats + (size_t)(vb ? jb : 0) * C + 4 * g;
      float4 fa[CPL], fb[CPL];
#pragma unroll
      for (int c = 0; c < CPL; c++) {
        fa[c] = *reinterpret_cast<const float4*>(fa_p + 32 * c);
        fb[c] = *reinterpret_cast<const float4*>(fb_p + 32 * c);
      }
      if (g == 0) cnt += (va && pa.w != 0.f) + (vb && pb.w != 0.f);
      const float ax = pa.x - qx, ay = pa.y - qy, az = pa.z - qz;
      const float bx = pb.x - qx, by = pb.y - qy, bz = pb.z - qz;
      float w[4];
      w[0] = va ? influence(ax, ay, az, k0x, k0y, k0z) : 0.f;            // (k = g,     h = 8 s + t)
      w[1] = va ? influence(ax, ay, az, k1x, k1y, k1z) * pad1 : 0.f;     // (k = g + 8, h = 8 s + t)
      w[2] = vb ? influence(bx, by, bz, k0x, k0y, k0z) : 0.f;            // (k = g,     h = 8 s + t + 4)
      w[3] = vb ? influence(bx, by, bz, k1x, k1y, k1z) * pad1 : 0.f;     // (k = g + 8, h = 8 s + t + 4)
      unsigned a_hi[4], a_lo[4];
#pragma unroll
      for (int i = 0; i < 4; i++) split_tf32(w[i], a_hi[i], a_lo[i]);
#pragma unroll
      for (int c = 0; c < CPL; c++) {
        const float va4[4] = {fa[c].x, fa[c].y, fa[c].z, fa[c].w}, vb4[4] = {fb[c].x, fb[c].y, fb[c].z, fb[c].w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          unsigned b0h, b0l, b1h, b1l;
          split_tf32(va ? va4[j] : 0.f, b0h, b0l);
          split_tf32(vb ? vb4[j] : 0.f, b1h, b1l);
          mma_tf32_16x8x8(d[4 * c + j], a_hi, b0h, b1h);
          mma_tf32_16x8x8(d[4 * c + j], a_hi, b0l, b1l);
          mma_tf32_16x8x8(d[4 * c + j], a_lo, b0h, b1h);
        }
      }
    }
  }
  cnt = lcr_warp_sum(cnt);
  // D fragment of n-tile j: d[0] (k = g, col 2t), d[1] (k = g, col 2t + 1), d[2] / d[3] the same for k = g + 8;
  // column col of n-tile j of channel set c is channel 32 c + 4 col + j
  float* o = wf + (size_t)m * (KP * C);
#pragma unroll
  for (int c = 0; c < CPL; c++) {
    float* o0 = o + (size_t)g * C + 32 * c + 8 * t;
    *reinterpret_cast<float4*>(o0) = make_float4(d[4 * c][0], d[4 * c + 1][0], d[4 * c + 2][0], d[4 * c + 3][0]);
    *reinterpret_cast<float4*>(o0 + 4) = make_float4(d[4 * c][1], d[4 * c + 1][1], d[4 * c + 2][1], d[4 * c + 3][1]);
    if (g + 8 < KP) {
      float* o1 = o + (size_t)(g + 8) * C + 32 * c + 8 * t;
      *reinterpret_cast<float4*>(o1) = make_float4(d[4 * c][2], d[4 * c + 1][2], d[4 * c + 2][2], d[4 * c + 3][2]);
      *reinterpret_cast<float4*>(o1 + 4) = make_float4(d[4 * c][3], d[4 * c + 1][3], d[4 * c + 2][3], d[4 * c + 3][3]);
    }
  }
  if (lane == 0) rowscale[m] = 1.f / (float)max(cnt, 1);
}

// ------------------------------------------------------------------ KPConv gather, sparse form
// A neighbour lies inside the influence radius sigma of only ~1.6 of the 15 kernel points (measured
// on the synthetic and demo pyramids; never more than 5), so ~90 % of the dense FFMAs above multiply
// by zero.  This kernel builds, per query and kernel point, the compacted list of (neighbour, weight)
// entries with non-zero influence (warp ballot + popc, neighbour order preserved), then accumulates
//   wf[k][:] = sum_{e in list k} w_e * feat[j_e][:]
// with one shared-memory broadcast, one vector feature load and CPL FFMAs per ENTRY instead of per
// (neighbour, kernel point) pair.  Lists are padded with zero-weight entries to a multiple of UN so
// UN feature rows are in flight without predication.  Kernel points are passed by value (constant
// bank operands).  The weight is 1 - d/sigma with d = d2 * rsqrt(d2) (<= 3e-7 absolute from the
// reference's sqrt and divide; parity bar 1e-4).

template <int CPL>
__global__ void __launch_bounds__(128)
kpconv_gather_sparse_kernel(const float* __restrict__ s_feats, const float* __restrict__ q_pts,
                            const float* __restrict__ s_pts, const int32_t* __restrict__ idx, int ld_idx, int H,
                            const KpArg kp, float sigma, const uint8_t* __restrict__ flags, int M, int N,
                            float* __restrict__ wf, float* __restrict__ rowscale) {
  constexpr int C = 32 * CPL;
  constexpr int VEC = CPL >= 4 ? 4 : CPL;       // channels per vector load: lane owns [lane*VEC, +VEC) of every 32*VEC slab
  constexpr int NV = CPL / VEC;
  constexpr int UN = CPL >= 8 ? 2 : 4;          // feature rows in flight per lane
  constexpr int CAP = 64 + 4;                   // entries per kernel point per 64-neighbour super-chunk
  __shared__ __align__(16) float2 s_ent[4][KP][CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * 4 + warp;
  if (m >= M) return;
  const float qx = q_pts[3 * m], qy = q_pts[3 * m + 1], qz = q_pts[3 * m + 2];
  const float sig2 = sigma * sigma * 1.000001f, inv_sigma = 1.f / sigma;
  const unsigned lt = (1u << lane) - 1u;
  float acc[KP][CPL];
#pragma unroll
  for (int k = 0; k < KP; k++)
#pragma unroll
    for (int i = 0; i < CPL; i++) acc[k][i] = 0.f;
  int cnt = 0;
  const int32_t* row = idx + (size_t)m * ld_idx;
  float2(*ent)[CAP] = s_ent[warp];
  for (int h0 = 0; h0 < H; h0 += 64) {
    int nk[KP];
#pragma unroll
    for (int k = 0; k < KP; k++) nk[k] = 0;
#pragma unroll
    for (int sub = 0; sub < 2; sub++) {
      const int h = h0 + sub * 32 + lane;
      const int j = h < H ? row[h] : N;
      const bool valid = j < N;
      if (!__any_sync(0xffffffffu, valid)) continue;
      float rx = 0.f, ry = 0.f, rz = 0.f;
      if (valid) {
        rx = s_pts[3 * (size_t)j] - qx;
        ry = s_pts[3 * (size_t)j + 1] - qy;
        rz = s_pts[3 * (size_t)j + 2] - qz;
        cnt += flags ? (int)flags[j] : 1;
      }
#pragma unroll
      for (int k = 0; k < KP; k++) {
        const float dx = rx - kp.v[3 * k], dy = ry - kp.v[3 * k + 1], dz = rz - kp.v[3 * k + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        const bool in = valid && d2 < sig2;
        const unsigned mask = __ballot_sync(0xffffffffu, in);
        if (in) {
          const float d = d2 * rsqrtf(fmaxf(d2, 1e-30f));
          ent[k][nk[k] + __popc(mask & lt)] = make_float2(__int_as_float(j), fmaxf(fmaf(-d, inv_sigma, 1.f), 0.f));
        }
        nk[k] += __popc(mask);
      }
    }
    if (lane < UN) {
#pragma unroll
      for (int k = 0; k < KP; k++) ent[k][nk[k] + lane] = make_float2(__int_as_float(0), 0.f);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < KP; k++) {
      for (int e = 0; e < nk[k]; e += UN) {
        float2 en[UN];
#pragma unroll
        for (int u = 0; u < UN; u += 2) {
          const float4 t = *reinterpret_cast<const float4*>(&ent[k][e + u]);
          en[u] = make_float2(t.x, t.y);
          en[u + 1] = make_float2(t.z, t.w);
        }
        float fv[UN][CPL];
#pragma unroll
        for (int u = 0; u < UN; u++) {
          const float* f = s_feats + (size_t)__float_as_int(en[u].x) * C + lane * VEC;
#pragma unroll
          for (int v = 0; v < NV; v++) {
            if (VEC == 4) {
              const float4 t = *reinterpret_cast<const float4*>(f + v * 128);
              fv[u][4 * v] = t.x; fv[u][4 * v + 1] = t.y; fv[u][4 * v + 2] = t.z; fv[u][4 * v + 3] = t.w;
            } else if (VEC == 2) {
              const float2 t = *reinterpret_cast<const float2*>(f);
              fv[u][0] = t.x; fv[u][1] = t.y;
            } else {
              fv[u][0] = f[0];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UN; u++)
#pragma unroll
          for (int i = 0; i < CPL; i++) acc[k][i] = fmaf(en[u].y, fv[u][i], acc[k][i]);
      }
    }
    __syncwarp();
  }
  cnt = lcr_warp_sum(cnt);
  float* o = wf + (size_t)m * (KP * C) + lane * VEC;
#pragma unroll
  for (int k = 0; k < KP; k++)
#pragma unroll
    for (int v = 0; v < NV; v++) {
      if (VEC == 4)
        *reinterpret_cast<float4*>(o + k * C + v * 128) =
            make_float4(acc[k][4 * v], acc[k][4 * v + 1], acc[k][4 * v + 2], acc[k][4 * v + 3]);
      else if (VEC == 2)
        *reinterpret_cast<float2*>(o + k * C) = make_float2(acc[k][0], acc[k][1]);
      else
        o[k * C] = acc[k][0];
    }
  if (lane == 0) rowscale[m] = 1.f / (float)max(cnt, 1);
}

// ------------------------------------------------------------------ KPConv gather, mask-dispatch form
// Same register layout as the dense kernel (lane owns C/32 channels, 15 x C/32 accumulators), but the
// accumulation visits only the NON-ZERO influences: the influence pass leaves, per neighbour, a 15-bit mask
// of the kernel points that reach it (~1.6 of 15) and drops neighbours no kernel point reaches; the
// accumulation walks the set bits (warp-uniform), and a 15-way switch (one indirect branch, no divergence)
// selects the statically-indexed accumulator row.  Per neighbour: 1 feature-row load + ~1.6 x (bit scan,
// weight broadcast, branch, C/32 FFMA) instead of 15 x C/32 FFMA (or 15 compare-and-skip pairs).
template <int CPL>
__global__ void __launch_bounds__(128)
kpconv_gather_mask_kernel(const float* __restrict__ s_feats, const float* __restrict__ q_pts,
                          const float* __restrict__ s_pts, const int32_t* __restrict__ idx, int ld_idx, int H,
                          const KpArg kp, float sigma, const uint8_t* __restrict__ flags, int M, int N,
                          float* __restrict__ wf, float* __restrict__ rowscale) {
  constexpr int C = 32 * CPL;
  __shared__ float s_w[4][KP][32];
  __shared__ int s_j[4][32];
  __shared__ unsigned s_m[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * 4 + warp;
  if (m >= M) return;
  const float qx = q_pts[3 * m], qy = q_pts[3 * m + 1], qz = q_pts[3 * m + 2];
  const float inv_sigma = 1.f / sigma;
  float acc[KP][CPL];
#pragma unroll
  for (int k = 0; k < KP; k++)
#pragma unroll
    for (int i = 0; i < CPL; i++) acc[k][i] = 0.f;
  int cnt = 0;
  const int32_t* row = idx + (size_t)m * ld_idx;
  for (int h0 = 0; h0 < H; h0 += 32) {
    const int h = h0 + lane;
    const int j = h < H ? row[h] : N;
    const bool valid = j < N;
    float w[KP];
    unsigned mask = 0;
    if (valid) {
      const float rx = s_pts[3 * (size_t)j] - qx, ry = s_pts[3 * (size_t)j + 1] - qy,
                  rz = s_pts[3 * (size_t)j + 2] - qz;
#pragma unroll
      for (int k = 0; k < KP; k++) {
        const float dx = rx - kp.v[3 * k], dy = ry - kp.v[3 * k + 1], dz = rz - kp.v[3 * k + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        w[k] = fmaxf(fmaf(-d2 * rsqrtf(fmaxf(d2, 1e-30f)), inv_sigma, 1.f), 0.f);
        mask |= (w[k] > 0.f ? 1u : 0u) << k;
      }
      cnt += flags ? (int)flags[j] : 1;
    }
    const unsigned vmask = __ballot_sync(0xffffffffu, mask != 0);
    const int slot = __popc(vmask & ((1u << lane) - 1u));
    const int nv = __popc(vmask);
    if (mask != 0) {
#pragma unroll
      for (int k = 0; k < KP; k++) s_w[warp][k][slot] = w[k];
      s_j[warp][slot] = j;
      s_m[warp][slot] = mask;
    }
    __syncwarp();
    constexpr int UN = CPL >= 4 ? 2 : 4;  // neighbours whose feature rows are in flight together
#define LCR_ACC_CASE(K)                                                         \
  case K:                                                                       \
    _Pragma("unroll") for (int i = 0; i < CPL; i++) acc[K][i] = fmaf(wk, f[i], acc[K][i]); \
    break;
    auto accumulate = [&](const float (&f)[CPL], int hh) {
      unsigned mm = s_m[warp][hh];
      while (mm) {
        const int k = __ffs(mm) - 1;
        mm &= mm - 1;
        const float wk = s_w[warp][k][hh];
        switch (k) {
          LCR_ACC_CASE(0) LCR_ACC_CASE(1) LCR_ACC_CASE(2) LCR_ACC_CASE(3) LCR_ACC_CASE(4) LCR_ACC_CASE(5)
          LCR_ACC_CASE(6) LCR_ACC_CASE(7) LCR_ACC_CASE(8) LCR_ACC_CASE(9) LCR_ACC_CASE(10) LCR_ACC_CASE(11)
          LCR_ACC_CASE(12) LCR_ACC_CASE(13) LCR_ACC_CASE(14)
          default: break;
        }
      }
    };
#undef LCR_ACC_CASE
    int hh = 0;
    for (; hh + UN <= nv; hh += UN) {
      float fv[UN][CPL];
#pragma unroll
      for (int u = 0; u < UN; u++) {
        const float* f = s_feats + (size_t)s_j[warp][hh + u] * C + lane;
#pragma unroll
        for (int i = 0; i < CPL; i++) fv[u][i] = f[32 * i];
      }
#pragma unroll
      for (int u = 0; u < UN; u++) accumulate(fv[u], hh + u);
    }
    for (; hh < nv; hh++) {
      const float* f = s_feats + (size_t)s_j[warp][hh] * C + lane;
      float fv[CPL];
#pragma unroll
      for (int i = 0; i < CPL; i++) fv[i] = f[32 * i];
      accumulate(fv, hh);
    }
    __syncwarp();
  }
  cnt = lcr_warp_sum(cnt);
  float* o = wf + (size_t)m * (KP * C) + lane;
#pragma unroll
  for (int k = 0; k < KP; k++)
#pragma unroll
    for (int i = 0; i < CPL; i++) o[k * C + 32 * i] = acc[k][i];
  if (lane == 0) rowscale[m] = 1.f / (float)max(cnt, 1);
}

// C_in = 1 (encoder1_1): lanes over neighbours, 15 warp-reduced sums, then O outputs per query.
template <bool FAST>
__global__ void __launch_bounds__(128)
kpconv_c1_kernel(const float* __restrict__ s_feats, const float* __restrict__ q_pts, const float* __restrict__ s_pts,
                 const int32_t* __restrict__ idx, int ld_idx, int H, const float* __restrict__ kpts, const KpArg kp,
                 float sigma, const float* __restrict__ weights /*[15,1,O]*/, const float* __restrict__ bias, int O,
                 int M, int N, float* __restrict__ out) {
  __shared__ float s_kp[KP * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (!FAST) {
    if (threadIdx.x < KP * 3) s_kp[threadIdx.x] = kpts[threadIdx.x];
    __syncthreads();
  }
  const int m = blockIdx.x * 4 + warp;
  if (m >= M) return;
  const float qx = q_pts[3 * m], qy = q_pts[3 * m + 1], qz = q_pts[3 * m + 2];
  const float inv_sigma = 1.f / sigma;
  float ws[KP];
#pragma unroll
  for (int k = 0; k < KP; k++) ws[k] = 0.f;
  int cnt = 0;
  const int32_t* row = idx + (size_t)m * ld_idx;
  for (int h = lane; h < H; h += 32) {
    const int j = row[h];
    if (j < N) {
      const float f = s_feats[j];
      const float rx = s_pts[3 * (size_t)j] - qx, ry = s_pts[3 * (size_t)j + 1] - qy,
                  rz = s_pts[3 * (size_t)j + 2] - qz;
#pragma unroll
      for (int k = 0; k < KP; k++) {
        if (FAST) {
          const float dx = rx - kp.v[3 * k], dy = ry - kp.v[3 * k + 1], dz = rz - kp.v[3 * k + 2];
          const float d2 = dx * dx + dy * dy + dz * dz;
          ws[k] = fmaf(fmaxf(fmaf(-d2 * rsqrtf(fmaxf(d2, 1e-30f)), inv_sigma, 1.f), 0.f), f, ws[k]);
        } else {
          const float dx = rx - s_kp[3 * k], dy = ry - s_kp[3 * k + 1], dz = rz - s_kp[3 * k + 2];
          const float d2 = dx * dx + dy * dy + dz * dz;
          ws[k] = fmaf(fmaxf(1.f - __fdiv_rn(sqrtf(d2), sigma), 0.f), f, ws[k]);
        }
      }
      cnt += f > 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < KP; k++) ws[k] = lcr_warp_sum(ws[k]);
  cnt = lcr_warp_sum(cnt);
  const float num = (float)max(cnt, 1);
  for (int o = lane; o < O; o += 32) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < KP; k++) a = fmaf(ws[k], weights[k * O + o], a);
    a = __fdiv_rn(a, num);
    out[(size_t)m * O + o] = bias ? a + bias[o] : a;
  }
}

// ------------------------------------------------------------------ GroupNorm statistics
// partial[(s * n_chunks + chunk) * G + g] = (sum, sumsq) over rows of the chunk x channels of g.
// A chunk is gn_chunk_rows(C) rows (>= 64 KB of x per block for every C).
constexpr int kGnRows = 64;   // smallest chunk (workspace sizing)
static inline int gn_chunk_rows(int C) { return C >= 256 ? 64 : 64 * (256 / C); }

__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ x, int C, int G, const int64_t* __restrict__ stack_off, int n_chunks,
                  int chunk_rows, double2* __restrict__ partial) {
  // 16-byte loads: a thread owns one quad of columns and every (256 / (C/4))-th row of the chunk;
  // its per-column partial sums go through shared memory, then one thread per group adds its
  // columns in a fixed order (deterministic; no float atomics)
  __shared__ float s_sum[1024], s_sq[1024];
  const int s = blockIdx.y, chunk = blockIdx.x;
  const int64_t r0 = stack_off[s] + (int64_t)chunk * chunk_rows;
  const int64_t r1 = min(r0 + chunk_rows, stack_off[s + 1]);
  const int quads = C >> 2;                       // <= 256
  const int cq = threadIdx.x % quads, rl = threadIdx.x / quads, nrl = 256 / quads;
  if (rl < nrl) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    const float* xp = x + 4 * cq;
#pragma unroll 4
    for (int64_t r = r0 + rl; r < r1; r += nrl) {
      const float4 v = *reinterpret_cast<const float4*>(xp + r * C);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      b.x = fmaf(v.x, v.x, b.x); b.y = fmaf(v.y, v.y, b.y); b.z = fmaf(v.z, v.z, b.z); b.w = fmaf(v.w, v.w, b.w);
    }
    *reinterpret_cast<float4*>(s_sum + rl * C + 4 * cq) = a;
    *reinterpret_cast<float4*>(s_sq + rl * C + 4 * cq) = b;
  }
  __syncthreads();
  if (threadIdx.x < G) {
    const int cpg = C / G;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < nrl; r++)
      for (int cc = 0; cc < cpg; cc++) {
        a += (double)s_sum[r * C + threadIdx.x * cpg + cc];
        b += (double)s_sq[r * C + threadIdx.x * cpg + cc];
      }
    partial[((size_t)s * n_chunks + chunk) * G + threadIdx.x] = make_double2(a, b);
  }
}

// one warp per (stack, group): fixed-order reduction over chunks -> mean, rstd
__global__ void gn_finalize_kernel(const double2* __restrict__ partial, int G, int C,
                                   const int64_t* __restrict__ stack_off, int n_chunks, int chunk_rows, int S,
                                   float eps, float2* __restrict__ stats) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= S * G) return;
  const int s = w / G, g = w % G;
  const int64_t rows = stack_off[s + 1] - stack_off[s];
  const int used = (int)((rows + chunk_rows - 1) / chunk_rows);
  double a = 0.0, b = 0.0;
  for (int ch = lane; ch < used; ch += 32) {
    const double2 p = partial[((size_t)s * n_chunks + ch) * G + g];
    a += p.x;
    b += p.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    const double n = (double)rows * (double)(C / G);
    const double mean = n > 0 ? a / n : 0.0;
    const double var = n > 0 ? fmax(b / n - mean * mean, 0.0) : 0.0;
    stats[w] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

// Statistics fused into the GEMM epilogue (gemm_tc.cu, GnFuse): one 128-thread CTA per (stack, group)
// adds the per-32-row-block column partials of its stack in a fixed order (thread t takes blocks
// b0 + t, b0 + t + 128, ...; then a fixed shuffle / shared-memory tree).  Block b covers rows
// [32b, 32b+32); slot 0 holds the rows of the block's first stack, slot 1 the rest (only the block that
// contains a stack's first row can start in the previous stack: every stack has >= 32 rows).
constexpr int kGnFinThreads = 128;
__global__ void __launch_bounds__(kGnFinThreads)
gn_finalize_blocks_kernel(const float2* __restrict__ partial, int G, int C, const int64_t* __restrict__ stack_off,
                          int S, float eps, float2* __restrict__ stats) {
  __shared__ double s_a[kGnFinThreads / 32], s_b[kGnFinThreads / 32];
  const int w = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = w / G, g = w % G, cpg = C / G;
  const int64_t r0 = stack_off[s], r1 = stack_off[s + 1];
  double a = 0.0, b = 0.0;
  if (r1 > r0) {
    const int64_t b0 = r0 >> 5, b1 = (r1 - 1) >> 5;
    for (int64_t blk = b0 + threadIdx.x; blk <= b1; blk += kGnFinThreads) {
      const int slot = (blk << 5) < r0 ? 1 : 0;
      const float2* p = partial + ((size_t)blk * 2 + slot) * C + g * cpg;
      double aa = 0.0, bb = 0.0;
#pragma unroll 4
      for (int cc = 0; cc < cpg; cc++) {
        const float2 v = p[cc];
        aa += (double)v.x;
        bb += (double)v.y;
      }
      a += aa;
      b += bb;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    s_a[warp] = a;
    s_b[warp] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = (s_a[0] + s_a[1]) + (s_a[2] + s_a[3]);
    b = (s_b[0] + s_b[1]) + (s_b[2] + s_b[3]);
    const double n = (double)(r1 - r0) * (double)cpg;
    const double mean = n > 0 ? a / n : 0.0;
    const double var = n > 0 ? fmax(b / n - mean * mean, 0.0) : 0.0;
    stats[w] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
  }
}

// ------------------------------------------------------------------ GroupNorm apply (+add, +LeakyReLU, +row flags)
// y = act( gn(x; stats, gamma, beta) + other ),  other = none | raw tensor | gn(x2; stats2, gamma2, beta2)
struct GnApplyArgs {
  const float* x; const float2* stats; const float* gamma; const float* beta;
  const float* x2; const float2* stats2; const float* gamma2; const float* beta2;
  float* y; uint8_t* flags;
  const int64_t* stack_off; int S; int64_t rows; int C; int G; float slope; int act;
};

__global__ void __launch_bounds__(256) gn_apply_kernel(GnApplyArgs a) {
  const int lpr = min(32, a.C / 4);          // lanes per row
  const int rpw = 32 / lpr;                  // rows per warp
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t r = warp * rpw + lane / lpr;
  const int sub = lane % lpr;
  const bool active = r < a.rows;
  float rowsum = 0.f;
  if (active) {
    const int s = lcr_find_segment(a.stack_off, a.S, r);
    const int cpg = a.C / a.G;
    for (int c = sub * 4; c < a.C; c += lpr * 4) {
      float4 v = *reinterpret_cast<const float4*>(a.x + r * a.C + c);
      float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 st = a.stats[s * a.G + (c + i) / cpg];
        o[i] = (o[i] - st.x) * st.y * a.gamma[c + i] + a.beta[c + i];
      }
      if (a.x2) {
        const float4 w = *reinterpret_cast<const float4*>(a.x2 + r * a.C + c);
        float p[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          if (a.stats2) {
            const float2 st = a.stats2[s * a.G + (c + i) / cpg];
            p[i] = (p[i] - st.x) * st.y * a.gamma2[c + i] + a.beta2[c + i];
          }
          o[i] += p[i];
        }
      }
      if (a.act) {
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = o[i] > 0.f ? o[i] : o[i] * a.slope;
      }
      rowsum += (o[0] + o[1]) + (o[2] + o[3]);
      *reinterpret_cast<float4*>(a.y + r * a.C + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  if (a.flags) {
    for (int o = lpr >> 1; o > 0; o >>= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, o);
    if (active && sub == 0) a.flags[r] = rowsum > 0.f;
  }
}

// Streaming variant (every call except row flags on C > 128): a lane keeps ONE quad of columns for
// kGnIter consecutive row groups, so gamma/beta and the per-stack scale/shift
//   y = x * (rstd * gamma) + (beta - mean * rstd * gamma)
// live in registers and are recomputed only when the rows cross into the next stack; the inner
// loop is one 16-byte load, 4 FFMA (+4 for the normalised shortcut), the activation and one store.
constexpr int kGnIter = 8;

struct GnAffine {
  float sc[4], sh[4];
  __device__ __forceinline__ void set(const float2* __restrict__ stats, int s, int G, int cpg, int c, float4 g, float4 b) {
    const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 st = stats[s * G + (c + i) / cpg];
      sc[i] = st.y * gg[i];
      sh[i] = fmaf(-st.x, sc[i], bb[i]);
    }
  }
};

template <int MODE>  // 0: no other, 1: + raw other, 2: + normalised other
__global__ void __launch_bounds__(256) gn_apply_stream_kernel(GnApplyArgs a) {
  const int seg_lanes = min(32, a.C >> 2);       // lanes covering one segment of min(C, 128) channels
  const int nseg = a.C > 128 ? a.C >> 7 : 1;
  const int rpw = 32 / seg_lanes;                // rows per warp iteration
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int seg = (int)(warp % nseg);
  const int64_t rb = warp / nseg;
  const int sub = lane % seg_lanes, rsub = lane / seg_lanes;
  const int c = seg * 128 + sub * 4;
  const int cpg = a.C / a.G;
  int64_t r = rb * (rpw * kGnIter) + rsub;
  if (r >= a.rows) {
    if (a.flags == nullptr) return;
    r = a.rows;                                   // stay for the shuffles below
  }
  const float4 g1 = *reinterpret_cast<const float4*>(a.gamma + c), b1 = *reinterpret_cast<const float4*>(a.beta + c);
  float4 g2 = g1, b2 = b1;
  if (MODE == 2) {
    g2 = *reinterpret_cast<const float4*>(a.gamma2 + c);
    b2 = *reinterpret_cast<const float4*>(a.beta2 + c);
  }
  int s = r < a.rows ? lcr_find_segment(a.stack_off, a.S, r) : a.S - 1;
  int64_t next = a.stack_off[s + 1];
  GnAffine f1, f2;
  f1.set(a.stats, s, a.G, cpg, c, g1, b1);
  if (MODE == 2) f2.set(a.stats2, s, a.G, cpg, c, g2, b2);
#pragma unroll 2
  for (int it = 0; it < kGnIter; it++, r += rpw) {
    const bool active = r < a.rows;
    float rowsum = 0.f;
    if (active) {
      if (r >= next) {
        while (r >= a.stack_off[s + 1]) s++;
        next = a.stack_off[s + 1];
        f1.set(a.stats, s, a.G, cpg, c, g1, b1);
        if (MODE == 2) f2.set(a.stats2, s, a.G, cpg, c, g2, b2);
      }
      const float4 v = *reinterpret_cast<const float4*>(a.x + r * a.C + c);
      float o[4] = {fmaf(v.x, f1.sc[0], f1.sh[0]), fmaf(v.y, f1.sc[1], f1.sh[1]), fmaf(v.z, f1.sc[2], f1.sh[2]),
                    fmaf(v.w, f1.sc[3], f1.sh[3])};
      if (MODE >= 1) {
        const float4 w = *reinterpret_cast<const float4*>(a.x2 + r * a.C + c);
        if (MODE == 2) {
          o[0] += fmaf(w.x, f2.sc[0], f2.sh[0]); o[1] += fmaf(w.y, f2.sc[1], f2.sh[1]);
          o[2] += fmaf(w.z, f2.sc[2], f2.sh[2]); o[3] += fmaf(w.w, f2.sc[3], f2.sh[3]);
        } else {
          o[0] += w.x; o[1] += w.y; o[2] += w.z; o[3] += w.w;
        }
      }
      if (a.act) {
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = o[i] > 0.f ? o[i] : o[i] * a.slope;
      }
      rowsum = (o[0] + o[1]) + (o[2] + o[3]);
      *reinterpret_cast<float4*>(a.y + r * a.C + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (a.flags) {   // only launched with nseg == 1: the row lives inside the warp
      for (int o = seg_lanes >> 1; o > 0; o >>= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, o);
      if (active && sub == 0) a.flags[r] = rowsum > 0.f;
    }
  }
}

// ------------------------------------------------------------------ max-pool over neighbours
__global__ void __launch_bounds__(256)
maxpool_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int ld_idx, int H, int M, int N, int C,
               float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  const int32_t* row = idx + (size_t)m * ld_idx;
  // the 32 lanes fetch 32 neighbour indices at once (one coalesced load) and hand them round by
  // shuffle, so four independent 16-byte feature loads are in flight per lane
  for (int c0 = 0; c0 < C; c0 += 128) {
    const int c = c0 + lane * 4;
    const bool on = c < C;
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int h0 = 0; h0 < H; h0 += 32) {
      const int jl = h0 + lane < H ? row[h0 + lane] : -1;
      const int n = min(32, H - h0);
      int hh = 0;
      for (; hh + 4 <= n; hh += 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int j = __shfl_sync(0xffffffffu, jl, hh + u);
          // the pad row of the reference is a row of zeros appended to x (functional.py:64)
          v[u] = (on && j < N) ? *reinterpret_cast<const float4*>(x + (size_t)j * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          best.x = fmaxf(best.x, v[u].x); best.y = fmaxf(best.y, v[u].y);
          best.z = fmaxf(best.z, v[u].z); best.w = fmaxf(best.w, v[u].w);
        }
      }
      for (; hh < n; hh++) {
        const int j = __shfl_sync(0xffffffffu, jl, hh);
        const float4 v = (on && j < N) ? *reinterpret_cast<const float4*>(x + (size_t)j * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        best.x = fmaxf(best.x, v.x); best.y = fmaxf(best.y, v.y);
        best.z = fmaxf(best.z, v.z); best.w = fmaxf(best.w, v.w);
      }
    }
    if (on) *reinterpret_cast<float4*>(out + (size_t)m * C + c) = best;
  }
}

__global__ void rowflag_kernel(const float* __restrict__ x, int64_t rows, int C, uint8_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += x[r * C + c];
  s = lcr_warp_sum(s);
  if (lane == 0) flags[r] = s > 0.f;
}

}  // namespace

// gather variant when host kernel points are given: 0 exact dense loop (IEEE sqrt/divide influences),
// 1 fast dense loop, 2 sparse influence lists, 3 auto (default).  LCR_GATHER={exact,dense,sparse,auto}
// sets the default; lcr_set_gather_mode overrides it.
static int g_gather_mode = -1;
static int gather_mode() {
  if (g_gather_mode < 0) {
    const char* e = getenv("LCR_GATHER");
    g_gather_mode = !e ? 3 : !strcmp(e, "exact") ? 0 : !strcmp(e, "dense") ? 1 : !strcmp(e, "sparse") ? 2 : !strcmp(e, "mask") ? 4 : !strcmp(e, "packed") ? 5 : !strcmp(e, "mma") ? 6 : !strcmp(e, "group") ? 7 : 3;
  }
  return g_gather_mode;
}
extern "C" void lcr_set_gather_mode(int mode) { g_gather_mode = mode < 0 || mode > 7 ? 3 : mode; }

// ================================================================== C ABI
extern "C" size_t lcr_kpconv_ws_bytes(int64_t m_rows, int c_in) {
  return lcr_align_up((size_t)m_rows * KP * c_in * sizeof(float)) + lcr_align_up((size_t)m_rows * sizeof(float)) + 256;
}
// with room for the packed (x, y, z, flag) copy of the n_support points the fast gather reads
extern "C" size_t lcr_kpconv_ws_bytes2(int64_t m_rows, int64_t n_support, int c_in) {
  return lcr_kpconv_ws_bytes(m_rows, c_in) + lcr_align_up((size_t)n_support * sizeof(float4));
}

static int kpconv_impl(const float* s_feats, const uint8_t* s_flags, int64_t n_support, const float* q_points,
                       int64_t m_query, const float* s_points, const int32_t* idx, int ld_idx, int H,
                       const float* kernel_points, const float* kernel_points_host, float sigma,
                       const float* weights, const float* weights_nk, const float* weights_nk_lo,
                       const float* bias, int c_in, int c_out, float* out, void* ws, size_t ws_bytes,
                       float* gn_partial, const int64_t* stack_off, int n_stacks, cudaStream_t stream) {
  LCR_REQUIRE(m_query >= 0 && n_support >= 0 && m_query < (1ll << 31) && n_support < (1ll << 31), "kpconv: sizes");
  LCR_REQUIRE(H >= 1 && ld_idx >= H, "kpconv: bad neighbour table width");
  LCR_REQUIRE(c_in == 1 || c_in == 32 || c_in == 64 || c_in == 128 || c_in == 256,
              "kpconv: c_in must be 1, 32, 64, 128 or 256");
  LCR_REQUIRE(c_out % 32 == 0, "kpconv: c_out must be a multiple of 32");
  if (m_query == 0) return LCR_OK;
  const int M = (int)m_query, N = (int)n_support;
  const unsigned grid = (unsigned)((M + 3) / 4);
  if (c_in == 1) {
    LcrProfScope prof("kpconv_c1", 2.0 * M * KP * (H + c_out), 4.0 * M * H + 16.0 * (M + N) + 4.0 * M * c_out, stream);
    KpArg kp;
    const bool fast = kernel_points_host != nullptr && gather_mode() != 0;
    if (fast) memcpy(kp.v, kernel_points_host, sizeof(kp.v));
    else memset(kp.v, 0, sizeof(kp.v));
    if (fast)
      kpconv_c1_kernel<true><<<grid, 128, 0, stream>>>(s_feats, q_points, s_points, idx, ld_idx, H, kernel_points, kp,
                                                       sigma, weights, bias, c_out, M, N, out);
    else
      kpconv_c1_kernel<false><<<grid, 128, 0, stream>>>(s_feats, q_points, s_points, idx, ld_idx, H, kernel_points, kp,
                                                        sigma, weights, bias, c_out, M, N, out);
    LCR_LAUNCHED(1);
    LCR_CUDA_CHECK_LAUNCH();
    return LCR_OK;
  }
  LCR_REQUIRE(ws && ws_bytes >= lcr_kpconv_ws_bytes(m_query, c_in), "kpconv: workspace too small");
  LcrArena a(ws, ws_bytes);
  float* wf = a.take<float>((size_t)M * KP * c_in);
  float* rowscale = a.take<float>(M);
  float4* pts4 = a.take<float4>((size_t)N);
  const bool have_pts4 = a.ok();      // callers that sized the workspace with lcr_kpconv_ws_bytes2
  {
  LcrProfScope prof("kpconv_gather", 2.0 * M * KP * (double)H * c_in,
                    4.0 * M * H + 4.0 * (double)N * c_in + 12.0 * (M + N) + 4.0 * (double)M * KP * c_in, stream);
  // variant (see gather_mode): exact dense / fast dense / sparse lists
  int mode = kernel_points_host == nullptr ? 0 : gather_mode();
  // auto, measured on B200 (scripts/bench_gather.py): the warp-MMA kernel wins for C = 32 (0.67 vs 0.79 ms on 465 k
  // queries), ties at C = 64 and loses above (register pressure); the fast dense loop wins or ties over the sparse /
  // mask / packed-FMA forms for every width
  if (mode == 3) mode = c_in == 32 ? 6 : 1;
  if ((mode == 1 || mode == 6 || mode == 7) && !have_pts4) mode = 5;   // no room for the packed points: the packed-FMA loop reads the 12-byte points
  if (mode == 1 || mode == 6 || mode == 7) {
    pack_points_kernel<<<(N + 255) / 256, 256, 0, stream>>>(s_points, s_flags, N, pts4);
    LCR_LAUNCHED(1);
  }
  KpArg kp;
  if (mode != 0) memcpy(kp.v, kernel_points_host, sizeof(kp.v));
  else memset(kp.v, 0, sizeof(kp.v));
#define LCR_GATHER(CPL)                                                                                             \
  if (mode == 0)                                                                                                    \
    kpconv_gather_kernel<CPL, false><<<grid, 128, 0, stream>>>(s_feats, q_points, s_points, idx, ld_idx, H,         \
                                                               kernel_points, kp, sigma, s_flags, nullptr, M, N, wf, \
                                                               rowscale);                                           \
  else if (mode == 1)                                                                                               \
    kpconv_gather_kernel<CPL, true><<<grid, 128, 0, stream>>>(s_feats, q_points, s_points, idx, ld_idx, H,          \
                                                              kernel_points, kp, sigma, s_flags, pts4, M, N, wf,    \
                                                              rowscale);                                            \
  else if (mode == 7)                                                                                               \
    kpconv_gather_kernel<CPL, true, true><<<grid, 128, 0, stream>>>(s_feats, q_points, s_points, idx, ld_idx, H,    \
                                                                    kernel_points, kp, sigma, s_flags, pts4, M, N,  \
                                                                    wf, rowscale);                                  \
  else if (mode == 6)                                                                                               \
    kpconv_gather_mma_kernel<CPL><<<grid, 128, 0, stream>>>(s_feats, q_points, pts4, idx, ld_idx, H, kernel_points, \
                                                            sigma, M, N, wf, rowscale);                             \
  else if (mode == 5)                                                                                               \
    kpconv_gather_packed_kernel<CPL><<<grid, 128, 0, stream>>>(s_feats, q_points, s_points, idx, ld_idx, H, kp,     \
                                                               sigma, s_flags, M, N, wf, rowscale);                 \
  else if (mode == 4)                                                                                               \
    kpconv_gather_mask_kernel<CPL><<<grid, 128, 0, stream>>>(s_feats, q_points, s_points, idx, ld_idx, H, kp,       \
                                                             sigma, s_flags, M, N, wf, rowscale);                   \
  else                                                                                                              \
    kpconv_gather_sparse_kernel<CPL><<<grid, 128, 0, stream>>>(s_feats, q_points, s_points, idx, ld_idx, H, kp,     \
                                                               sigma, s_flags, M, N, wf, rowscale)
  switch (c_in) {
    case 32: LCR_GATHER(1); break;
    case 64: LCR_GATHER(2); break;
    case 128: LCR_GATHER(4); break;
    default: LCR_GATHER(8); break;
  }
#undef LCR_GATHER
  }
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  if (weights_nk)  // tensor-core contraction: weights as [c_out, 15 * c_in]
    return lcr_gemm_tf32x3_gn(wf, KP * c_in, weights_nk, weights_nk_lo, KP * c_in, out, c_out, M, c_out, KP * c_in,
                              rowscale, bias, 0, gn_partial, stack_off, n_stacks, stream);
  LCR_REQUIRE(!gn_partial, "kpconv: fused GroupNorm statistics need the tensor-core contraction (weights_nk)");
  return lcr_gemm_f32(wf, KP * c_in, weights, c_out, out, c_out, M, c_out, KP * c_in, rowscale, bias, stream);
}

extern "C" int lcr_kpconv(const float* s_feats, const uint8_t* s_flags, int64_t n_support, const float* q_points,
                          int64_t m_query, const float* s_points, const int32_t* idx, int ld_idx, int H,
                          const float* kernel_points, const float* kernel_points_host, float sigma,
                          const float* weights, const float* weights_nk, const float* weights_nk_lo,
                          const float* bias, int c_in, int c_out, float* out, void* ws, size_t ws_bytes,
                          void* stream) {
  return kpconv_impl(s_feats, s_flags, n_support, q_points, m_query, s_points, idx, ld_idx, H, kernel_points,
                     kernel_points_host, sigma, weights, weights_nk, weights_nk_lo, bias, c_in, c_out, out, ws, ws_bytes,
                     nullptr, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int lcr_kpconv_gn(const float* s_feats, const uint8_t* s_flags, int64_t n_support, const float* q_points,
                             int64_t m_query, const float* s_points, const int32_t* idx, int ld_idx, int H,
                             const float* kernel_points, const float* kernel_points_host, float sigma,
                             const float* weights, const float* weights_nk, const float* weights_nk_lo,
                             const float* bias, int c_in, int c_out, float* out, void* ws, size_t ws_bytes,
                             float* gn_partial, const int64_t* stack_off, int n_stacks, void* stream) {
  LCR_REQUIRE(c_in > 1 && weights_nk && gn_partial && stack_off && n_stacks >= 1,
              "kpconv_gn: needs c_in > 1, the [c_out, 15 c_in] weights and the partial buffer");
  return kpconv_impl(s_feats, s_flags, n_support, q_points, m_query, s_points, idx, ld_idx, H, kernel_points,
                     kernel_points_host, sigma, weights, weights_nk, weights_nk_lo, bias, c_in, c_out, out, ws, ws_bytes,
                     gn_partial, stack_off, n_stacks, (cudaStream_t)stream);
}

extern "C" size_t lcr_gn_blocks_ws_bytes(int64_t rows, int channels) {
  return lcr_align_up((size_t)((rows + 31) / 32) * 2 * (size_t)channels * sizeof(float2)) + 256;
}

extern "C" int lcr_group_norm_finalize_blocks(const float* gn_partial, int64_t rows, int channels, int groups,
                                              const int64_t* stack_off, int n_stacks, float eps, float* stats,
                                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(gn_partial && stack_off && stats, "group_norm_finalize_blocks: null argument");
  LCR_REQUIRE(groups >= 1 && channels % groups == 0 && n_stacks >= 1 && rows >= 0, "group_norm_finalize_blocks: shape");
  LcrProfScope prof("group_norm_stats", 0.0, 16.0 * (double)((rows + 31) / 32) * channels, stream);
  gn_finalize_blocks_kernel<<<n_stacks * groups, kGnFinThreads, 0, stream>>>(
      reinterpret_cast<const float2*>(gn_partial), groups, channels, stack_off, n_stacks, eps,
      reinterpret_cast<float2*>(stats));
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" size_t lcr_group_norm_ws_bytes(int64_t max_stack_rows, int n_stacks, int groups) {
  const size_t chunks = (size_t)((max_stack_rows + kGnRows - 1) / kGnRows) + 1;
  return lcr_align_up(chunks * n_stacks * groups * sizeof(double2)) + 256;
}

extern "C" int lcr_group_norm_stats(const float* x, int64_t rows, int channels, int groups, const int64_t* stack_off,
                                    int n_stacks, int64_t max_stack_rows, float eps, float* stats_out, void* ws,
                                    size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(channels % groups == 0 && groups <= 32 && channels % 4 == 0, "group_norm: bad channel/group count");
  LCR_REQUIRE(channels <= 256 ? (256 % channels == 0) : (channels % 256 == 0 && channels <= 1024),
              "group_norm: channels must divide 256 or be a multiple of 256 (<= 1024)");
  LCR_REQUIRE(n_stacks >= 1 && ws && ws_bytes >= lcr_group_norm_ws_bytes(max_stack_rows, n_stacks, groups),
              "group_norm: workspace too small");
  if (rows == 0) return LCR_OK;
  const int chunk_rows = gn_chunk_rows(channels);
  const int n_chunks = (int)((max_stack_rows + chunk_rows - 1) / chunk_rows) + 1;
  double2* partial = (double2*)ws;
  LcrProfScope prof("group_norm_stats", 3.0 * rows * channels, 4.0 * rows * channels, stream);
  dim3 grid(n_chunks, n_stacks);
  gn_partial_kernel<<<grid, 256, 0, stream>>>(x, channels, groups, stack_off, n_chunks, chunk_rows, partial);
  const int warps = n_stacks * groups;
  gn_finalize_kernel<<<(warps * 32 + 255) / 256, 256, 0, stream>>>(partial, groups, channels, stack_off, n_chunks,
                                                                   chunk_rows, n_stacks, eps, (float2*)stats_out);
  LCR_LAUNCHED(2);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_group_norm_apply(const float* x, const float* stats, const float* gamma, const float* beta,
                                    const float* x2, const float* stats2, const float* gamma2, const float* beta2,
                                    int64_t rows, int channels, int groups, const int64_t* stack_off, int n_stacks,
                                    int leaky, float slope, float* y, uint8_t* row_flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(channels % 32 == 0 && channels % groups == 0, "group_norm_apply: bad channel count");
  if (rows == 0) return LCR_OK;
  GnApplyArgs a;
  a.x = x; a.stats = (const float2*)stats; a.gamma = gamma; a.beta = beta;
  a.x2 = x2; a.stats2 = (const float2*)stats2; a.gamma2 = gamma2; a.beta2 = beta2;
  a.y = y; a.flags = row_flags; a.stack_off = stack_off; a.S = n_stacks; a.rows = rows; a.C = channels;
  a.G = groups; a.slope = slope; a.act = leaky;
  const int lpr = channels / 4 < 32 ? channels / 4 : 32;
  const int rpw = 32 / lpr;
  const int64_t warps = (rows + rpw - 1) / rpw;
  LcrProfScope prof("group_norm_apply", 6.0 * rows * channels, 4.0 * rows * channels * (x2 ? 3.0 : 2.0), stream);
  if (channels % 4 == 0 && (channels <= 128 || (channels % 128 == 0 && row_flags == nullptr))) {
    const int seg_lanes = channels / 4 < 32 ? channels / 4 : 32;
    const int nseg = channels > 128 ? channels / 128 : 1;
    const int64_t rows_per_warp = (32 / seg_lanes) * kGnIter;
    const int64_t nwarps = ((rows + rows_per_warp - 1) / rows_per_warp) * nseg;
    const unsigned grid = (unsigned)((nwarps * 32 + 255) / 256);
    if (!x2) gn_apply_stream_kernel<0><<<grid, 256, 0, stream>>>(a);
    else if (!stats2) gn_apply_stream_kernel<1><<<grid, 256, 0, stream>>>(a);
    else gn_apply_stream_kernel<2><<<grid, 256, 0, stream>>>(a);
  } else {
    gn_apply_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(a);
  }
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_maxpool(const float* x, int64_t n_support, const int32_t* idx, int ld_idx, int H, int64_t m_query,
                           int channels, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(channels % 4 == 0 && H >= 1 && ld_idx >= H, "maxpool: bad shape");
  if (m_query == 0) return LCR_OK;
  LcrProfScope prof("maxpool", (double)m_query * H * channels,
                    4.0 * m_query * H + 4.0 * (double)n_support * channels + 4.0 * (double)m_query * channels, stream);
  maxpool_kernel<<<(unsigned)((m_query * 32 + 255) / 256), 256, 0, stream>>>(x, idx, ld_idx, H, (int)m_query,
                                                                            (int)n_support, channels, out);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

extern "C" int lcr_row_flags(const float* x, int64_t rows, int channels, uint8_t* flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows == 0) return LCR_OK;
  rowflag_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, stream>>>(x, rows, channels, flags);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
