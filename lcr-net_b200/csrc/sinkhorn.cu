// sinkhorn.cu -- a11: Sinkhorn with a learnable dustbin (reference: experiments/lcrnet/modules/sinkhorn/
// learnable_sinkhorn.py:13-66): scores [B, M, N] -> log transport scores [B, M+1, N+1] after `iters` iterations of
//     u = log_mu - logsumexp_j(S + v),   v = log_nu - logsumexp_i(S + u)          (S = scores padded with alpha,
// masked rows / columns = -1e12), out = S + u + v - norm.
//
// One CTA per problem.  The iterations are the reference's, evaluated in one of two equivalent forms:
//   LOG  the update exactly as written above (max-subtracted logsumexp): always safe, one exp per matrix entry;
//   LIN  with the potentials absorbed into the plan K = exp(S + u + v) the same update reads
//            a_i = mu_i / sum_j K_ij b_j ,   b_j = nu_j / sum_i K_ij a_i      (u_i += log a_i, v_j += log b_j)
//        one FMA per entry, K constant between absorptions.
// Round 1 switched to LIN after four LOG iterations unconditionally; on real node-level problems (scores up to
// 450, round-2 full-size parity test) entries of K underflowed before their column scaling had grown and the
// result was off by O(100).  Now the form is chosen per iteration:
//   * a LOG iteration whose largest potential step is < 20 is followed by an absorption and LIN iterations;
//   * every LIN iteration checks its scalings: all a_i, b_j of unmasked rows / columns must stay inside
//     [1e-13, 1e13] (this also catches empty sums: inf / NaN).  Inside that band an entry of K lost to fp32
//     underflow (< 1e-38) can contribute at most 1e-25 to a sum that is at least mu_i * 1e-13 ~ 1e-16, i.e. 1e-9
//     relative.  If the check fails the iteration is discarded, the scalings of the PREVIOUS iteration (a
//     consistent pair, kept in a second buffer) are absorbed into (u, v) and the iteration is redone in LOG form.
// Measured against an fp64 evaluation of the reference (tests/test_gpu_pair.py): 7e-6 .. 2e-5 absolute on the log
// scores for score ranges up to +-100; the torch fp32 reference itself is 4e-5 .. 3e-4 away from fp64 there.
//
// Kernels:
//   sinkhorn_patch_kernel    the point-level problems (128 x 128 (+1), thousands per batch): K lives in REGISTERS
//                            (every thread a quarter row and a quarter column); a LIN half-iteration is 33 FFMA per
//                            thread against a vector broadcast from shared memory with 16-byte loads -- the round-1
//                            kernel read K from shared memory: one LDS per FMA, bound by the 128 B/clk port.
//   sinkhorn_general_kernel  any size (node level, ~370 x 360): K in the output buffer (L2), warp per row, row
//                            slabs per warp for the column sums.
#include "common.cuh"

namespace {

struct SinkhornArgs {
  const float* scores;       // [B, M, N]
  const uint8_t* row_mask;   // [B, M] (1 = valid) or NULL
  const uint8_t* col_mask;   // [B, N] or NULL
  const float* alpha;        // device scalar
  float* out;                // [B, M+1, N+1]
  int M, N, iters;
};

// iteration statistics (debug / tuning): [0] LOG iterations, [1] LIN iterations, [2] discarded LIN iterations,
// [3] absorptions; summed over all problems since the last lcr_sinkhorn_stats(reset = 1)
__device__ unsigned long long g_sk_stats[4];

constexpr float kInf = 1e12f;
constexpr float kBigStep = 20.f;             // LOG step above which the next iteration stays in LOG form
constexpr float kHi = 1e13f, kLo = 1e-13f;   // admissible band of the LIN scalings

// exp(x) on the SFU with a compensated argument: ex2.approx(x * log2e) loses |x| * 2^-24 in the
// product; the FMA residual restores it, leaving the ~2 ulp of ex2.approx itself (the same order
// as expf) at a quarter of the instructions.
__device__ __forceinline__ float sk_exp(float x) {
  const float kL2E = 1.4426950408889634f, kL2E_lo = 1.925963033500011e-8f;
  const float y = x * kL2E;
  const float e = fmaf(x, kL2E_lo, fmaf(x, kL2E, -y));
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return fmaf(r, e * 0.6931471805599453f, r);
}

// entry (i, j) of the padded, masked score matrix
__device__ __forceinline__ float sk_score(const float* __restrict__ src, const uint8_t* __restrict__ rm,
                                          const uint8_t* __restrict__ cm, int M, int N, float alpha, int i, int j) {
  if ((i < M && rm && !rm[i]) || (j < N && cm && !cm[j])) return -kInf;
  return (i < M && j < N) ? src[(size_t)i * N + j] : alpha;
}

// ================================================================== point level: 128 x 128 (+ dustbins)
// Thread t = 4 q + h (q < 129, h < 4) holds a QUARTER of row q of the plan (columns 32 h .. 32 h + 31, plus the
// dustbin column for h = 3) and a quarter of column q (rows 32 h .. 32 h + 31, plus the dustbin row): 66 registers.
// Every thread works in both half-iterations -- 33 FFMA against 8 + 1 broadcast loads of the other side's scalings,
// a 4-lane butterfly, one division -- so the 17 warps (4-5 per scheduler) cover the shared-memory latency that a
// one-warp-per-scheduler layout (thread = whole row, first version of this kernel: ptxas kept only two 16-byte
// loads in flight) left exposed: 160 -> ~30 us per problem.
constexpr int PN = 128, PR = 129, PLD = 129, PV = 136, PQ = 33;
constexpr int kPatchThreads = 544;                           // 17 warps: 516 threads carry data
constexpr int kPatchSFloats = 16644;                         // 129 * 129 rounded up to a multiple of 4
constexpr size_t kPatchSmem = sizeof(float) * (kPatchSFloats + 10 * PV);

// Shared-memory layout of a scaling vector x[0..128]: the quarter h of the vector (x[32 h .. 32 h + 31]) is cut into
// eight float4 groups m and stored group-interleaved, group (m, h) at float4 slot 4 m + h; x[128] stays at float
// index 128.  The four threads of a row then read four CONSECUTIVE float4 (64 bytes, the same for all eight rows of
// the warp): one conflict-free wavefront, where the plain layout put the four quarters 128 bytes apart on the same
// banks (4-way conflict on every load).
__device__ __forceinline__ int vec_slot(int x) {          // float index of element x (x < 128)
  return ((((x & 31) >> 2) * 4 + (x >> 5)) << 2) | (x & 3);
}

// sum over this thread's quarter h: k[4 m + r] against x[32 h + 4 m + r], k[32] against x[128].  All nine loads are
// issued before the first FMA (see the __syncwarp below).
__device__ __forceinline__ float dot_quarter(const float (&k)[PQ], const float* __restrict__ vec, int h) {
  const unsigned base = (unsigned)__cvta_generic_to_shared(vec) + 16u * (unsigned)h;
  float4 t[8];
#pragma unroll
  for (int m = 0; m < 8; m++)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(t[m].x), "=f"(t[m].y), "=f"(t[m].z), "=f"(t[m].w)
                 : "r"(base + 64u * (unsigned)m));
  float last;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(last) : "r"((unsigned)__cvta_generic_to_shared(vec) + 4u * PN));
  // ptxas sinks each load to just before its use to save registers (two loads in flight, one ~30-cycle stall per
  // float4); a warp-level barrier is a memory fence it cannot move shared loads across, so all nine loads are
  // issued back to back here and their latencies overlap
  __syncwarp();
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int m = 0; m < 8; m++) {
    s0 = fmaf(k[4 * m], t[m].x, s0);
    s1 = fmaf(k[4 * m + 1], t[m].y, s1);
    s2 = fmaf(k[4 * m + 2], t[m].z, s2);
    s3 = fmaf(k[4 * m + 3], t[m].w, s3);
  }
  s0 = fmaf(k[32], last, s0);
  float s = (s0 + s1) + (s2 + s3);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  return s;
}

__global__ void __launch_bounds__(kPatchThreads, 1) sinkhorn_patch_kernel(SinkhornArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* S = sm;                       // padded log scores, row stride 129
  float* u = sm + kPatchSFloats;       // absorbed log potentials
  float* v = u + PV;
  float* la = v + PV;                  // LIN scalings, two buffers each (current / previous iteration)
  float* lb = la + 2 * PV;
  float* mu = lb + 2 * PV;             // linear marginals (0 for masked rows / columns and the pads)
  float* nu = mu + PV;
  float* log_mu = nu + PV;
  float* log_nu = log_mu + PV;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int q = tid >> 2, h = tid & 3, c0 = 32 * h;
  const bool active = q < PR;
  const int slot = q < PN ? vec_slot(q) : q;     // where element q of a scaling vector lives
  const float alpha = *a.alpha;
  const uint8_t* rm = a.row_mask ? a.row_mask + (size_t)b * PN : nullptr;
  const uint8_t* cm = a.col_mask ? a.col_mask + (size_t)b * PN : nullptr;
  const float* src = a.scores + (size_t)b * PN * PN;

  const float nvr = (float)__syncthreads_count(tid < PN && (!rm || rm[tid]));
  const float nvc = (float)__syncthreads_count(tid < PN && (!cm || cm[tid]));
  const float norm = -logf(nvr + nvc);
  if (tid < PV) {
    const int i = tid;
    const bool masked_r = i >= PR || (i < PN && rm && !rm[i]);
    const bool masked_c = i >= PR || (i < PN && cm && !cm[i]);
    const float lm = masked_r ? -kInf : (i < PN ? norm : logf(nvc) + norm);
    const float ln = masked_c ? -kInf : (i < PN ? norm : logf(nvr) + norm);
    log_mu[i] = lm;
    log_nu[i] = ln;
    mu[i] = lm > -1e11f ? expf(lm) : 0.f;
    nu[i] = ln > -1e11f ? expf(ln) : 0.f;
    u[i] = 0.f;
    v[i] = 0.f;
  }
  for (int e = tid; e < PR * PR; e += kPatchThreads) {
    const int i = e / PR, j = e - i * PR;
    S[i * PLD + j] = sk_score(src, rm, cm, PN, PN, alpha, i, j);
  }
  __syncthreads();

  float kr[PQ], kc[PQ];   // quarter row / quarter column of the absorbed plan
#pragma unroll
  for (int w = 0; w < PQ; w++) kr[w] = kc[w] = 0.f;
  int p = 0;              // current scaling buffer
  bool lin = false, need_absorb = false;   // block-uniform
  int it = 0, streak = 0;                  // streak: LIN iterations since the last absorption
  unsigned n_log = 0, n_lin = 0, n_disc = 0, n_abs = 0;
  while (it < a.iters) {
    if (need_absorb) {
      // K = exp(S + u + v) into registers, scalings = 1
      if (active) {
        const float uq = u[q], vq = v[q];
#pragma unroll
        for (int w = 0; w < 32; w++) {
          kr[w] = sk_exp(S[q * PLD + c0 + w] + uq + v[c0 + w]);
          kc[w] = sk_exp(S[(c0 + w) * PLD + q] + u[c0 + w] + vq);
        }
        kr[32] = h == 3 ? sk_exp(S[q * PLD + PN] + uq + v[PN]) : 0.f;
        kc[32] = h == 3 ? sk_exp(S[PN * PLD + q] + u[PN] + vq) : 0.f;
      }
      if (tid < PV) {
        la[p * PV + tid] = tid < PR ? 1.f : 0.f;
        lb[p * PV + tid] = tid < PR ? 1.f : 0.f;
        la[(p ^ 1) * PV + tid] = 0.f;
        lb[(p ^ 1) * PV + tid] = 0.f;
      }
      __syncthreads();
      need_absorb = false;
      streak = 0;
      n_abs++;
    }
    if (lin) {
      float* a_new = la + (p ^ 1) * PV;
      float* b_new = lb + (p ^ 1) * PV;
      bool bad = false;
      {
        const float s = dot_quarter(kr, lb + p * PV, h);
        const float m_ = mu[q];
        float an = 0.f;
        if (m_ > 0.f) {
          an = m_ / s;
          bad = !(an < kHi && an > kLo);
        }
        if (active && h == 0) a_new[slot] = an;
      }
      bool fail = __syncthreads_or(bad) != 0;
      if (!fail) {
        const float s = dot_quarter(kc, a_new, h);
        const float n_ = nu[q];
        float bn = 0.f;
        bad = false;
        if (n_ > 0.f) {
          bn = n_ / s;
          bad = !(bn < kHi && bn > kLo);
        }
        if (active && h == 0) b_new[slot] = bn;
        fail = __syncthreads_or(bad) != 0;
      }
      if (!fail) {
        p ^= 1;
        it++;
        streak++;
        n_lin++;
        continue;
      }
      // discard this iteration and absorb the previous (consistent) scalings into the potentials
      n_disc++;
      if (tid < PR) {
        const int sl = tid < PN ? vec_slot(tid) : tid;
        const float av = la[p * PV + sl], bv = lb[p * PV + sl];
        if (av > 0.f) u[tid] += logf(av);
        if (bv > 0.f) v[tid] += logf(bv);
      }
      __syncthreads();
      if (streak > 0) {      // the drift since the last absorption left the band: re-absorb, retry in LIN form
        need_absorb = true;
        continue;
      }
      lin = false;           // a single step left the band (> e^30): redo the iteration in LOG form
    }
    // ---- LOG iteration on the shared-memory scores: one thread per row, then one thread per column
    float step = 0.f;
    if (tid < PR) {
      const float* row = S + tid * PLD;
      float mx = -INFINITY;
#pragma unroll 4
      for (int j = 0; j < PR; j++) mx = fmaxf(mx, row[j] + v[j]);
      float sum = 0.f;
#pragma unroll 4
      for (int j = 0; j < PR; j++) sum += sk_exp(row[j] + v[j] - mx);
      const float un = log_mu[tid] - (mx + logf(sum));
      if (mu[tid] > 0.f) step = fabsf(un - u[tid]);
      u[tid] = un;
    }
    __syncthreads();
    if (tid < PR) {
      float mx = -INFINITY;
#pragma unroll 4
      for (int i = 0; i < PR; i++) mx = fmaxf(mx, S[i * PLD + tid] + u[i]);
      float sum = 0.f;
#pragma unroll 4
      for (int i = 0; i < PR; i++) sum += sk_exp(S[i * PLD + tid] + u[i] - mx);
      const float vn = log_nu[tid] - (mx + logf(sum));
      if (nu[tid] > 0.f) step = fmaxf(step, fabsf(vn - v[tid]));
      v[tid] = vn;
    }
    const bool big = __syncthreads_or(!(step < kBigStep)) != 0;
    it++;
    n_log++;
    if (!big && it < a.iters) {
      need_absorb = true;
      lin = true;
    }
  }
  if (tid == 0) {
    atomicAdd(&g_sk_stats[0], (unsigned long long)n_log);
    atomicAdd(&g_sk_stats[1], (unsigned long long)n_lin);
    atomicAdd(&g_sk_stats[2], (unsigned long long)n_disc);
    atomicAdd(&g_sk_stats[3], (unsigned long long)n_abs);
  }
  if (lin && !need_absorb && tid < PR) {
    const int sl = tid < PN ? vec_slot(tid) : tid;
    const float av = la[p * PV + sl], bv = lb[p * PV + sl];
    if (av > 0.f) u[tid] += logf(av);
    if (bv > 0.f) v[tid] += logf(bv);
  }
  __syncthreads();
  float* dst = a.out + (size_t)b * PR * PR;
  for (int e = tid; e < PR * PR; e += kPatchThreads) {
    const int i = e / PR, j = e - i * PR;
    dst[e] = S[i * PLD + j] + u[i] + v[j] - norm;
  }
}

// ================================================================== general size: plan in global memory (L2)
constexpr int kGenThreads = 1024;

__global__ void __launch_bounds__(kGenThreads) sinkhorn_general_kernel(SinkhornArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int M = a.M, N = a.N, R = M + 1, C = N + 1;
  float* u = sm;
  float* v = u + R;
  float* la = v + C;             // two buffers of R
  float* lb = la + 2 * R;        // two buffers of C
  float* mu = lb + 2 * C;
  float* nu = mu + R;
  float* log_mu = nu + C;
  float* log_nu = log_mu + R;
  float* part_m = log_nu + C;    // per-warp column partials [nw x C]
  float* part_s = part_m + (size_t)(kGenThreads / 32) * C;
  const int b = blockIdx.x, tid = threadIdx.x, nt = kGenThreads;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const float alpha = *a.alpha;
  const uint8_t* rm = a.row_mask ? a.row_mask + (size_t)b * M : nullptr;
  const uint8_t* cm = a.col_mask ? a.col_mask + (size_t)b * N : nullptr;
  const float* src = a.scores + (size_t)b * M * N;
  float* K = a.out + (size_t)b * R * C;

  __shared__ int s_cnt[2];
  if (tid < 2) s_cnt[tid] = 0;
  __syncthreads();
  int c0 = 0, c1 = 0;
  for (int i = tid; i < M; i += nt) c0 += rm ? rm[i] : 1;
  for (int j = tid; j < N; j += nt) c1 += cm ? cm[j] : 1;
  c0 = lcr_warp_sum(c0);
  c1 = lcr_warp_sum(c1);
  if (lane == 0) {
    atomicAdd(&s_cnt[0], c0);
    atomicAdd(&s_cnt[1], c1);
  }
  __syncthreads();
  const float nvr = (float)s_cnt[0], nvc = (float)s_cnt[1];
  const float norm = -logf(nvr + nvc);
  for (int i = tid; i < R; i += nt) {
    const bool masked = i < M && rm && !rm[i];
    const float lm = masked ? -kInf : (i < M ? norm : logf(nvc) + norm);
    log_mu[i] = lm;
    mu[i] = lm > -1e11f ? expf(lm) : 0.f;
    u[i] = 0.f;
  }
  for (int j = tid; j < C; j += nt) {
    const bool masked = j < N && cm && !cm[j];
    const float ln = masked ? -kInf : (j < N ? norm : logf(nvr) + norm);
    log_nu[j] = ln;
    nu[j] = ln > -1e11f ? expf(ln) : 0.f;
    v[j] = 0.f;
  }
  __syncthreads();
  const int rs = (R + nw - 1) / nw, r0 = warp * rs, r1 = min(r0 + rs, R);   // this warp's row slab (column sums)

  int p = 0;
  bool lin = false, need_absorb = false;
  int it = 0, streak = 0;
  unsigned n_log = 0, n_lin = 0, n_disc = 0, n_abs = 0;
  while (it < a.iters) {
    if (need_absorb) {
      for (int e = tid; e < R * C; e += nt) {
        const int i = e / C, j = e - i * C;
        K[e] = sk_exp(sk_score(src, rm, cm, M, N, alpha, i, j) + u[i] + v[j]);
      }
      for (int i = tid; i < R; i += nt) {
        la[p * R + i] = 1.f;
        la[(p ^ 1) * R + i] = 0.f;
      }
      for (int j = tid; j < C; j += nt) {
        lb[p * C + j] = 1.f;
        lb[(p ^ 1) * C + j] = 0.f;
      }
      __syncthreads();
      need_absorb = false;
      streak = 0;
      n_abs++;
    }
    if (lin) {
      float* a_new = la + (p ^ 1) * R;
      float* b_new = lb + (p ^ 1) * C;
      const float* b_cur = lb + p * C;
      bool bad = false;
      for (int i = warp; i < R; i += nw) {       // warp per row, coalesced
        const float* row = K + (size_t)i * C;
        float sum = 0.f;
        for (int j = lane; j < C; j += 32) sum = fmaf(row[j], b_cur[j], sum);
        sum = lcr_warp_sum(sum);
        if (lane == 0) {
          const float m_ = mu[i];
          float an = 0.f;
          if (m_ > 0.f) {
            an = m_ / sum;
            bad |= !(an < kHi && an > kLo);
          }
          a_new[i] = an;
        }
      }
      bool fail = __syncthreads_or(bad) != 0;
      if (!fail) {
        for (int j = lane; j < C; j += 32) {      // row slab per warp, lanes over columns
          float sum = 0.f;
          for (int i = r0; i < r1; i++) sum = fmaf(K[(size_t)i * C + j], a_new[i], sum);
          part_s[(size_t)warp * C + j] = sum;
        }
        __syncthreads();
        bad = false;
        for (int j = tid; j < C; j += nt) {
          float sum = 0.f;
          for (int w = 0; w < nw; w++) sum += part_s[(size_t)w * C + j];
          const float n_ = nu[j];
          float bn = 0.f;
          if (n_ > 0.f) {
            bn = n_ / sum;
            bad |= !(bn < kHi && bn > kLo);
          }
          b_new[j] = bn;
        }
        fail = __syncthreads_or(bad) != 0;
      }
      if (!fail) {
        p ^= 1;
        it++;
        streak++;
        n_lin++;
        continue;
      }
      n_disc++;
      for (int i = tid; i < R; i += nt) {
        const float av = la[p * R + i];
        if (av > 0.f) u[i] += logf(av);
      }
      for (int j = tid; j < C; j += nt) {
        const float bv = lb[p * C + j];
        if (bv > 0.f) v[j] += logf(bv);
      }
      __syncthreads();
      if (streak > 0) {
        need_absorb = true;
        continue;
      }
      lin = false;
    }
    // ---- LOG iteration straight from the input scores
    bool big = false;
    for (int i = warp; i < R; i += nw) {
      float mx = -INFINITY;
      for (int j = lane; j < C; j += 32) mx = fmaxf(mx, sk_score(src, rm, cm, M, N, alpha, i, j) + v[j]);
      mx = lcr_warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < C; j += 32) sum += sk_exp(sk_score(src, rm, cm, M, N, alpha, i, j) + v[j] - mx);
      sum = lcr_warp_sum(sum);
      if (lane == 0) {
        const float un = log_mu[i] - (mx + logf(sum));
        if (mu[i] > 0.f) big |= !(fabsf(un - u[i]) < kBigStep);
        u[i] = un;
      }
    }
    __syncthreads();
    for (int j = lane; j < C; j += 32) {
      float mx = -INFINITY;
      for (int i = r0; i < r1; i++) mx = fmaxf(mx, sk_score(src, rm, cm, M, N, alpha, i, j) + u[i]);
      float sum = 0.f;
      for (int i = r0; i < r1; i++) sum += sk_exp(sk_score(src, rm, cm, M, N, alpha, i, j) + u[i] - mx);
      part_m[(size_t)warp * C + j] = mx;
      part_s[(size_t)warp * C + j] = sum;
    }
    __syncthreads();
    for (int j = tid; j < C; j += nt) {
      float mx = -INFINITY;
      for (int w = 0; w < nw; w++) mx = fmaxf(mx, part_m[(size_t)w * C + j]);
      float sum = 0.f;
      for (int w = 0; w < nw; w++) {
        const float pm = part_m[(size_t)w * C + j];
        if (pm > -INFINITY) sum += part_s[(size_t)w * C + j] * sk_exp(pm - mx);
      }
      const float vn = log_nu[j] - (mx + logf(sum));
      if (nu[j] > 0.f) big |= !(fabsf(vn - v[j]) < kBigStep);
      v[j] = vn;
    }
    big = __syncthreads_or(big) != 0;
    it++;
    n_log++;
    if (!big && it < a.iters) {
      need_absorb = true;
      lin = true;
    }
  }
  if (tid == 0) {
    atomicAdd(&g_sk_stats[0], (unsigned long long)n_log);
    atomicAdd(&g_sk_stats[1], (unsigned long long)n_lin);
    atomicAdd(&g_sk_stats[2], (unsigned long long)n_disc);
    atomicAdd(&g_sk_stats[3], (unsigned long long)n_abs);
  }
  if (lin && !need_absorb) {
    for (int i = tid; i < R; i += nt) {
      const float av = la[p * R + i];
      if (av > 0.f) u[i] += logf(av);
    }
    for (int j = tid; j < C; j += nt) {
      const float bv = lb[p * C + j];
      if (bv > 0.f) v[j] += logf(bv);
    }
  }
  __syncthreads();
  for (int e = tid; e < R * C; e += nt) {
    const int i = e / C, j = e - i * C;
    K[e] = sk_score(src, rm, cm, M, N, alpha, i, j) + u[i] + v[j] - norm;
  }
}

}  // namespace

// Debug / tuning: iteration statistics of all Sinkhorn problems since the last reset (see g_sk_stats); synchronises.
extern "C" int lcr_sinkhorn_stats(int64_t* out4, int reset) {
  unsigned long long h[4] = {0, 0, 0, 0};
  LCR_CUDA_TRY(cudaMemcpyFromSymbol(h, g_sk_stats, sizeof(h)));
  if (out4)
    for (int i = 0; i < 4; i++) out4[i] = (int64_t)h[i];
  if (reset) {
    const unsigned long long z[4] = {0, 0, 0, 0};
    LCR_CUDA_TRY(cudaMemcpyToSymbol(g_sk_stats, z, sizeof(z)));
  }
  return LCR_OK;
}

extern "C" int lcr_sinkhorn(const float* scores, int batch, int rows, int cols, const uint8_t* row_mask,
                            const uint8_t* col_mask, const float* alpha, int iters, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(batch >= 0 && rows >= 1 && cols >= 1 && iters >= 0, "sinkhorn: sizes");
  LCR_REQUIRE(scores && alpha && out, "sinkhorn: null pointer");
  if (batch == 0) return LCR_OK;
  SinkhornArgs a{scores, row_mask, col_mask, alpha, out, rows, cols, iters};
  const bool patch = rows == PN && cols == PN;
  LcrProfScope prof(patch ? "sinkhorn_point" : "sinkhorn_node", 4.0 * batch * (double)(rows + 1) * (cols + 1) * iters,
                    4.0 * batch * ((double)rows * cols + (double)(rows + 1) * (cols + 1)), stream);
  if (patch) {
    static LcrOncePerDevice once;
    const int dev = once.need();
    if (dev != -1) {
      LCR_CUDA_TRY(cudaFuncSetAttribute(sinkhorn_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kPatchSmem));
      once.done(dev);
    }
    sinkhorn_patch_kernel<<<batch, kPatchThreads, kPatchSmem, stream>>>(a);
  } else {
    const size_t R = (size_t)rows + 1, C = (size_t)cols + 1;
    const size_t smem = sizeof(float) * (5 * R + 5 * C + 2 * (kGenThreads / 32) * C) + 64;
    LCR_REQUIRE(smem <= 200 * 1024, "sinkhorn: problem too large (5 (rows + cols) + 64 cols floats of shared memory)");
    LCR_CUDA_TRY(cudaFuncSetAttribute(sinkhorn_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sinkhorn_general_kernel<<<batch, kGenThreads, smem, stream>>>(a);
  }
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
