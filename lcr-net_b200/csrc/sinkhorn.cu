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
// Exact early exit (point level): a LIN iteration whose scalings all equal, bit for bit, those of two iterations before
// has closed a cycle of period 1 or 2; the remaining iterations only replay it, so the loop stops and the final state is
// picked by the parity of the iterations left (lcr_set_sinkhorn_early_exit(0) runs the full count: same bits).
//
// Kernels:
//   sinkhorn_patch_kernel    the point-level problems (128 x 128 (+1), thousands per batch): K lives in REGISTERS in
//                            row form and in column form (3 x 17 blocks per thread each); a LIN half-iteration is
//                            51 FFMA per thread against 17 words of shared memory -- the round-1 kernel read K from
//                            shared memory: one LDS per FMA, bound by the 128 B/clk port.
//   sinkhorn_cluster_kernel  node level (~390 x 385, one problem per pair): a cluster of 4 / 8 CTAs per problem, row
//                            slabs of K in shared memory, column partials exchanged through DSMEM.
//   sinkhorn_general_kernel  any size: K in the output buffer (L2), warp per row, row slabs per warp for the column
//                            sums (fallback when the slabs do not fit, LCR_SINKHORN_CLUSTER=0).
#include <cooperative_groups.h>

#include "common.cuh"

namespace {

struct SinkhornArgs {
  const float* scores;       // [B, M, N]
  const uint8_t* row_mask;   // [B, M] (1 = valid) or NULL
  const uint8_t* col_mask;   // [B, N] or NULL
  const float* alpha;        // device scalar
  float* out;                // [B, M+1, N+1]
  int M, N, iters;
  int early_exit;            // 1: stop at a bitwise cycle of the scalings (exact), 0: always run `iters` iterations
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
// The plan (129 x 129, zero-padded to 132 x 136) lives in REGISTERS, twice: thread (g, h) -- 44 groups x 8 chunks,
// 352 of the 384 threads -- holds the ROW-form block  K[3 g .. 3 g + 2][17 h .. 17 h + 16]  and the COLUMN-form block
// K[17 h .. 17 h + 16][3 g .. 3 g + 2]  (2 x 51 registers).  A half-iteration is 51 FFMA per thread against 17 words
// of the other side's scalings: every word loaded from shared memory feeds THREE multiply-adds.  That ratio is the
// point: shared memory delivers one 32-bit word per lane and wavefront whatever the load width or broadcast, so a
// thread-per-row (1 x 129) or quarter-row (1 x 33) layout needs 129 x 129 words per half-iteration -- 520 cycles of
// the 128 B/clk port, measured 50 % busy and short-scoreboard bound (ncu, profiles/r2_ncu_sinkhorn_quarter_rows.txt).
// Both forms in EVERY thread (rather than 3 x 33 row blocks in one half of the CTA and column blocks in the other,
// the previous version of this kernel) keeps all 12 warps busy in both half-iterations: with role-split warps only 6
// warps -- 1.5 per scheduler -- were runnable at any time and the half-iteration sat on FMA / shuffle / barrier
// latency (ncu: issue 33 %, FMA pipe 20 %, profiles/r2_ncu_sinkhorn_patch.txt).  The partial sums of a row
// (column) group are combined by an 8-lane butterfly; lane h < 3 finishes row (column) 3 g + h.  Scaling vectors sit
// in shared memory as eight 20-float segments (element x at 20 (x / 17) + x % 17): the 17 words of a thread are four
// aligned 16-byte loads + one, and the eight segments of a quarter-warp fall on distinct banks (20 h mod 32).
constexpr int PN = 128, PR = 129, PLD = 129, PP = 132, PB = 17, PH = 8, PSEG = 20, PV = PH * PSEG;
constexpr int RB = 3;                                        // rows (columns) per thread block: 3 x 17 = 51 registers
constexpr int NG = PP / RB;                                  // 44 row (column) groups
constexpr int kPatchThreads = 384;                           // 12 warps, 352 threads carry data
constexpr int kPatchSFloats = 16644;                         // 129 * 129 rounded up to a multiple of 4
constexpr size_t kPatchSmem = sizeof(float) * (kPatchSFloats + 4 * PV + 6 * PP + 4);

__device__ __forceinline__ int vec_slot(int x) { return PSEG * (x / PB) + x % PB; }   // x < 136

// s[r] = sum_c k[r][c] * x[17 h + c], then the 8-lane butterfly over h.  All five loads are issued before the
// first FMA: left alone, ptxas sinks every load to just before its use and recycles ONE register quad (ncu source
// view: each LDS.128 exposed its full latency, ~30 % of the half-iteration), and neither volatile asm
// loads nor a warp barrier keep it from doing so.  The accumulators therefore start from a zero that is
// data-dependent on all loads, which forces them to be in flight together.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// `seg` = 32-bit shared-window address of this thread's 20-float segment of the scaling vector (generic pointers cost
// a window conversion per access: S2R + LEA showed up in every half-iteration)
__device__ __forceinline__ void dot_block(const float (&k)[RB][PB], uint32_t seg, float (&s)[RB]) {
  float4 t[4];
#pragma unroll
  for (int m = 0; m < 4; m++) t[m] = lds128(seg + 16u * m);
  const float last = lds32(seg + 64u);
  // zero, but data-dependent on every load (x * 0 is not foldable in IEEE arithmetic; the scalings are finite)
  const float z = (((t[0].x + t[1].x) + (t[2].x + t[3].x)) + last) * 0.f;
#pragma unroll
  for (int r = 0; r < RB; r++) s[r] = z;
#pragma unroll
  for (int m = 0; m < 4; m++) {
#pragma unroll
    for (int r = 0; r < RB; r++) {
      s[r] = fmaf(k[r][4 * m], t[m].x, s[r]);
      s[r] = fmaf(k[r][4 * m + 1], t[m].y, s[r]);
      s[r] = fmaf(k[r][4 * m + 2], t[m].z, s[r]);
      s[r] = fmaf(k[r][4 * m + 3], t[m].w, s[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < RB; r++) s[r] = fmaf(k[r][16], last, s[r]);
#pragma unroll
  for (int d = 1; d < PH; d <<= 1) {
#pragma unroll
    for (int r = 0; r < RB; r++) s[r] += __shfl_xor_sync(0xffffffffu, s[r], d);
  }
}

__global__ void __launch_bounds__(kPatchThreads, 1) sinkhorn_patch_kernel(SinkhornArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* S = sm;                       // padded log scores, row stride 129
  float* la = sm + kPatchSFloats;      // LIN scalings in slot layout, two buffers each (current / previous iteration)
  float* lb = la + 2 * PV;
  float* u = lb + 2 * PV;              // absorbed log potentials (plain layout, padded to 132)
  float* v = u + PP;
  float* mu = v + PP;                  // linear marginals (0 for masked rows / columns and the pads)
  float* nu = mu + PP;
  float* log_mu = nu + PP;
  float* log_nu = log_mu + PP;
  float* chg = log_nu + PP;            // [2] "a scaling changed in iteration it" flags (iteration numbers as bits)
  const int b = blockIdx.x, tid = threadIdx.x;
  const int g = tid >> 3, h = tid & (PH - 1);
  const bool active = g < NG;                    // 44 groups of 3 rows (columns)
  const bool finisher = active && h < RB;        // lane h < 3 finishes row AND column 3 g + h
  const int mine = finisher ? RB * g + h : PP - 1;   // pad row otherwise (marginal 0)
  const float alpha = *a.alpha;
  const uint8_t* rm = a.row_mask ? a.row_mask + (size_t)b * PN : nullptr;
  const uint8_t* cm = a.col_mask ? a.col_mask + (size_t)b * PN : nullptr;
  const float* src = a.scores + (size_t)b * PN * PN;

  const float nvr = (float)__syncthreads_count(tid < PN && (!rm || rm[tid]));
  const float nvc = (float)__syncthreads_count(tid < PN && (!cm || cm[tid]));
  const float norm = -logf(nvr + nvc);
  if (tid < PP) {
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
  if (tid < 2) chg[tid] = __int_as_float(-1);
  for (int e = tid; e < 2 * PV; e += kPatchThreads) {
    la[e] = 0.f;
    lb[e] = 0.f;
  }
  // padded, masked scores -> S.  The body streams in as 16-byte loads (11 per thread, all in flight together: the
  // element-wise form with its two mask bytes per element cost 26 us per problem, a fifth of the kernel)
  {
    float* mr = la;            // masks as floats, staged in the (not yet used) scaling buffers: 1 = valid
    float* mc = la + PV;
    __syncthreads();           // the zero-fill of la / lb above is complete
    if (tid < PN) {
      mr[tid] = (!rm || rm[tid]) ? 1.f : 0.f;
      mc[tid] = (!cm || cm[tid]) ? 1.f : 0.f;
    }
    __syncthreads();
    constexpr int kVec = PN * PN / 4;                              // 4096 float4
    constexpr int kPer = (kVec + kPatchThreads - 1) / kPatchThreads;
    float4 val[kPer];
#pragma unroll
    for (int q = 0; q < kPer; q++) {
      const int e = tid + q * kPatchThreads;
      val[q] = e < kVec ? *reinterpret_cast<const float4*>(src + 4 * (size_t)e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < kPer; q++) {
      const int e = tid + q * kPatchThreads;
      if (e < kVec) {
        const int i = e >> 5, j = (e & 31) << 2;
        const bool ri = mr[i] != 0.f;
        float* d = S + i * PLD + j;
        d[0] = (ri && mc[j] != 0.f) ? val[q].x : -kInf;
        d[1] = (ri && mc[j + 1] != 0.f) ? val[q].y : -kInf;
        d[2] = (ri && mc[j + 2] != 0.f) ? val[q].z : -kInf;
        d[3] = (ri && mc[j + 3] != 0.f) ? val[q].w : -kInf;
      }
    }
    if (tid < PN) {                                                // dustbin column and row
      S[tid * PLD + PN] = mr[tid] != 0.f ? alpha : -kInf;
      S[PN * PLD + tid] = mc[tid] != 0.f ? alpha : -kInf;
    }
    if (tid == 0) S[PN * PLD + PN] = alpha;
    __syncthreads();
    for (int e = tid; e < 2 * PV; e += kPatchThreads) la[e] = 0.f;
  }
  __syncthreads();

  float kr[RB][PB], kc[RB][PB];        // this thread's row-form and column-form blocks of the absorbed plan
#pragma unroll
  for (int r = 0; r < RB; r++)
#pragma unroll
    for (int c = 0; c < PB; c++) kr[r][c] = kc[r][c] = 0.f;
  int p = 0;              // current scaling buffer
  bool lin = false, need_absorb = false;   // block-uniform
  int it = 0, streak = 0;                  // streak: LIN iterations since the last absorption
  unsigned n_log = 0, n_lin = 0, n_disc = 0, n_abs = 0;
  const float my_mu = mu[mine], my_nu = nu[mine];
  const int my_slot = vec_slot(mine);
  while (it < a.iters) {
    if (need_absorb) {
      // K = exp(S + u + v) into registers (zero outside 129 x 129), scalings = 1
      if (active) {
#pragma unroll
        for (int r = 0; r < RB; r++) {
          const int x = RB * g + r;                // row of the row-form block, column of the column-form block
#pragma unroll
          for (int c = 0; c < PB; c++) {
            const int y = PB * h + c;              // column of the row-form block, row of the column-form block
            const bool in = x < PR && y < PR;
            kr[r][c] = in ? sk_exp(S[x * PLD + y] + u[x] + v[y]) : 0.f;
            kc[r][c] = in ? sk_exp(S[y * PLD + x] + u[y] + v[x]) : 0.f;
          }
        }
      }
      if (tid < PP) {
        const int sl = vec_slot(tid);
        la[p * PV + sl] = tid < PR ? 1.f : 0.f;
        lb[p * PV + sl] = tid < PR ? 1.f : 0.f;
        la[(p ^ 1) * PV + sl] = 0.f;
        lb[(p ^ 1) * PV + sl] = 0.f;
      }
      __syncthreads();
      need_absorb = false;
      streak = 0;
      n_abs++;
    }
    if (lin) {
      // tight loop over consecutive LIN iterations; leaves on the first out-of-band scaling (fail) or at the end
      const uint32_t la_s = (uint32_t)__cvta_generic_to_shared(la), lb_s = (uint32_t)__cvta_generic_to_shared(lb);
      const uint32_t seg_off = 4u * PSEG * (uint32_t)h, slot_off = 4u * (uint32_t)my_slot;
      bool fail = false, fixed = false;
      // Exact early exit: when an iteration reproduces BIT FOR BIT every scaling of the iteration two steps back
      // (whose values still sit in the buffer it is about to overwrite), the sequence has entered a cycle of period
      // 1 or 2 -- same inputs, same arithmetic -- and the state after the full iteration count is the current one or
      // the previous one, by the parity of the iterations left: both are in shared memory.  Finishers that see a
      // changed value store the iteration number into chg[it & 1]; the flag of iteration `it` is read after its
      // second barrier and is rewritten two iterations (four barriers) later.
      const uint32_t chg_s = (uint32_t)__cvta_generic_to_shared(chg);
      while (it < a.iters) {
        const uint32_t cur = 4u * PV * (uint32_t)p, nxt = 4u * PV * (uint32_t)(p ^ 1);
        const uint32_t chg_it = chg_s + 4u * (uint32_t)(it & 1);
        bool bad;
        {
          float s[RB];
          const float prev = finisher ? lds32(la_s + nxt + slot_off) : 0.f;      // a of iteration it - 2
          dot_block(kr, lb_s + cur + seg_off, s);
          const float sum = h == 0 ? s[0] : h == 1 ? s[1] : s[2];
          const float an = my_mu > 0.f ? __fdividef(my_mu, sum) : 0.f;
          bad = my_mu > 0.f && !(an < kHi && an > kLo);
          if (finisher) {
            sts32(la_s + nxt + slot_off, an);
            if (an != prev) sts32(chg_it, __int_as_float(it));
          }
        }
        fail = __syncthreads_or(bad) != 0;
        if (fail) break;
        {
          float s[RB];
          const float prev = finisher ? lds32(lb_s + nxt + slot_off) : 0.f;
          dot_block(kc, la_s + nxt + seg_off, s);
          const float sum = h == 0 ? s[0] : h == 1 ? s[1] : s[2];
          const float bn = my_nu > 0.f ? __fdividef(my_nu, sum) : 0.f;
          bad = my_nu > 0.f && !(bn < kHi && bn > kLo);
          if (finisher) {
            sts32(lb_s + nxt + slot_off, bn);
            if (bn != prev) sts32(chg_it, __int_as_float(it));
          }
        }
        fail = __syncthreads_or(bad) != 0;
        if (fail) break;
        fixed = __float_as_int(lds32(chg_it)) != it;
        p ^= 1;
        it++;
        streak++;
        n_lin++;
        if (fixed && a.early_exit) break;
        fixed = false;
      }
      if (fixed) {               // the remaining iterations only alternate between the last two states
        if ((a.iters - it) & 1) p ^= 1;
        it = a.iters;
      }
      if (!fail) break;          // all iterations done
      // discard this iteration and absorb the previous (consistent) scalings into the potentials
      n_disc++;
      if (tid < PR) {
        const int sl = vec_slot(tid);
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
    const int sl = vec_slot(tid);
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


// ================================================================== node level on a thread-block cluster
// One CLUSTER of CL CTAs per problem: rank r keeps rows [r rs, (r + 1) rs) of the plan in SHARED memory (a 98 x 388
// slab = 150 KB for the ~390 x 385 node problems), so a LIN iteration reads the plan from shared memory instead of
// L2 (the single-CTA kernel above is bound by L2 latency: 55 us per iteration, 32 CTAs on a 148-SM GPU).  The row
// update is local to a rank (a_i only feeds the column sums of the same rows); the column update needs the sum over
// all ranks: every rank writes its partial column sums (LOG form: partial max / sum pairs) and its "row check
// failed" flag into EVERY peer's shared memory through DSMEM, one cluster barrier, then all ranks combine the CL
// partials in rank order -- bit-identical b_j, potentials and control decisions on every rank, no second barrier.
// The exchange buffers are double-buffered on the exchange counter: a peer can run at most one exchange ahead.
constexpr int kClThreads = 512;
constexpr int kClChunks = 16;                      // columns <= 32 * kClChunks
constexpr size_t kClSmemMax = 226 * 1024;          // 227 KB opt-in limit minus the static shared memory

__host__ __device__ inline size_t cluster_smem_floats(int CL, int C, int rs) {
  const size_t Cs = (size_t)(C + 3) & ~(size_t)3;
  return (size_t)rs * Cs      // plan slab
         + 6 * (size_t)rs     // u, la x 2, mu, log_mu, row mask of the local rows
         + 6 * Cs             // v, lb x 2, nu, log_nu, column mask (replicated)
         + 4 * CL * Cs        // exchange: [2 buffers][2 planes][CL ranks][Cs]
         + 2 * CL + 8;        // flags [2][CL]
}

template <int CL>
__global__ void __launch_bounds__(kClThreads, 1) sinkhorn_cluster_kernel(SinkhornArgs a, int rs) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  const int M = a.M, N = a.N, R = M + 1, C = N + 1;
  const int Cs = (C + 3) & ~3;
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / CL, tid = threadIdx.x, nt = kClThreads;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int i0 = rank * rs, nloc = max(0, min(rs, R - i0));
  float* Ks = sm;                          // [rs][Cs]
  float* u = Ks + (size_t)rs * Cs;         // local rows
  float* la = u + rs;                      // [2][rs]
  float* mu = la + 2 * rs;
  float* log_mu = mu + rs;
  float* rmask = log_mu + rs;              // 1 = valid row (the dustbin row is valid)
  float* v = rmask + rs;                   // replicated columns
  float* lb = v + Cs;                      // [2][Cs]
  float* nu = lb + 2 * Cs;
  float* log_nu = nu + Cs;
  float* cmask = log_nu + Cs;
  float* exch = cmask + Cs;                // [2][2][CL][Cs]
  int* flags = reinterpret_cast<int*>(exch + 4 * CL * Cs);   // [2][CL]
  const float alpha = *a.alpha;
  const uint8_t* rm = a.row_mask ? a.row_mask + (size_t)b * M : nullptr;
  const uint8_t* cm = a.col_mask ? a.col_mask + (size_t)b * N : nullptr;
  const float* src = a.scores + (size_t)b * M * N;
  float* out = a.out + (size_t)b * R * C;
  float* peer_exch[CL];
  int* peer_flags[CL];
#pragma unroll
  for (int r = 0; r < CL; r++) {
    peer_exch[r] = cluster.map_shared_rank(exch, r);
    peer_flags[r] = cluster.map_shared_rank(flags, r);
  }
  // padded, masked score of local row i (global row i0 + i) and column j, masks from shared memory
  auto score = [&](int i, int j) -> float {
    if (rmask[i] == 0.f || cmask[j] == 0.f) return -kInf;
    return (i0 + i < M && j < N) ? src[(size_t)(i0 + i) * N + j] : alpha;
  };

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
  for (int i = tid; i < nloc; i += nt) {
    const int gi = i0 + i;
    const bool masked = gi < M && rm && !rm[gi];
    const float lm = masked ? -kInf : (gi < M ? norm : logf(nvc) + norm);
    log_mu[i] = lm;
    mu[i] = lm > -1e11f ? expf(lm) : 0.f;
    rmask[i] = masked ? 0.f : 1.f;
    u[i] = 0.f;
  }
  for (int j = tid; j < C; j += nt) {
    const bool masked = j < N && cm && !cm[j];
    const float ln = masked ? -kInf : (j < N ? norm : logf(nvr) + norm);
    log_nu[j] = ln;
    nu[j] = ln > -1e11f ? expf(ln) : 0.f;
    cmask[j] = masked ? 0.f : 1.f;
    v[j] = 0.f;
  }
  cluster.sync();        // every rank's shared memory is live before the first remote store

  int p = 0, x = 0;      // scaling buffer, exchange counter
  bool lin = false, need_absorb = false;
  int it = 0, streak = 0;
  unsigned n_log = 0, n_lin = 0, n_disc = 0, n_abs = 0;
  while (it < a.iters) {
    if (need_absorb) {
      for (int i = warp; i < nloc; i += nw) {      // warp per row, lanes over columns: coalesced score loads
        const float ui = u[i];
        float sc[kClChunks];
#pragma unroll
        for (int k = 0; k < kClChunks; k++) {
          const int j = lane + 32 * k;
          sc[k] = j < C ? score(i, j) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kClChunks; k++) {
          const int j = lane + 32 * k;
          if (j < C) Ks[(size_t)i * Cs + j] = sk_exp(sc[k] + ui + v[j]);
        }
      }
      for (int i = tid; i < nloc; i += nt) {
        la[p * rs + i] = 1.f;
        la[(p ^ 1) * rs + i] = 0.f;
      }
      for (int j = tid; j < C; j += nt) {
        lb[p * Cs + j] = 1.f;
        lb[(p ^ 1) * Cs + j] = 0.f;
      }
      __syncthreads();
      need_absorb = false;
      streak = 0;
      n_abs++;
    }
    float* ex = exch + (size_t)(x & 1) * 2 * CL * Cs;           // this exchange: [2][CL][Cs]
    const size_t ex_off = (size_t)(x & 1) * 2 * CL * Cs + (size_t)rank * Cs;
    int* fl = flags + (x & 1) * CL;
    if (lin) {
      float* a_new = la + (p ^ 1) * rs;
      float* b_new = lb + (p ^ 1) * Cs;
      const float* b_cur = lb + p * Cs;
      bool bad = false, chg = false;                // chg: a scaling differs from its value two iterations back
      {
        float bc[kClChunks];                        // this lane's columns of the current scalings
#pragma unroll
        for (int k = 0; k < kClChunks; k++) {
          const int j = lane + 32 * k;
          bc[k] = j < C ? b_cur[j] : 0.f;
        }
        for (int i = warp; i < nloc; i += nw) {     // warp per row of the slab
          const float* row = Ks + (size_t)i * Cs;
          float kv[kClChunks];
#pragma unroll
          for (int k = 0; k < kClChunks; k++) {
            const int j = lane + 32 * k;
            kv[k] = j < C ? row[j] : 0.f;
          }
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int k = 0; k < kClChunks; k += 4) {
            s0 = fmaf(kv[k], bc[k], s0);
            s1 = fmaf(kv[k + 1], bc[k + 1], s1);
            s2 = fmaf(kv[k + 2], bc[k + 2], s2);
            s3 = fmaf(kv[k + 3], bc[k + 3], s3);
          }
          const float sum = lcr_warp_sum((s0 + s1) + (s2 + s3));
          if (lane == 0) {
            const float m_ = mu[i];
            float an = 0.f;
            if (m_ > 0.f) {
              an = __fdividef(m_, sum);
              bad |= !(an < kHi && an > kLo);
            }
            chg |= an != a_new[i];                  // (a_new still holds the scaling of iteration it - 2)
            a_new[i] = an;
          }
        }
      }
      const int row_bad = (__syncthreads_or(bad) ? 1 : 0) | (__syncthreads_or(chg) ? 2 : 0);
      for (int j = tid; j < C; j += nt) {           // thread per column over the slab's rows
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int i = 0;
        for (; i + 3 < nloc; i += 4) {
          s0 = fmaf(Ks[(size_t)i * Cs + j], a_new[i], s0);
          s1 = fmaf(Ks[(size_t)(i + 1) * Cs + j], a_new[i + 1], s1);
          s2 = fmaf(Ks[(size_t)(i + 2) * Cs + j], a_new[i + 2], s2);
          s3 = fmaf(Ks[(size_t)(i + 3) * Cs + j], a_new[i + 3], s3);
        }
        for (; i < nloc; i++) s0 = fmaf(Ks[(size_t)i * Cs + j], a_new[i], s0);
        const float part = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int r = 0; r < CL; r++) peer_exch[r][ex_off + j] = part;
      }
      if (tid < CL) peer_flags[tid][(x & 1) * CL + rank] = row_bad;
      cluster.sync();
      bad = false;
      chg = false;
      for (int j = tid; j < C; j += nt) {
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < CL; r++) sum += ex[(size_t)r * Cs + j];
        const float n_ = nu[j];
        float bn = 0.f;
        if (n_ > 0.f) {
          bn = __fdividef(n_, sum);
          bad |= !(bn < kHi && bn > kLo);
        }
        chg |= bn != b_new[j];
        b_new[j] = bn;
      }
#pragma unroll
      for (int r = 0; r < CL; r++) {
        bad |= (fl[r] & 1) != 0;
        chg |= (fl[r] & 2) != 0;
      }
      const bool fail = __syncthreads_or(bad) != 0;
      const bool moving = __syncthreads_or(chg) != 0;       // identical on every rank (replicated inputs)
      x++;
      if (!fail) {
        p ^= 1;
        it++;
        streak++;
        n_lin++;
        if (!moving && a.early_exit) {   // period-1 / period-2 cycle: exact early exit (see the point-level kernel)
          if ((a.iters - it) & 1) p ^= 1;
          it = a.iters;
        }
        continue;
      }
      n_disc++;
      for (int i = tid; i < nloc; i += nt) {
        const float av = la[p * rs + i];
        if (av > 0.f) u[i] += logf(av);
      }
      for (int j = tid; j < C; j += nt) {
        const float bv = lb[p * Cs + j];
        if (bv > 0.f) v[j] += logf(bv);
      }
      __syncthreads();
      if (streak > 0) {
        need_absorb = true;
        continue;
      }
      lin = false;
      continue;            // (the exchange buffers of the next round are selected at the top of the loop)
    }
    // ---- LOG iteration straight from the input scores
    bool big = false;
    for (int i = warp; i < nloc; i += nw) {
      float sc[kClChunks];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < kClChunks; k++) {
        const int j = lane + 32 * k;
        sc[k] = j < C ? score(i, j) + v[j] : -INFINITY;
        mx = fmaxf(mx, sc[k]);
      }
      mx = lcr_warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < kClChunks; k++)
        if (lane + 32 * k < C) sum += sk_exp(sc[k] - mx);
      sum = lcr_warp_sum(sum);
      if (lane == 0) {
        const float un = log_mu[i] - (mx + logf(sum));
        if (mu[i] > 0.f) big |= !(fabsf(un - u[i]) < kBigStep);
        u[i] = un;
      }
    }
    const int row_big = __syncthreads_or(big);
    for (int j = tid; j < C; j += nt) {            // online max / sum over the slab's rows, 4 rows in flight
      float mx = -INFINITY, sum = 0.f;
      for (int i = 0; i < nloc; i += 4) {
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; q++) t[q] = i + q < nloc ? score(i + q, j) + u[i + q] : -INFINITY;
        const float m4 = fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3]));
        if (m4 > mx) {
          sum = mx > -INFINITY ? sum * sk_exp(mx - m4) : 0.f;   // (first group: nothing summed yet)
          mx = m4;
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (i + q < nloc) sum += sk_exp(t[q] - mx);
      }
#pragma unroll
      for (int r = 0; r < CL; r++) {
        peer_exch[r][ex_off + j] = mx;
        peer_exch[r][ex_off + (size_t)CL * Cs + j] = sum;
      }
    }
    if (tid < CL) peer_flags[tid][(x & 1) * CL + rank] = row_big;
    cluster.sync();
    big = false;
    for (int j = tid; j < C; j += nt) {
      float mx = -INFINITY;
#pragma unroll
      for (int r = 0; r < CL; r++) mx = fmaxf(mx, ex[(size_t)r * Cs + j]);
      float sum = 0.f;
#pragma unroll
      for (int r = 0; r < CL; r++) {
        const float pm = ex[(size_t)r * Cs + j];
        if (pm > -INFINITY) sum += ex[(size_t)(CL + r) * Cs + j] * sk_exp(pm - mx);
      }
      const float vn = log_nu[j] - (mx + logf(sum));
      if (nu[j] > 0.f) big |= !(fabsf(vn - v[j]) < kBigStep);
      v[j] = vn;
    }
#pragma unroll
    for (int r = 0; r < CL; r++) big |= fl[r] != 0;
    big = __syncthreads_or(big) != 0;
    x++;
    it++;
    n_log++;
    if (!big && it < a.iters) {
      need_absorb = true;
      lin = true;
    }
  }
  if (tid == 0 && rank == 0) {
    atomicAdd(&g_sk_stats[0], (unsigned long long)n_log);
    atomicAdd(&g_sk_stats[1], (unsigned long long)n_lin);
    atomicAdd(&g_sk_stats[2], (unsigned long long)n_disc);
    atomicAdd(&g_sk_stats[3], (unsigned long long)n_abs);
  }
  if (lin && !need_absorb) {
    for (int i = tid; i < nloc; i += nt) {
      const float av = la[p * rs + i];
      if (av > 0.f) u[i] += logf(av);
    }
    for (int j = tid; j < C; j += nt) {
      const float bv = lb[p * Cs + j];
      if (bv > 0.f) v[j] += logf(bv);
    }
  }
  __syncthreads();
  for (int i = warp; i < nloc; i += nw) {
    const float ui = u[i] - norm;
    float sc[kClChunks];
#pragma unroll
    for (int k = 0; k < kClChunks; k++) {
      const int j = lane + 32 * k;
      sc[k] = j < C ? score(i, j) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kClChunks; k++) {
      const int j = lane + 32 * k;
      if (j < C) out[(size_t)(i0 + i) * C + j] = sc[k] + ui + v[j];
    }
  }
  cluster.sync();          // no rank may exit while a peer can still write into its shared memory
}

template <int CL>
int launch_sinkhorn_cluster(const SinkhornArgs& a, int batch, int rs, size_t smem, cudaStream_t stream) {
  static LcrOncePerDevice once;
  const int dev = once.need();
  if (dev != -1) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(sinkhorn_cluster_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kClSmemMax));
    if (CL > 8)
      LCR_CUDA_TRY(cudaFuncSetAttribute(sinkhorn_cluster_kernel<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    once.done(dev);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)batch * CL);
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LCR_CUDA_TRY(cudaLaunchKernelEx(&cfg, sinkhorn_cluster_kernel<CL>, a, rs));
  return LCR_OK;
}

}  // namespace

// Node-level kernel selection: 0 single-CTA kernel, 1 (default) smallest cluster that fits, 4 / 8 that cluster size.
// LCR_SINKHORN_CLUSTER sets the default; lcr_set_sinkhorn_cluster overrides it (tests).
static int g_sinkhorn_early_exit = 1;
extern "C" void lcr_set_sinkhorn_early_exit(int on) { g_sinkhorn_early_exit = on ? 1 : 0; }
static int g_sinkhorn_cluster = -1;
static int sinkhorn_cluster_mode() {
  if (g_sinkhorn_cluster < 0) g_sinkhorn_cluster = getenv("LCR_SINKHORN_CLUSTER") ? atoi(getenv("LCR_SINKHORN_CLUSTER")) : 1;
  return g_sinkhorn_cluster;
}
extern "C" void lcr_set_sinkhorn_cluster(int mode) { g_sinkhorn_cluster = (mode == 0 || mode == 4 || mode == 8) ? mode : 1; }

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
  SinkhornArgs a{scores, row_mask, col_mask, alpha, out, rows, cols, iters, g_sinkhorn_early_exit};
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
    const int R = rows + 1, C = cols + 1;
    // 0: single-CTA kernel; otherwise the smallest cluster (4, then 8 CTAs per problem) whose row slab fits
    const int use_cluster = sinkhorn_cluster_mode();
    int cl = 0;
    for (int c : {4, 8}) {
      if (use_cluster > 1 && c != use_cluster) continue;       // 4 / 8: force that cluster size (tests)
      if (C <= 32 * kClChunks && sizeof(float) * cluster_smem_floats(c, C, (R + c - 1) / c) <= kClSmemMax) {
        cl = c;
        break;
      }
    }
    if (use_cluster && cl) {
      const int rs = (R + cl - 1) / cl;
      const size_t cl_smem = sizeof(float) * cluster_smem_floats(cl, C, rs);
      const int rc = cl == 4 ? launch_sinkhorn_cluster<4>(a, batch, rs, cl_smem, stream)
                             : launch_sinkhorn_cluster<8>(a, batch, rs, cl_smem, stream);
      if (rc != LCR_OK) return rc;
    } else {
      const size_t smem = sizeof(float) * (5 * (size_t)R + 5 * (size_t)C + 2 * (kGenThreads / 32) * (size_t)C) + 64;
      LCR_REQUIRE(smem <= 200 * 1024, "sinkhorn: problem too large (5 (rows + cols) + 64 cols floats of shared memory)");
      LCR_CUDA_TRY(cudaFuncSetAttribute(sinkhorn_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sinkhorn_general_kernel<<<batch, kGenThreads, smem, stream>>>(a);
    }
  }
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
