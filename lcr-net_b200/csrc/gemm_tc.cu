// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMM with fp32-class accuracy by 3xTF32 splitting:
//     C[M,N] = act(rowscale[m] * (A[M,K] . W[N,K]^T) + bias[n])
// A = activations (row-major, K contiguous), W = weights in nn.Linear layout [N, K] (K contiguous),
// i.e. both operands are K-major for the MMA.
//
// Why split: tensor cores have no fp32 MMA; kind::tf32 keeps 10 mantissa bits.  Each operand is
// split on the fly into hi = rna_tf32(x) and lo = x - hi, and three MMAs accumulate
// hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator (the dropped lo*lo term is ~2^-22
// relative), which keeps the encoder inside the 1e-4 descriptor parity bar.
//
// Structure (one 128 x BN output tile per CTA, BK = 32 fp32 = one 128-byte swizzle row):
//   warps 0-7  producers: coalesced 16-B global loads -> split -> st.shared into the canonical
//              K-major SWIZZLE_128B layout (chunk ^= row & 7), fence.proxy.async, mbarrier arrive
//   warp 8     one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (4 k-steps x 3 products
//              per stage), tcgen05.commit releases the stage / signals the epilogue
//   warps 0-7  epilogue: tcgen05.ld 32x32b.x32 from TMEM, row scale + bias (+ReLU), 128-B row stores
// 3-stage shared-memory ring; accumulator: BN fp32 columns of TMEM.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 32;
constexpr int kProducerThreads = 256;
constexpr int kThreads = kProducerThreads + 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 | LBO (=1, unused for swizzled K-major) << 16 | SBO (1024 B between 8-row
// groups) >> 4 << 32 | version 1 << 46 | layout SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Round-to-nearest fp32 -> tf32 (10-bit mantissa) as two integer ops: add half a tf32 ulp to the
// magnitude bits, clear the 13 low bits.  (cvt.rna.tf32.f32 expands to ~5 instructions with
// inf/nan guards on sm_100a and made the producers ALU-bound; activations and weights are finite.)
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ void split_store(float4 v, float* hi_tile, float* lo_tile, int row, int chunk) {
  float4 h, l;
  h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
  // the residual is exact in fp32; it is rounded (not left to the MMA's operand truncation) so the
  // low-order products carry no bias
  l.x = tf32_rn(v.x - h.x); l.y = tf32_rn(v.y - h.y); l.z = tf32_rn(v.z - h.z); l.w = tf32_rn(v.w - h.w);
  const int off = row * 32 + ((chunk ^ (row & 7)) << 2);  // floats: 128-B rows, 16-B chunks XOR-swizzled
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// GroupNorm statistics fused into the epilogue (the GEMM output is always followed by a GroupNorm in
// the encoder: modules.py:53-83,104-146): while a warp streams its 32 x 32 chunk of finished values
// out, lane = column, it also sums the column over the chunk's rows.  Rows are cut into 32-row
// blocks (aligned to the tile); a block that straddles a stack boundary reports two partials:
//   gn_partial[(block * 2 + 0) * N + col] = (sum, sumsq) over the rows of the block's FIRST stack,
//   gn_partial[(block * 2 + 1) * N + col] = the same over the remaining rows (written only when the
//                                           block straddles; every stack must have >= 32 rows),
// combined per (stack, group) in a fixed order by gn_finalize_blocks_kernel (encoder.cu): no float
// atomics, run-to-run deterministic.  The 32-row sums are fp32 (chains of <= 32 terms), the combination
// fp64.  (fp64 in-loop sums were measured: +0.6 ms per step on the 4 epilogue warps, more than half of
// what the fusion saves.)  Where the block boundaries fall inside a stack depends on the stack's position
// in the batch, so the statistics of a stack evaluated inside a batch can differ from the same stack
// evaluated alone in the last fp32 bits (~1e-7 relative), like any blocked summation.
struct GnFuse {
  float2* partial;              // nullptr: no statistics
  const int64_t* stack_off;     // [S + 1] row offsets of the stacks
  int S;
};

// Stores one 32-row x 32-column chunk held transposed in tb (tb[row * 33 + col], row scale applied):
// adds the bias, applies ReLU, writes 128-byte row segments, and accumulates the column statistics.
__device__ __forceinline__ void store_chunk(const float* tb, int lane, int row0, int M, int gn, int N,
                                            const float* __restrict__ bias, int relu, float* __restrict__ C, int ldc,
                                            const GnFuse& gnf, int split) {
  const int nrows = min(32, M - row0);
  if (gn >= N || nrows <= 0) return;
  const float bv = bias ? bias[gn] : 0.f;
  float* out = C + (size_t)row0 * ldc + gn;
  const int cut = min(split, nrows);
  if (!gnf.partial) {
#pragma unroll 8
    for (int rr = 0; rr < nrows; rr++) {
      float o = tb[rr * 33 + lane] + bv;
      if (relu) o = fmaxf(o, 0.f);
      out[(size_t)rr * ldc] = o;        // 32 lanes -> one 128-byte row segment
    }
    return;
  }
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll 8
  for (int rr = 0; rr < cut; rr++) {
    float o = tb[rr * 33 + lane] + bv;
    if (relu) o = fmaxf(o, 0.f);
    out[(size_t)rr * ldc] = o;
    s0 += o;
    q0 = fmaf(o, o, q0);
  }
#pragma unroll 4
  for (int rr = cut; rr < nrows; rr++) {
    float o = tb[rr * 33 + lane] + bv;
    if (relu) o = fmaxf(o, 0.f);
    out[(size_t)rr * ldc] = o;
    s1 += o;
    q1 = fmaf(o, o, q1);
  }
  float2* p = gnf.partial + ((size_t)(row0 >> 5) * 2) * N + gn;
  p[0] = make_float2(s0, q0);
  if (cut < nrows) p[N] = make_float2(s1, q1);
}

// rows of the 32-row block starting at row0 that belong to the block's first stack
__device__ __forceinline__ int gn_split(const GnFuse& gnf, int row0, int M) {
  if (!gnf.partial || row0 >= M) return 32;
  const int s = lcr_find_segment(gnf.stack_off, gnf.S, (int64_t)row0);
  return (int)min((int64_t)32, gnf.stack_off[s + 1] - (int64_t)row0);
}

// The tensor core accumulates with truncation, which biases long K chains (measured: error grows
// linearly with K, 2.7e-5 at K = 3840).  The K loop is therefore cut into chunks of STAGES
// k-blocks (96 values of K): each chunk is accumulated in one of two TMEM buffers and then
// PROMOTED: added, with IEEE rounding, into fp32 register accumulators by the producer warps
// while the MMA warp already works on the next chunk in the other buffer.
// PS: the weights arrive PRE-SPLIT (W = tf32 hi part, W_lo = tf32 lo part, same layout; made once per
// weight tensor by lcr_tf32_split): both halves of a B tile are then plain cp.async copies and the
// producers only split the activations.
template <int BN, int STAGES, bool PS, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
gemm_tf32x3_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ W_lo,
                   int ldw, float* __restrict__ C, int ldc, int M, int N, int K, const float* __restrict__ rowscale,
                   const float* __restrict__ bias, int relu, const GnFuse gnf) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned tiles (SWIZZLE_128B atom = 8 rows x 128 B)
  // aligned by OFFSET arithmetic on the shared array (a uintptr_t round trip hides the address space from the
  // compiler: every access through `base` became a generic LD/ST that it also had to order against global stores)
  uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int kATile = BM * BK * 4, kBTile = BN * BK * 4;
  constexpr int kStageBytes = 2 * kATile + 2 * kBTile;
  constexpr int CH = STAGES;                               // k-blocks per promotion chunk
  constexpr int kColsPerWarp = BN >= 64 ? BN / 2 : BN;     // accumulator columns owned by a thread
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nk = K / BK;
  const int n_chunks = (nk + CH - 1) / CH;
  const bool promoter = BN >= 64 || warp < 4;              // BN = 32: warps 0-3 hold the accumulators
  constexpr int kPromoters = BN >= 64 ? kProducerThreads : 128;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], kProducerThreads / 32);   // one arrive per producer warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kPromoters / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 8) {
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    const int c_begin = BN >= 64 ? (warp >> 2) * kColsPerWarp : 0;
    float acc[kColsPerWarp];
#pragma unroll
    for (int j = 0; j < kColsPerWarp; j++) acc[j] = 0.f;
    auto promote = [&](int chunk) {
      const int b = chunk & 1;
      mbar_wait(&acc_full[b], (chunk >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + c_begin + c0), r);
#pragma unroll
        for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(r[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    };
    // Staging: every thread owns a fixed set of 16-byte chunks of the A and W tiles.  The raw fp32
    // chunks are fetched with cp.async (LDGSTS, zero-fill outside the matrix) straight into their
    // swizzled slot of the HI tile of a stage, STAGES-1 k-blocks ahead and without holding
    // registers; when its own copies have landed the thread converts IN PLACE: hi = rna_tf32(x)
    // back into the same slot, lo = rna_tf32(x - hi) into the LO tile.  No cross-thread
    // dependency exists before the full-barrier arrive.
    constexpr int kALoads = BM * 8 / kProducerThreads, kBLoads = BN * 8 / kProducerThreads;
    auto issue_block = [&](int kb) {
      if (kb < nk) {
        const int s = kb % STAGES;
        const uint32_t a_hi = smem_u32(base + s * kStageBytes), b_hi = a_hi + 2 * kATile;
        const int k0 = kb * BK;
#pragma unroll
        for (int i = 0; i < kALoads; i++) {
          const int idx = tid + i * kProducerThreads;
          const int row = idx >> 3, chunk = idx & 7;
          const int gm = m0 + row;
          const float* src = A + (size_t)(gm < M ? gm : 0) * lda + k0 + chunk * 4;
          const uint32_t dst = a_hi + row * 128 + ((chunk ^ (row & 7)) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(gm < M ? 16 : 0));
        }
#pragma unroll
        for (int i = 0; i < kBLoads; i++) {
          const int idx = tid + i * kProducerThreads;
          const int row = idx >> 3, chunk = idx & 7;
          const int gn = n0 + row;
          const size_t goff = (size_t)(gn < N ? gn : 0) * ldw + k0 + chunk * 4;
          const uint32_t dst = b_hi + row * 128 + ((chunk ^ (row & 7)) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(W + goff), "r"(gn < N ? 16 : 0));
          if (PS)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + kBTile), "l"(W_lo + goff),
                         "r"(gn < N ? 16 : 0));
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // always: keeps the group count uniform
    };
#pragma unroll
    for (int kb = 0; kb < STAGES - 1; kb++) issue_block(kb);
    for (int c = 0; c < n_chunks; c++) {
      // ---------------------------------------------------------- produce the chunk's k-blocks
      const int kb_end = min((c + 1) * CH, nk);
      for (int kb = c * CH; kb < kb_end; kb++) {
        const int s = kb % STAGES;
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");   // this thread's block kb landed
        float* a_hi = reinterpret_cast<float*>(base + s * kStageBytes);
        float* a_lo = a_hi + BM * BK;
        float* b_hi = a_lo + BM * BK;
        float* b_lo = b_hi + BN * BK;
#pragma unroll
        for (int i = 0; i < kALoads; i++) {
          const int idx = tid + i * kProducerThreads;
          const int row = idx >> 3, chunk = idx & 7;
          const float4 v = *reinterpret_cast<const float4*>(a_hi + row * 32 + ((chunk ^ (row & 7)) << 2));
          split_store(v, a_hi, a_lo, row, chunk);
        }
        if (!PS) {
#pragma unroll
          for (int i = 0; i < kBLoads; i++) {
            const int idx = tid + i * kProducerThreads;
            const int row = idx >> 3, chunk = idx & 7;
            const float4 v = *reinterpret_cast<const float4*>(b_hi + row * 32 + ((chunk ^ (row & 7)) << 2));
            split_store(v, b_hi, b_lo, row, chunk);
          }
        }
        // make the generic-proxy writes visible to the tensor-core (async) proxy, then signal
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);   // 8 arrivals per stage instead of 256 shared-memory atomics
        // refill the stage that block kb-1 occupied (free once its MMAs have completed)
        const int nb = kb + STAGES - 1;
        if (nb < nk && kb >= 1) mbar_wait(&empty_bar[nb % STAGES], ((nb / STAGES) - 1) & 1);
        issue_block(nb);
      }
      // ---------------------------------------------------------- promote the previous chunk
      if (c >= 1 && promoter) promote(c - 1);
    }
    if (promoter) {
      promote(n_chunks - 1);
      // -------------------------------------------------------- epilogue from registers
      // All MMAs have completed (acc_full of the last chunk), so the stage ring is free: every warp
      // transposes its 32 x 32 chunks through a private 32 x 33 buffer in it and writes 128-byte row
      // segments (+ the fused GroupNorm column statistics, see store_chunk).
      float* tb = reinterpret_cast<float*>(base) + warp * (32 * 33);
      const int row0 = m0 + q * 32;
      const int gm = row0 + lane;
      const float rs = (rowscale && gm < M) ? rowscale[gm] : 1.f;
      const int split = gn_split(gnf, row0, M);
#pragma unroll
      for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
#pragma unroll
        for (int j = 0; j < 32; j++) tb[lane * 33 + j] = acc[c0 + j] * rs;
        __syncwarp();
        store_chunk(tb, lane, row0, M, n0 + c_begin + c0 + lane, N, bias, relu, C, ldc, gnf, split);
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ MMA issuer (one lane)
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      for (int c = 0; c < n_chunks; c++) {
        const int b = c & 1;
        if (c >= 2) {
          mbar_wait(&acc_empty[b], ((c >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * BN);
        const int kb_end = min((c + 1) * CH, nk);
        for (int kb = c * CH; kb < kb_end; kb++) {
          const int s = kb % STAGES;
          mbar_wait(&full_bar[s], (kb / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(base + s * kStageBytes);
          const uint32_t a_lo = a_hi + kATile, b_hi = a_lo + kATile, b_lo = b_hi + kBTile;
#pragma unroll
          for (int k = 0; k < BK / 8; k++) {
            const uint32_t koff = k * 32;  // 8 tf32 = 32 bytes inside the 128-B swizzle row
            mma_tf32(tmem_d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, (kb > c * CH) || k != 0);
            mma_tf32(tmem_d, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc, 1);
            mma_tf32(tmem_d, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc, 1);
          }
          // frees the stage once the MMAs above have consumed it (implies fence::before_thread_sync)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(&empty_bar[s]))
                       : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&acc_full[b]))
                     : "memory");
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
  }
}

// ---------------------------------------------------------------------------------------------
// Small-K variant (K <= 128: the unary layers).  These GEMMs are bandwidth-bound (a 128 x BN output
// tile is produced from at most four k-blocks), so the kernel is built for OCCUPANCY instead of
// pipelining: one single-use stage per k-block (no ring), accumulators stay in TMEM (one chunk, no
// promotion), ~64 registers, so 2-3 CTAs share an SM and hide each other's prologue / epilogue.
// The epilogue goes TMEM -> registers -> per-warp shared-memory transpose -> 128-byte row stores.
template <int BN, bool PS>
__global__ void __launch_bounds__(kThreads, BN >= 256 ? 2 : 3)
gemm_tf32x3_small_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                         const float* __restrict__ W_lo, int ldw, float* __restrict__ C, int ldc, int M, int N, int K,
                         const float* __restrict__ rowscale, const float* __restrict__ bias, int relu,
                         const GnFuse gnf) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET arithmetic on the shared array (a uintptr_t round trip hides the address space from the
  // compiler: every access through `base` became a generic LD/ST that it also had to order against global stores)
  uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int kATile = BM * BK * 4, kBTile = BN * BK * 4;
  __shared__ uint64_t full_bar, empty_bar, done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nk = K / BK;
  if (tid == 0) {
    mbar_init(&full_bar, kProducerThreads / 32);
    mbar_init(&empty_bar, 1);
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  float* a_hi = reinterpret_cast<float*>(base);
  float* a_lo = a_hi + BM * BK;
  float* b_hi = a_lo + BM * BK;
  float* b_lo = b_hi + BN * BK;

  if (warp < 8) {
    constexpr int kALoads = BM * 8 / kProducerThreads, kBLoads = BN * 8 / kProducerThreads;
    for (int kb = 0; kb < nk; kb++) {
      const int k0 = kb * BK;
      float4 ra[kALoads], rb[kBLoads], rbl[PS ? kBLoads : 1];
#pragma unroll
      for (int i = 0; i < kALoads; i++) {
        const int idx = tid + i * kProducerThreads;
        const int gm = m0 + (idx >> 3);
        ra[i] = gm < M ? *reinterpret_cast<const float4*>(A + (size_t)gm * lda + k0 + (idx & 7) * 4)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < kBLoads; i++) {
        const int idx = tid + i * kProducerThreads;
        const int gn = n0 + (idx >> 3);
        rb[i] = gn < N ? *reinterpret_cast<const float4*>(W + (size_t)gn * ldw + k0 + (idx & 7) * 4)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        if (PS)
          rbl[i] = gn < N ? *reinterpret_cast<const float4*>(W_lo + (size_t)gn * ldw + k0 + (idx & 7) * 4)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (kb >= 1) mbar_wait(&empty_bar, (kb - 1) & 1);   // the MMAs of block kb-1 have read the stage
#pragma unroll
      for (int i = 0; i < kALoads; i++) {
        const int idx = tid + i * kProducerThreads;
        split_store(ra[i], a_hi, a_lo, idx >> 3, idx & 7);
      }
#pragma unroll
      for (int i = 0; i < kBLoads; i++) {
        const int idx = tid + i * kProducerThreads;
        if (PS) {
          const int row = idx >> 3, chunk = idx & 7;
          const int off = row * 32 + ((chunk ^ (row & 7)) << 2);
          *reinterpret_cast<float4*>(b_hi + off) = rb[i];
          *reinterpret_cast<float4*>(b_lo + off) = rbl[i];
        } else {
          split_store(rb[i], b_hi, b_lo, idx >> 3, idx & 7);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar);
    }
    // ------------------------------------------------------------ epilogue
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // the stage is free now: per-warp 32 x 33 transpose buffers live in it
    float* tb = reinterpret_cast<float*>(base) + warp * (32 * 33);
    const int q = warp & 3;
    constexpr int kColsPerWarp = BN >= 64 ? BN / 2 : BN;
    const int c_begin = BN >= 64 ? (warp >> 2) * kColsPerWarp : 0;
    if (BN >= 64 || warp < 4) {
      const int row0 = m0 + q * 32;
      const float rs = (rowscale && row0 + lane < M) ? rowscale[row0 + lane] : 1.f;
      const int split = gn_split(gnf, row0, M);
#pragma unroll 1
      for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c_begin + c0), r);
#pragma unroll
        for (int j = 0; j < 32; j++) tb[lane * 33 + j] = __uint_as_float(r[j]) * rs;
        __syncwarp();
        store_chunk(tb, lane, row0, M, n0 + c_begin + c0 + lane, N, bias, relu, C, ldc, gnf, split);
        __syncwarp();
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      const uint32_t sa_hi = smem_u32(base), sa_lo = sa_hi + kATile, sb_hi = sa_lo + kATile, sb_lo = sb_hi + kBTile;
      for (int kb = 0; kb < nk; kb++) {
        mbar_wait(&full_bar, kb & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < BK / 8; k++) {
          const uint32_t koff = k * 32;
          mma_tf32(tmem_base, make_desc(sa_hi + koff), make_desc(sb_hi + koff), idesc, (kb | k) != 0);
          mma_tf32(tmem_base, make_desc(sa_hi + koff), make_desc(sb_lo + koff), idesc, 1);
          mma_tf32(tmem_base, make_desc(sa_lo + koff), make_desc(sb_hi + koff), idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(kb + 1 < nk ? &empty_bar : &done_bar))
                     : "memory");
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent, warp-specialised variant (default for pre-split weights).  The one-tile-per-CTA kernels
// above serialise load -> split -> MMA -> epilogue inside a CTA; for the bandwidth-bound layers (every
// unary layer, the narrow KPConv contractions) that leaves HBM idle during every prologue and epilogue
// (ncu: 37 % warps active, long-scoreboard + barrier stalls, 2.3 TB/s).  Here one CTA per SM walks over
// output tiles and three roles run concurrently:
//   warps 0-7   producers: cp.async (LDGSTS) of the raw fp32 A block and the pre-split W block into a
//               STAGES-deep ring that runs ACROSS tile boundaries (the loads of the next tiles are in
//               flight while this tile is multiplied and stored); the raw A block is used as the hi
//               operand as it lands (kind::tf32 reads only the upper 19 bits of an fp32 word, i.e. the
//               tensor core truncates), so the producers only compute lo = x - trunc(x) (exact in fp32)
//               into the lo block: one LOP3 + one FADD per element.
//   warp 8      one lane issues the 3 x (BK/8) tcgen05.mma per block; tcgen05.commit frees the stage
//               and, per promotion chunk, hands the TMEM buffer to the epilogue warps.
//   warps 9-12  epilogue: promote each finished chunk (TMEM -> fp32 registers, see above), and after a
//               tile's last chunk release the TMEM buffer and stream the tile out (row scale, bias, ReLU,
//               fused GroupNorm statistics) while the MMA warp is already in the next tile.
// Two TMEM buffers of BN columns alternate per chunk.
constexpr int kWsProducerWarps = 8, kWsEpilogueWarps = 4;
constexpr int kWsThreads = (kWsProducerWarps + 1 + kWsEpilogueWarps) * 32;
constexpr int kWsChunk = 3;   // k-blocks per promotion chunk (96 values of K)
// 416 threads start with 128 registers each (53 K of the 64 K file): the epilogue warpgroup grows to 200
// and the two producer warpgroups shrink to 72.  The register file is split between the four SM sub-partitions
// (16 K each; warp w lives in partition w % 4): partition 0 holds an epilogue warp, two producer warps and the MMA
// warp, so the growth (72 x 32) has to be paid for by the two producers on the same partition (2 x 56 x 32).
constexpr int kWsEpilogueRegs = 200, kWsProducerRegs = 72;

template <int BN, int STAGES>
__global__ void __launch_bounds__(kWsThreads, 1)
gemm_tf32x3_ws_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ W_lo,
                      int ldw, float* __restrict__ C, int ldc, int M, int N, int K, const float* __restrict__ rowscale,
                      const float* __restrict__ bias, int relu, const GnFuse gnf) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET arithmetic on the shared array (a uintptr_t round trip hides the address space from the
  // compiler: every access through `base` became a generic LD/ST that it also had to order against global stores)
  uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int kATile = BM * BK * 4, kBTile = BN * BK * 4;
  constexpr int kStageBytes = 2 * kATile + 2 * kBTile;
  constexpr int kProd = kWsProducerWarps * 32;
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = K / BK;
  const int tiles_n = (N + BN - 1) / BN;
  const int tiles = ((M + BM - 1) / BM) * tiles_n;
  const int my_tiles = ((int)blockIdx.x < tiles) ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_chunks = (nk + kWsChunk - 1) / kWsChunk;     // per tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], kWsProducerWarps);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kWsEpilogueWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr int kMmaWarp = kWsEpilogueWarps + kWsProducerWarps;
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp >= kWsEpilogueWarps && warp < kMmaWarp) {
    // ================================================================ producers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsProducerRegs));
    const int tid = (int)threadIdx.x - kWsEpilogueWarps * 32;   // index among the producer threads
    const int total = my_tiles * nk;                       // blocks this CTA streams, tile-major
    constexpr int kALoads = BM * 8 / kProd, kBLoads = (BN * 8 + kProd - 1) / kProd;
    auto issue_block = [&](int g) {
      if (g < total) {
        const int t = (int)blockIdx.x + (g / nk) * (int)gridDim.x;
        const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN, k0 = (g % nk) * BK;
        const int s = g % STAGES;
        const uint32_t a_hi = smem_u32(base + s * kStageBytes), b_hi = a_hi + 2 * kATile;
#pragma unroll
        for (int i = 0; i < kALoads; i++) {
          const int idx = tid + i * kProd;
          const int row = idx >> 3, chunk = idx & 7;
          const int gm = m0 + row;
          const float* src = A + (size_t)(gm < M ? gm : 0) * lda + k0 + chunk * 4;
          const uint32_t dst = a_hi + row * 128 + ((chunk ^ (row & 7)) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(gm < M ? 16 : 0));
        }
#pragma unroll
        for (int i = 0; i < kBLoads; i++) {
          const int idx = tid + i * kProd;
          if (BN * 8 % kProd == 0 || idx < BN * 8) {
            const int row = idx >> 3, chunk = idx & 7;
            const int gn = n0 + row;
            const size_t goff = (size_t)(gn < N ? gn : 0) * ldw + k0 + chunk * 4;
            const uint32_t dst = b_hi + row * 128 + ((chunk ^ (row & 7)) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(W + goff), "r"(gn < N ? 16 : 0));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + kBTile), "l"(W_lo + goff),
                         "r"(gn < N ? 16 : 0));
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // always: keeps the group count uniform
    };
#pragma unroll
    for (int g = 0; g < STAGES - 1; g++) issue_block(g);
    for (int g = 0; g < total; g++) {
      const int s = g % STAGES;
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");   // this thread's part of block g landed
      float* a_hi = reinterpret_cast<float*>(base + s * kStageBytes);
      float* a_lo = a_hi + BM * BK;
#pragma unroll
      for (int i = 0; i < kALoads; i++) {
        const int idx = tid + i * kProd;
        const int row = idx >> 3, chunk = idx & 7;
        const int off = row * 32 + ((chunk ^ (row & 7)) << 2);
        const float4 v = *reinterpret_cast<const float4*>(a_hi + off);
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        *reinterpret_cast<float4*>(a_lo + off) = l;
      }
      // generic-proxy writes (lo block) and this thread's landed cp.async data -> visible to the async proxy
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      const int nb = g + STAGES - 1;                       // refill the stage block g-1 occupied
      if (nb < total && g >= 1) mbar_wait(&empty_bar[nb % STAGES], ((nb / STAGES) - 1) & 1);
      issue_block(nb);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == kMmaWarp) {
    // ================================================================ MMA issuer (one lane)
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      int g = 0, gc = 0;                                   // running block / chunk counters
      for (int ti = 0; ti < my_tiles; ti++) {
        for (int c = 0; c < n_chunks; c++, gc++) {
          const int b = gc & 1;
          if (gc >= 2) {
            mbar_wait(&acc_empty[b], ((gc >> 1) - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          const uint32_t tmem_d = tmem_base + (uint32_t)(b * BN);
          const int kb_end = min((c + 1) * kWsChunk, nk);
          for (int kb = c * kWsChunk; kb < kb_end; kb++, g++) {
            const int s = g % STAGES;
            mbar_wait(&full_bar[s], (g / STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(base + s * kStageBytes);
            const uint32_t a_lo = a_hi + kATile, b_hi = a_lo + kATile, b_lo = b_hi + kBTile;
#pragma unroll
            for (int k = 0; k < BK / 8; k++) {
              const uint32_t koff = k * 32;
              mma_tf32(tmem_d, make_desc(a_hi + koff), make_desc(b_hi + koff), idesc, (kb > c * kWsChunk) || k != 0);
              mma_tf32(tmem_d, make_desc(a_hi + koff), make_desc(b_lo + koff), idesc, 1);
              mma_tf32(tmem_d, make_desc(a_lo + koff), make_desc(b_hi + koff), idesc, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(&empty_bar[s]))
                         : "memory");
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(&acc_full[b]))
                       : "memory");
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================ epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsEpilogueRegs));
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    float* tb = reinterpret_cast<float*>(base + STAGES * kStageBytes) + warp * (32 * 33);
    int gc = 0;
    for (int ti = 0; ti < my_tiles; ti++) {
      const int t = (int)blockIdx.x + ti * (int)gridDim.x;
      const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; j++) acc[j] = 0.f;
      for (int c = 0; c < n_chunks; c++, gc++) {
        const int b = gc & 1;
        mbar_wait(&acc_full[b], (gc >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + c0), r);
#pragma unroll
          for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[b]);
      }
      const int row0 = m0 + q * 32;
      const float rs = (rowscale && row0 + lane < M) ? rowscale[row0 + lane] : 1.f;
      const int split = gn_split(gnf, row0, M);
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
#pragma unroll
        for (int j = 0; j < 32; j++) tb[lane * 33 + j] = acc[c0 + j] * rs;
        __syncwarp();
        store_chunk(tb, lane, row0, M, n0 + c0 + lane, N, bias, relu, C, ldc, gnf, split);
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent kernel with the A operand in TENSOR MEMORY (tcgen05.mma "TS" form: D += A[tmem] . B[smem]^T).
// Why: in the kernel above every byte passes through shared memory several times -- per 128 x 32 k-block the
// tensor core itself reads A and B once per MMA (3 products x 4 k-steps x (4 KB of A + BN x 32 B of B): 96 KB at
// BN = 128), the cp.async engine writes 48 KB, and the lo-split reads and writes the A block again (32 KB): 176 KB
// per 768 MMA cycles against a 128 B/clk shared-memory port -- the tensor pipe cannot pass ~55 % (ncu: 48-52 %),
// and the narrow contractions (BN = 32: 60 KB of A re-reads per 192 MMA cycles) are worse.  Here the producers
// move each landed raw A block ONCE from shared memory into TMEM (tcgen05.st), already split: hi = the raw word
// (kind::tf32 truncates it), lo = x - trunc(x); the MMAs read A from TMEM (which has its own read path) and only
// the W blocks from shared memory: 112 KB per k-block at BN = 128, 52 KB at BN = 32.
//   TMEM columns: 2 accumulators x BN  +  STAGES x (32 hi + 32 lo) A columns  <= 512.
//   a producer warp may touch TMEM lanes 32 (warp % 4) .. + 31: thread = A row, all 32 columns of the k-block; its
//   128 bytes of the row come from eight conflict-free 16-byte loads of the swizzled landing buffer.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// TMA: the operand blocks are fetched by the tensor-memory-accelerator (cp.async.bulk.tensor.2d through 2-D tensor maps
// with the 128-byte swizzle, one elected thread per producer group issues the three copies of a k-block and arms
// the stage's transaction barrier) instead of 16-byte cp.async by every producer thread.
struct TmaMaps {
  CUtensorMap a, w, w_lo;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

template <int BN, int STAGES, bool TMA>
__global__ void __launch_bounds__(kWsThreads, 1)
gemm_tf32x3_ts_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ W_lo,
                      int ldw, float* __restrict__ C, int ldc, int M, int N, int K, const float* __restrict__ rowscale,
                      const float* __restrict__ bias, int relu, const GnFuse gnf,
                      const __grid_constant__ TmaMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  constexpr int kATile = BM * BK * 4, kBTile = BN * BK * 4;
  constexpr int kStageBytes = kATile + 2 * kBTile;          // raw A landing buffer | W hi | W lo
  constexpr int kProd = kWsProducerWarps * 32;
  constexpr uint32_t kAccCols = 2 * BN;                     // TMEM: [0, 2 BN) accumulators, then the A stages
  static_assert(kAccCols + STAGES * 64 <= 512, "TMEM budget");
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2], tma_bar[STAGES];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = K / BK;
  const int tiles_n = (N + BN - 1) / BN;
  const int tiles = ((M + BM - 1) / BM) * tiles_n;
  const int my_tiles = ((int)blockIdx.x < tiles) ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_chunks = (nk + kWsChunk - 1) / kWsChunk;     // per tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], kWsProducerWarps / 2);   // one arrive per warp of the group that owns the block
      mbar_init(&empty_bar[s], 1);
      mbar_init(&tma_bar[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kWsEpilogueWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int kMmaWarp = kWsEpilogueWarps + kWsProducerWarps;
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp >= kWsEpilogueWarps && warp < kMmaWarp) {
    // ================================================================ producers
    // Two independent groups of four warps, group j streaming the k-blocks g = j (mod 2).  A group's loop per
    // block -- wait for its copies, barrier, 8 x LDS, split, 2 x tcgen05.st, wait::st, fences, arrive, wait for a
    // free stage, issue the next copies -- is a chain of latencies (~1000 cycles); with all eight warps in one
    // lock-step group only one block entered the pipeline per such chain and the bandwidth-bound shapes stalled at
    // ~3.6 TB/s.  Two chains in flight double the rate at which loads are issued.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsProducerRegs));
    constexpr int kGrp = 128;                                   // threads per producer group
    const int ptid = (int)threadIdx.x - kWsEpilogueWarps * 32;  // index among the producer threads
    const int grp = ptid >> 7, tid = ptid & (kGrp - 1);         // group, index inside the group
    const int total = my_tiles * nk;                            // blocks this CTA streams, tile-major
    constexpr int kALoads = BM * 8 / kGrp, kBLoads = (BN * 8 + kGrp - 1) / kGrp;
    auto issue_block = [&](int g) {
      if (TMA) {
        if (g < total && tid == 0) {
          const int t = (int)blockIdx.x + (g / nk) * (int)gridDim.x;
          const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN, k0 = (g % nk) * BK;
          const int s = g % STAGES;
          const uint32_t a_raw = smem_u32(base + s * kStageBytes), b_hi = a_raw + kATile;
          mbar_expect_tx(&tma_bar[s], (uint32_t)kStageBytes);
          tma_load_2d(a_raw, &maps.a, k0, m0, &tma_bar[s]);
          tma_load_2d(b_hi, &maps.w, k0, n0, &tma_bar[s]);
          tma_load_2d(b_hi + kBTile, &maps.w_lo, k0, n0, &tma_bar[s]);
        }
        return;
      }
      if (g < total) {
        const int t = (int)blockIdx.x + (g / nk) * (int)gridDim.x;
        const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN, k0 = (g % nk) * BK;
        const int s = g % STAGES;
        const uint32_t a_raw = smem_u32(base + s * kStageBytes), b_hi = a_raw + kATile;
#pragma unroll
        for (int i = 0; i < kALoads; i++) {
          const int idx = tid + i * kGrp;
          const int row = idx >> 3, chunk = idx & 7;
          const int gm = m0 + row;
          const float* src = A + (size_t)(gm < M ? gm : 0) * lda + k0 + chunk * 4;
          const uint32_t dst = a_raw + row * 128 + ((chunk ^ (row & 7)) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(gm < M ? 16 : 0));
        }
#pragma unroll
        for (int i = 0; i < kBLoads; i++) {
          const int idx = tid + i * kGrp;
          if (BN * 8 % kGrp == 0 || idx < BN * 8) {
            const int row = idx >> 3, chunk = idx & 7;
            const int gn = n0 + row;
            const size_t goff = (size_t)(gn < N ? gn : 0) * ldw + k0 + chunk * 4;
            const uint32_t dst = b_hi + row * 128 + ((chunk ^ (row & 7)) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(W + goff), "r"(gn < N ? 16 : 0));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + kBTile), "l"(W_lo + goff),
                         "r"(gn < N ? 16 : 0));
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // always: keeps the group count uniform
    };
    // this group's blocks: g = grp, grp + 2, ...; it keeps kAhead of them in flight (stages g % STAGES are disjoint
    // between the groups as long as 2 * kAhead <= STAGES)
    constexpr int kAhead = STAGES / 2;
    static_assert(kAhead >= 1 && 2 * kAhead <= STAGES, "stage ring too short for two producer groups");
#pragma unroll
    for (int j = 0; j < kAhead; j++) issue_block(grp + 2 * j);
    const int arow = tid;                                  // A row (= TMEM lane) this thread moves: warp quarter = warp % 4
    const uint32_t t_lane = (uint32_t)(32 * (tid >> 5)) << 16;
    for (int g = grp; g < total; g += 2) {
      const int s = g % STAGES;
      if (TMA) {
        mbar_wait(&tma_bar[s], (g / STAGES) & 1);                              // the three copies of block g landed
        __syncwarp();
      } else {
        asm volatile("cp.async.wait_group %0;" ::"n"(kAhead - 1) : "memory");   // this thread's part of block g landed
        if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(kGrp) : "memory");   // ... and the rest of the group's
        else asm volatile("bar.sync 2, %0;" ::"n"(kGrp) : "memory");
      }
      const float* a_raw = reinterpret_cast<const float*>(base + s * kStageBytes);
      const uint32_t t_a = tmem_base + kAccCols + (uint32_t)(s * 64) + t_lane;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int chunk = 4 * half + c;
          const float4 v = *reinterpret_cast<const float4*>(a_raw + arow * 32 + ((chunk ^ (arow & 7)) << 2));
          const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const uint32_t hb = __float_as_uint(x[e]);
            hi[4 * c + e] = hb;                                                // the tensor core truncates it
            lo[4 * c + e] = __float_as_uint(x[e] - __uint_as_float(hb & 0xFFFFE000u));
          }
        }
        tmem_st16(t_a + (uint32_t)(16 * half), hi);
        tmem_st16(t_a + 32 + (uint32_t)(16 * half), lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      // W blocks (cp.async, generic proxy) -> visible to the async proxy; A block in TMEM -> ordered before the arrive
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      const int nb = g + 2 * kAhead;                       // this group's next block to put in flight
      if (nb < total && nb >= STAGES && (!TMA || tid == 0)) mbar_wait(&empty_bar[nb % STAGES], ((nb / STAGES) - 1) & 1);
      issue_block(nb);
    }
    if (!TMA) asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == kMmaWarp) {
    // ================================================================ MMA issuer (one lane)
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(BM >> 4) << 24);
      int g = 0, gc = 0;                                   // running block / chunk counters
      for (int ti = 0; ti < my_tiles; ti++) {
        for (int c = 0; c < n_chunks; c++, gc++) {
          const int b = gc & 1;
          if (gc >= 2) {
            mbar_wait(&acc_empty[b], ((gc >> 1) - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          const uint32_t tmem_d = tmem_base + (uint32_t)(b * BN);
          const int kb_end = min((c + 1) * kWsChunk, nk);
          for (int kb = c * kWsChunk; kb < kb_end; kb++, g++) {
            const int s = g % STAGES;
            mbar_wait(&full_bar[s], (g / STAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t b_hi = smem_u32(base + s * kStageBytes) + kATile, b_lo = b_hi + kBTile;
            const uint32_t a_hi = tmem_base + kAccCols + (uint32_t)(s * 64), a_lo = a_hi + 32;
#pragma unroll
            for (int k = 0; k < BK / 8; k++) {
              const uint32_t koff = k * 32, kc = k * 8;
              mma_tf32_ts(tmem_d, a_hi + kc, make_desc(b_hi + koff), idesc, (kb > c * kWsChunk) || k != 0);
              mma_tf32_ts(tmem_d, a_hi + kc, make_desc(b_lo + koff), idesc, 1);
              mma_tf32_ts(tmem_d, a_lo + kc, make_desc(b_hi + koff), idesc, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(&empty_bar[s]))
                         : "memory");
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(&acc_full[b]))
                       : "memory");
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================ epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsEpilogueRegs));
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    float* tb = reinterpret_cast<float*>(base + STAGES * kStageBytes) + warp * (32 * 33);
    int gc = 0;
    for (int ti = 0; ti < my_tiles; ti++) {
      const int t = (int)blockIdx.x + ti * (int)gridDim.x;
      const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; j++) acc[j] = 0.f;
      for (int c = 0; c < n_chunks; c++, gc++) {
        const int b = gc & 1;
        mbar_wait(&acc_full[b], (gc >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + c0), r);
#pragma unroll
          for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[b]);
      }
      const int row0 = m0 + q * 32;
      const float rs = (rowscale && row0 + lane < M) ? rowscale[row0 + lane] : 1.f;
      const int split = gn_split(gnf, row0, M);
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
#pragma unroll
        for (int j = 0; j < 32; j++) tb[lane * 33 + j] = acc[c0 + j] * rs;
        __syncwarp();
        store_chunk(tb, lane, row0, M, n0 + c0 + lane, N, bias, relu, C, ldc, gnf, split);
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// 2-D fp32 tensor [rows, cols] with row stride ld (floats): boxes of 32 columns (one 128-byte swizzle row) x box_rows
bool make_map(CUtensorMap* m, const float* ptr, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

template <int BN, int STAGES>
int launch_tc_ts(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N,
                 int K, const float* rowscale, const float* bias, int relu, const GnFuse& gnf, cudaStream_t stream,
                 bool tma) {
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 4 + 2 * BN * BK * 4) + kWsEpilogueWarps * 32 * 33 * 4 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static LcrOncePerDevice attr_done;
  const int attr_done_dev = attr_done.need();
  if (attr_done_dev != -1) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_ts_kernel<BN, STAGES, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LCR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_ts_kernel<BN, STAGES, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done.done(attr_done_dev);
  }
  const long tiles = (long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const unsigned grid = (unsigned)(tiles < LCR_SM_COUNT ? tiles : LCR_SM_COUNT);
  TmaMaps maps;
  if (tma) tma = make_map(&maps.a, A, M, K, lda, BM) && make_map(&maps.w, W, N, K, ldw, BN) &&
                 make_map(&maps.w_lo, W_lo, N, K, ldw, BN);
  if (tma)
    gemm_tf32x3_ts_kernel<BN, STAGES, true><<<grid, kWsThreads, smem, stream>>>(A, lda, W, W_lo, ldw, C, ldc, M, N, K,
                                                                                rowscale, bias, relu, gnf, maps);
  else
    gemm_tf32x3_ts_kernel<BN, STAGES, false><<<grid, kWsThreads, smem, stream>>>(A, lda, W, W_lo, ldw, C, ldc, M, N, K,
                                                                                 rowscale, bias, relu, gnf, maps);
  return LCR_OK;
}

template <int BN, int STAGES>
int launch_tc_ws(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N,
                 int K, const float* rowscale, const float* bias, int relu, const GnFuse& gnf, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (2 * BM * BK * 4 + 2 * BN * BK * 4) + kWsEpilogueWarps * 32 * 33 * 4 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static LcrOncePerDevice attr_done;
  const int attr_done_dev = attr_done.need();
  if (attr_done_dev != -1) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_ws_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    attr_done.done(attr_done_dev);
  }
  const long tiles = (long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const unsigned grid = (unsigned)(tiles < LCR_SM_COUNT ? tiles : LCR_SM_COUNT);
  gemm_tf32x3_ws_kernel<BN, STAGES><<<grid, kWsThreads, smem, stream>>>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale,
                                                                         bias, relu, gnf);
  return LCR_OK;
}

template <int BN, bool PS>
int launch_tc_small(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N,
                    int K, const float* rowscale, const float* bias, int relu, const GnFuse& gnf, cudaStream_t stream) {
  constexpr size_t stage = 2 * BM * BK * 4 + 2 * BN * BK * 4;
  constexpr size_t smem = (stage > 8 * 32 * 33 * 4 ? stage : 8 * 32 * 33 * 4) + 1024;
  static LcrOncePerDevice attr_done;
  const int attr_done_dev = attr_done.need();
  if (attr_done_dev != -1) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_small_kernel<BN, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    attr_done.done(attr_done_dev);
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_tf32x3_small_kernel<BN, PS><<<grid, kThreads, smem, stream>>>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale,
                                                                     bias, relu, gnf);
  return LCR_OK;
}

template <int BN, int STAGES, bool PS, int MINB = 1>
int launch_tc(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N, int K,
              const float* rowscale, const float* bias, int relu, const GnFuse& gnf, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (2 * BM * BK * 4 + 2 * BN * BK * 4) + 1024;
  static LcrOncePerDevice attr_done;
  const int attr_done_dev = attr_done.need();
  if (attr_done_dev != -1) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, STAGES, PS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    attr_done.done(attr_done_dev);
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_tf32x3_kernel<BN, STAGES, PS, MINB><<<grid, kThreads, smem, stream>>>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale,
                                                                       bias, relu, gnf);
  return LCR_OK;
}

// hi = rna_tf32(w), lo = rna_tf32(w - hi): the operand split of the 3xTF32 scheme, done once for weights
__global__ void tf32_split_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float x = w[i], h = tf32_rn(x);
    hi[i] = h;
    lo[i] = tf32_rn(x - h);
  }
}

template <bool PS>
int gemm_dispatch(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N,
                  int K, const float* rowscale, const float* bias, int relu, const GnFuse& gnf, cudaStream_t stream) {
  static const int ws = getenv("LCR_GEMM_WS") ? atoi(getenv("LCR_GEMM_WS")) : 5;   // 5: A operand in TMEM, TMA loads (default)
  if (PS && (ws >= 3 && ws <= 6) && ((ws & 1) == 0 || K > 4 * BK)) {   // persistent kernel, A operand in tensor memory
    const bool tma = ws >= 5;                                             // 5 / 6: operand blocks fetched by TMA
    // stages: as many as tensor memory allows (2 BN accumulator columns + 64 per stage <= 512): the narrow shapes are
    // bound by the bytes in flight (ncu: nothing above 46 %, DRAM latency ~2 us under load)
    if (N <= 32) return launch_tc_ts<32, 7>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream, tma);
    if (N <= 64) return launch_tc_ts<64, 6>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream, tma);
    return launch_tc_ts<128, 4>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream, tma);
  }
  if (PS && (ws == 2 || (ws == 1 && K > 4 * BK))) {   // persistent warp-specialised kernel (needs the pre-split weights)
    if (N <= 32) return launch_tc_ws<32, 5>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
    if (N <= 64) return launch_tc_ws<64, 4>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
    return launch_tc_ws<128, 3>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
  }
  if (K <= 4 * BK) {  // bandwidth-bound unary layers: occupancy-oriented kernel
    if (N <= 32) return launch_tc_small<32, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
    if (N <= 64) return launch_tc_small<64, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
    if (N <= 128) return launch_tc_small<128, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
    return launch_tc_small<256, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
  }
  static const int occ = getenv("LCR_GEMM_OCC") ? atoi(getenv("LCR_GEMM_OCC")) : 0;
  if (occ) {   // experiment: two co-resident CTAs per SM with a 2-stage ring instead of one CTA with 4 stages
    if (N <= 32) return launch_tc<32, 2, PS, 2>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
    if (N <= 64) return launch_tc<64, 2, PS, 2>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
  }
  if (N <= 32) return launch_tc<32, 4, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
  if (N <= 64) return launch_tc<64, 4, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
  // wide tiles amortise the A-operand staging (each 128 x 32 A block is split once per N tile); when
  // 256-wide tiles would leave the last wave of the 148 SMs mostly empty, 128-wide tiles fill it better
  const long tiles_m = (M + 127) / 128;
  const long t256 = tiles_m * ((N + 255) / 256), t128 = tiles_m * ((N + 127) / 128);
  const long w256 = (t256 + LCR_SM_COUNT - 1) / LCR_SM_COUNT, w128 = (t128 + LCR_SM_COUNT - 1) / LCR_SM_COUNT;
  if (N <= 128 || w128 < 2 * w256)   // a 128-wide tile costs about half a 256-wide one
    return launch_tc<128, 3, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
  return launch_tc<256, 2, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
}

}  // namespace

// Internal entry: tensor-core GEMM.  Requirements: K % 32 == 0, N % 4 == 0, 16-byte aligned rows.
// W_lo != NULL: W / W_lo are the pre-split tf32 hi / lo parts of the weights (lcr_tf32_split).
// gn_partial != NULL: also write the per-32-row-block column statistics of the output (GnFuse).
int lcr_gemm_tf32x3_gn(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M,
                       int N, int K, const float* rowscale, const float* bias, int relu, float* gn_partial,
                       const int64_t* stack_off, int n_stacks, cudaStream_t stream) {
  LCR_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_tc: bad shape");
  LCR_REQUIRE((K % BK) == 0 && (N % 4) == 0 && (lda % 4) == 0 && (ldw % 4) == 0 && (ldc % 4) == 0,
              "gemm_tc: K must be a multiple of 32; N and leading dimensions multiples of 4");
  LCR_REQUIRE((((uintptr_t)A | (uintptr_t)W | (uintptr_t)W_lo | (uintptr_t)C | (uintptr_t)bias) & 15) == 0,
              "gemm_tc: pointers must be 16-byte aligned");
  LCR_REQUIRE(!gn_partial || (stack_off && n_stacks >= 1 && ((uintptr_t)gn_partial & 7) == 0),
              "gemm_tc: fused GroupNorm statistics need the stack offsets");
  if (M == 0) return LCR_OK;
  LcrProfScope prof("gemm_tf32x3", 2.0 * M * N * K, 4.0 * ((double)M * K + (double)K * N + (double)M * N), stream);
  GnFuse gnf{reinterpret_cast<float2*>(gn_partial), stack_off, n_stacks};
  const int rc = W_lo ? gemm_dispatch<true>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream)
                      : gemm_dispatch<false>(A, lda, W, nullptr, ldw, C, ldc, M, N, K, rowscale, bias, relu, gnf, stream);
  if (rc != LCR_OK) return rc;
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

int lcr_gemm_tf32x3(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N,
                    int K, const float* rowscale, const float* bias, int relu, cudaStream_t stream) {
  return lcr_gemm_tf32x3_gn(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, nullptr, nullptr, 0, stream);
}

// C ABI: out = act(rowscale * (x . weight^T) + bias) with weight in nn.Linear layout [c_out, c_in];
// weight_lo != NULL: weight / weight_lo are the two halves written by lcr_tf32_split.
extern "C" int lcr_linear_tc(const float* x, int64_t n_rows, int c_in, int ld_x, const float* weight,
                             const float* weight_lo, int c_out, int ld_w, const float* bias, const float* rowscale,
                             int act, float* out, int ld_out, void* stream) {
  LCR_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "linear_tc: n_rows out of range");
  return lcr_gemm_tf32x3(x, ld_x, weight, weight_lo, ld_w, out, ld_out, (int)n_rows, c_out, c_in, rowscale, bias, act,
                         (cudaStream_t)stream);
}

// lcr_linear_tc + fused GroupNorm statistics of the output (see lcr_group_norm_finalize_blocks).
extern "C" int lcr_linear_tc_gn(const float* x, int64_t n_rows, int c_in, int ld_x, const float* weight,
                                const float* weight_lo, int c_out, int ld_w, const float* bias, float* out, int ld_out,
                                float* gn_partial, const int64_t* stack_off, int n_stacks, void* stream) {
  LCR_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "linear_tc_gn: n_rows out of range");
  LCR_REQUIRE(gn_partial && ld_out == c_out, "linear_tc_gn: needs the partial buffer and a dense output");
  return lcr_gemm_tf32x3_gn(x, ld_x, weight, weight_lo, ld_w, out, ld_out, (int)n_rows, c_out, c_in, nullptr, bias, 0,
                            gn_partial, stack_off, n_stacks, (cudaStream_t)stream);
}

extern "C" int lcr_tf32_split(const float* w, int64_t count, float* hi, float* lo, void* stream) {
  LCR_REQUIRE(count >= 0, "tf32_split: bad count");
  if (count == 0) return LCR_OK;
  tf32_split_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, count, hi, lo);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
