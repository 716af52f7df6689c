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

// The tensor core accumulates with truncation, which biases long K chains (measured: error grows
// linearly with K, 2.7e-5 at K = 3840).  The K loop is therefore cut into chunks of STAGES
// k-blocks (96 values of K): each chunk is accumulated in one of two TMEM buffers and then
// PROMOTED: added, with IEEE rounding, into fp32 register accumulators by the producer warps
// while the MMA warp already works on the next chunk in the other buffer.
// PS: the weights arrive PRE-SPLIT (W = tf32 hi part, W_lo = tf32 lo part, same layout; made once per
// weight tensor by lcr_tf32_split): both halves of a B tile are then plain cp.async copies and the
// producers only split the activations.
template <int BN, int STAGES, bool PS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32x3_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ W_lo,
                   int ldw, float* __restrict__ C, int ldc, int M, int N, int K, const float* __restrict__ rowscale,
                   const float* __restrict__ bias, int relu) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned tiles (SWIZZLE_128B atom = 8 rows x 128 B)
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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
      const int gm = m0 + q * 32 + lane;
      if (gm < M) {
        const float rs = rowscale ? rowscale[gm] : 1.f;
#pragma unroll
        for (int j = 0; j < kColsPerWarp; j += 4) {
          const int gn = n0 + c_begin + j;
          if (gn < N) {
            float4 o = make_float4(acc[j] * rs, acc[j + 1] * rs, acc[j + 2] * rs, acc[j + 3] * rs);
            if (bias) {
              const float4 bv = *reinterpret_cast<const float4*>(bias + gn);
              o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
            }
            if (relu) {
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            *reinterpret_cast<float4*>(C + (size_t)gm * ldc + gn) = o;
          }
        }
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
                         const float* __restrict__ rowscale, const float* __restrict__ bias, int relu) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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
      const int row_l = q * 32 + lane;
      const float rs = (rowscale && m0 + row_l < M) ? rowscale[m0 + row_l] : 1.f;
#pragma unroll 1
      for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c_begin + c0), r);
#pragma unroll
        for (int j = 0; j < 32; j++) tb[lane * 33 + j] = __uint_as_float(r[j]) * rs;
        __syncwarp();
        const int gn = n0 + c_begin + c0 + lane;
        const float bv = (bias && gn < N) ? bias[gn] : 0.f;
        if (gn < N) {
#pragma unroll 8
          for (int rr = 0; rr < 32; rr++) {
            const int gm = m0 + q * 32 + rr;
            if (gm < M) {
              float o = tb[rr * 33 + lane] + bv;
              if (relu) o = fmaxf(o, 0.f);
              C[(size_t)gm * ldc + gn] = o;        // 32 lanes -> one 128-byte row segment
            }
          }
        }
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

template <int BN, bool PS>
int launch_tc_small(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N,
                    int K, const float* rowscale, const float* bias, int relu, cudaStream_t stream) {
  constexpr size_t stage = 2 * BM * BK * 4 + 2 * BN * BK * 4;
  constexpr size_t smem = (stage > 8 * 32 * 33 * 4 ? stage : 8 * 32 * 33 * 4) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_small_kernel<BN, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    attr_done = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_tf32x3_small_kernel<BN, PS><<<grid, kThreads, smem, stream>>>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale,
                                                                     bias, relu);
  return LCR_OK;
}

template <int BN, int STAGES, bool PS>
int launch_tc(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N, int K,
              const float* rowscale, const float* bias, int relu, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (2 * BM * BK * 4 + 2 * BN * BK * 4) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, STAGES, PS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    attr_done = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_tf32x3_kernel<BN, STAGES, PS><<<grid, kThreads, smem, stream>>>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale,
                                                                       bias, relu);
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
                  int K, const float* rowscale, const float* bias, int relu, cudaStream_t stream) {
  if (K <= 4 * BK) {  // bandwidth-bound unary layers: occupancy-oriented kernel
    if (N <= 32) return launch_tc_small<32, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream);
    if (N <= 64) return launch_tc_small<64, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream);
    if (N <= 128) return launch_tc_small<128, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream);
    return launch_tc_small<256, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream);
  }
  if (N <= 32) return launch_tc<32, 4, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream);
  if (N <= 64) return launch_tc<64, 4, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream);
  // wide tiles amortise the A-operand staging (each 128 x 32 A block is split once per N tile); when
  // 256-wide tiles would leave the last wave of the 148 SMs mostly empty, 128-wide tiles fill it better
  const long tiles_m = (M + 127) / 128;
  const long t256 = tiles_m * ((N + 255) / 256), t128 = tiles_m * ((N + 127) / 128);
  const long w256 = (t256 + LCR_SM_COUNT - 1) / LCR_SM_COUNT, w128 = (t128 + LCR_SM_COUNT - 1) / LCR_SM_COUNT;
  if (N <= 128 || w128 < 2 * w256)   // a 128-wide tile costs about half a 256-wide one
    return launch_tc<128, 3, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream);
  return launch_tc<256, 2, PS>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream);
}

}  // namespace

// Internal entry: tensor-core GEMM.  Requirements: K % 32 == 0, N % 4 == 0, 16-byte aligned rows.
// W_lo != NULL: W / W_lo are the pre-split tf32 hi / lo parts of the weights (lcr_tf32_split).
int lcr_gemm_tf32x3(const float* A, int lda, const float* W, const float* W_lo, int ldw, float* C, int ldc, int M, int N,
                    int K, const float* rowscale, const float* bias, int relu, cudaStream_t stream) {
  LCR_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_tc: bad shape");
  LCR_REQUIRE((K % BK) == 0 && (N % 4) == 0 && (lda % 4) == 0 && (ldw % 4) == 0 && (ldc % 4) == 0,
              "gemm_tc: K must be a multiple of 32; N and leading dimensions multiples of 4");
  LCR_REQUIRE((((uintptr_t)A | (uintptr_t)W | (uintptr_t)W_lo | (uintptr_t)C | (uintptr_t)bias) & 15) == 0,
              "gemm_tc: pointers must be 16-byte aligned");
  if (M == 0) return LCR_OK;
  LcrProfScope prof("gemm_tf32x3", 2.0 * M * N * K, 4.0 * ((double)M * K + (double)K * N + (double)M * N), stream);
  const int rc = W_lo ? gemm_dispatch<true>(A, lda, W, W_lo, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream)
                      : gemm_dispatch<false>(A, lda, W, nullptr, ldw, C, ldc, M, N, K, rowscale, bias, relu, stream);
  if (rc != LCR_OK) return rc;
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
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

extern "C" int lcr_tf32_split(const float* w, int64_t count, float* hi, float* lo, void* stream) {
  LCR_REQUIRE(count >= 0, "tf32_split: bad count");
  if (count == 0) return LCR_OK;
  tf32_split_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, count, hi, lo);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
