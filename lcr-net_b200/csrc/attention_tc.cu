// attention_tc.cu -- a7: multi-head softmax attention of the 3D-RoFormer on the 5th-gen tensor
// cores (tcgen05) with operands fetched by the TMA engine (cp.async.bulk + mbarrier tx counts).
// Reference: dynamic_attention (thdroformer/rpetransformer.py:19-24) and MultiHeadAttention
// (vanilla_transformer.py:58-72): out = softmax(q k^T / sqrt(32)) v per head; rotary embedding is
// applied to q/k beforehand by lcr_rope.
//
// One CTA = (problem, head, 128-query tile); keys stream through in tiles of 64:
//   TMA     : every worker thread issues one 128-byte bulk copy (one q / k / v row slice of the
//             head) into a raw staging buffer; completion is tracked by an mbarrier tx count.
//   workers : 4 warps = 128 threads, thread t owns query row t.  They split the raw fp32 tiles into
//             tf32 hi/lo pairs in the canonical K-major SWIZZLE_128B layout (V is transposed on
//             the fly so that keys become its K dimension), run the online softmax on the score
//             row they read back from TMEM, write P (hi/lo) back INTO TMEM (tcgen05.st) where the
//             second MMA reads it as its A operand, and fold each tile's P.V result into register
//             accumulators.  (P never touches shared memory: 96 KB per CTA, two CTAs per SM.)
//   MMA     : one elected lane issues tcgen05.mma.kind::tf32: S = Q.K^T (128 x 64 x 32) and
//             O_j = P.V (128 x 32 x 64), each as 3 products (hi.hi + hi.lo + lo.hi) -> fp32-class
//             accuracy, which the discrete stages downstream of the transformer (NMS, arg-max
//             correspondences) need; completion via tcgen05.commit -> mbarrier.
// TMEM: S in columns [0, 64), O_j in [64, 96), P_hi in [96, 160), P_lo in [160, 224).
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int QT = 128, KT = 64, HD = 32;
constexpr int kWorkers = 128, kThreadsA = kWorkers + 32;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                   smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_addr(bar);
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
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ float rn_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void ld_tmem32(uint32_t taddr, uint32_t* r) {
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

__device__ __forceinline__ void st_tmem32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// exp(x) on the SFU with a compensated argument (see matching.cu:sk_exp): ~2 ulp
__device__ __forceinline__ float fast_exp(float x) {
  x = fmaxf(x, -100.f);  // masked scores are -inf: exp(-100) flushes to 0 and no inf - inf appears below
  const float kL2E = 1.4426950408889634f, kL2E_lo = 1.925963033500011e-8f;
  const float y = x * kL2E;
  const float e = fmaf(x, kL2E_lo, fmaf(x, kL2E, -y));
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return fmaf(r, e * 0.6931471805599453f, r);
}

// float offset of element (row, col) inside a K-major SWIZZLE_128B tile with 32-float rows
__device__ __forceinline__ int sw_off(int row, int col) { return row * 32 + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3)); }

// Shared-memory map (bytes, all tile bases 1024-aligned)
constexpr int kQhi = 0, kQlo = 16384;                  // 128 x 32
constexpr int kKhi = 32768, kKlo = 40960;              // 64 x 32
constexpr int kVhi = 49152, kVlo = 57344;              // V^T: 2 sub-tiles of 32 (dims) x 32 (keys)
constexpr int kKraw = 65536, kVraw = 73728;            // raw TMA staging, double-buffered: + buf * 16384
constexpr int kQraw = 81920;                           // raw Q (used once) aliases raw buffer 1
constexpr int kSmemBytes = 98304 + 1024;

__global__ void __launch_bounds__(kThreadsA, 2)
attention_tc_kernel(const float* __restrict__ q, int ld_q, const float* __restrict__ k, int ld_k,
                    const float* __restrict__ v, int ld_v, const int64_t* __restrict__ q_off,
                    const int64_t* __restrict__ k_off, int heads, float scale_mul, float* __restrict__ out,
                    int ld_o) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET arithmetic on the shared array (a uintptr_t round trip hides the address space from the
  // compiler: every access through `base` became a generic LD/ST that it also had to order against global stores)
  uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_tq, bar_tkv[2], bar_kv_ready, bar_s_full, bar_p_ready, bar_o_full;
  __shared__ uint32_t tmem_base_s;
  const int prob = blockIdx.y / heads, head = blockIdx.y % heads;
  const int64_t q0 = q_off[prob] + (int64_t)blockIdx.x * QT, q1 = q_off[prob + 1];
  if (q0 >= q1) return;
  const int64_t k0 = k_off[prob], k1 = k_off[prob + 1];
  const int n_tiles = (int)((k1 - k0 + KT - 1) / KT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    bar_init(&bar_tq, 1);
    bar_init(&bar_tkv[0], 1);
    bar_init(&bar_tkv[1], 1);
    bar_init(&bar_kv_ready, kWorkers / 32);
    bar_init(&bar_s_full, 1);
    bar_init(&bar_p_ready, kWorkers / 32);
    bar_init(&bar_o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)),
                 "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 64, tmem_phi = tmem_base + 96, tmem_plo = tmem_base + 160;

  float* Qhi = reinterpret_cast<float*>(base + kQhi);
  float* Qlo = reinterpret_cast<float*>(base + kQlo);
  float* Khi = reinterpret_cast<float*>(base + kKhi);
  float* Klo = reinterpret_cast<float*>(base + kKlo);
  float* Vhi = reinterpret_cast<float*>(base + kVhi);
  float* Vlo = reinterpret_cast<float*>(base + kVlo);
  float* Qraw = reinterpret_cast<float*>(base + kQraw);
  // the TMA engine fetches tile j+1 into the other raw buffer while tile j is being processed
  auto issue_kv = [&](int j) {
    const int64_t kb = k0 + (int64_t)j * KT;
    const int nkv = (int)min((int64_t)KT, k1 - kb);
    const int b = j & 1;
    float* Kr = reinterpret_cast<float*>(base + kKraw + b * 16384);
    float* Vr = reinterpret_cast<float*>(base + kVraw + b * 16384);
    if (tid == 0) bar_expect_tx(&bar_tkv[b], (uint32_t)nkv * HD * 4 * 2);
    if (tid < KT) {
      if (tid < nkv) tma_row(Kr + tid * HD, k + (kb + tid) * ld_k + head * HD, HD * 4, &bar_tkv[b]);
    } else if (tid - KT < nkv) {
      tma_row(Vr + (tid - KT) * HD, v + (kb + tid - KT) * ld_v + head * HD, HD * 4, &bar_tkv[b]);
    }
  };

  if (warp < 4) {
    // ------------------------------------------------------------ Q tile: TMA rows -> split
    const int nq = (int)min((int64_t)QT, q1 - q0);
    if (tid == 0) bar_expect_tx(&bar_tq, (uint32_t)nq * HD * 4);
    if (tid < nq) tma_row(Qraw + tid * HD, q + (q0 + tid) * ld_q + head * HD, HD * 4, &bar_tq);
    issue_kv(0);
    bar_wait(&bar_tq, 0);
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
      float4 x = tid < nq ? *reinterpret_cast<const float4*>(Qraw + tid * HD + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 h = make_float4(rn_tf32(x.x), rn_tf32(x.y), rn_tf32(x.z), rn_tf32(x.w));
      float4 l = make_float4(rn_tf32(x.x - h.x), rn_tf32(x.y - h.y), rn_tf32(x.z - h.z), rn_tf32(x.w - h.w));
      *reinterpret_cast<float4*>(Qhi + sw_off(tid, c)) = h;
      *reinterpret_cast<float4*>(Qlo + sw_off(tid, c)) = l;
    }
    float o_acc[HD];
#pragma unroll
    for (int c = 0; c < HD; c++) o_acc[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n_tiles; j++) {
      const int64_t kb = k0 + (int64_t)j * KT;
      const int nkv = (int)min((int64_t)KT, k1 - kb);
      const uint32_t ph = j & 1;
      // ---------------------------------------------------------- K / V tile (fetched by TMA one tile ahead)
      const float* Kraw = reinterpret_cast<const float*>(base + kKraw + (j & 1) * 16384);
      const float* Vraw = reinterpret_cast<const float*>(base + kVraw + (j & 1) * 16384);
      if (j + 1 < n_tiles) issue_kv(j + 1);
      bar_wait(&bar_tkv[j & 1], (j >> 1) & 1);
      if (tid < KT) {  // K row -> hi / lo (K-major: row = key)
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          float4 x = tid < nkv ? *reinterpret_cast<const float4*>(Kraw + tid * HD + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 h = make_float4(rn_tf32(x.x), rn_tf32(x.y), rn_tf32(x.z), rn_tf32(x.w));
          float4 l = make_float4(rn_tf32(x.x - h.x), rn_tf32(x.y - h.y), rn_tf32(x.z - h.z), rn_tf32(x.w - h.w));
          *reinterpret_cast<float4*>(Khi + sw_off(tid, c)) = h;
          *reinterpret_cast<float4*>(Klo + sw_off(tid, c)) = l;
        }
      } else {         // V row -> transposed: V^T[dim][key], keys are the K dimension of the second MMA
        const int key = tid - KT, sub = key >> 5, kc = key & 31;
#pragma unroll
        for (int d = 0; d < HD; d++) {
          const float x = key < nkv ? Vraw[key * HD + d] : 0.f;
          const float h = rn_tf32(x);
          Vhi[sub * 1024 + sw_off(d, kc)] = h;
          Vlo[sub * 1024 + sw_off(d, kc)] = rn_tf32(x - h);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(&bar_kv_ready);
      // ---------------------------------------------------------- softmax on the score row
      bar_wait(&bar_s_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t sr[KT];
      ld_tmem32(tmem_s + ((uint32_t)(warp * 32) << 16), sr);
      ld_tmem32(tmem_s + ((uint32_t)(warp * 32) << 16) + 32, sr + 32);
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < KT; c++) {
        const float s = c < nkv ? __uint_as_float(sr[c]) * scale_mul : -INFINITY;
        sr[c] = __float_as_uint(s);
        mx = fmaxf(mx, s);
      }
      const float m_new = fmaxf(m_run, mx);
      const float corr = fast_exp(m_run - m_new);
      float psum = 0.f;
      const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
      for (int c0 = 0; c0 < KT; c0 += 32) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int e = 0; e < 32; e++) {
          const float p = fast_exp(__uint_as_float(sr[c0 + e]) - m_new);
          psum += p;
          const float h = rn_tf32(p);
          hi[e] = __float_as_uint(h);
          lo[e] = __float_as_uint(rn_tf32(p - h));
        }
        st_tmem32(tmem_phi + lane_base + c0, hi);
        st_tmem32(tmem_plo + lane_base + c0, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      l_run = l_run * corr + psum;
      m_run = m_new;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(&bar_p_ready);
      // ---------------------------------------------------------- fold this tile's P.V
      bar_wait(&bar_o_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t orr[HD];
      ld_tmem32(tmem_o + ((uint32_t)(warp * 32) << 16), orr);
#pragma unroll
      for (int c = 0; c < HD; c++) o_acc[c] = fmaf(o_acc[c], corr, __uint_as_float(orr[c]));
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (tid < nq) {
      float* o = out + (q0 + tid) * ld_o + head * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 4)
        *reinterpret_cast<float4*>(o + c) =
            make_float4(o_acc[c] / l_run, o_acc[c + 1] / l_run, o_acc[c + 2] / l_run, o_acc[c + 3] / l_run);
    }
  } else if (lane == 0) {
    // -------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
    constexpr uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
    const uint32_t qhi = smem_addr(base + kQhi), qlo = smem_addr(base + kQlo), khi = smem_addr(base + kKhi),
                   klo = smem_addr(base + kKlo), vhi = smem_addr(base + kVhi), vlo = smem_addr(base + kVlo);
    for (int j = 0; j < n_tiles; j++) {
      const uint32_t ph = j & 1;
      bar_wait(&bar_kv_ready, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < HD / 8; ks++) {   // S = Q . K^T
        const uint32_t koff = ks * 32;
        umma_tf32(tmem_s, sw128_desc(qhi + koff), sw128_desc(khi + koff), idesc_s, ks != 0);
        umma_tf32(tmem_s, sw128_desc(qhi + koff), sw128_desc(klo + koff), idesc_s, 1);
        umma_tf32(tmem_s, sw128_desc(qlo + koff), sw128_desc(khi + koff), idesc_s, 1);
      }
      umma_commit(&bar_s_full);
      bar_wait(&bar_p_ready, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < KT / 8; ks++) {   // O_j = P . V  (K = keys: 2 sub-tiles of 32)
        const uint32_t sub = ks >> 2, koff = (ks & 3) * 32;
        const uint32_t vb = sub * 4096 + koff, pc = ks * 8;   // A (P) from TMEM: 8 tf32 columns per k-step
        umma_tf32_ts(tmem_o, tmem_phi + pc, sw128_desc(vhi + vb), idesc_o, ks != 0);
        umma_tf32_ts(tmem_o, tmem_phi + pc, sw128_desc(vlo + vb), idesc_o, 1);
        umma_tf32_ts(tmem_o, tmem_plo + pc, sw128_desc(vhi + vb), idesc_o, 1);
      }
      umma_commit(&bar_o_full);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}


// ================================================================== tensor-map variant (default)
// Same pipeline, but every tile is ONE TMA tensor copy (cp.async.bulk.tensor.2d through a CUtensorMap over the
// [rows, heads * 32] operand with the 128-byte swizzle) instead of 64 / 128 per-row bulk copies issued lane by lane:
// the tile lands in the canonical K-major SWIZZLE_128B layout, so the raw fp32 Q and K tiles ARE the `hi` operands
// (kind::tf32 reads the upper 19 bits: the tensor core truncates) and the workers only write lo = x - trunc(x);
// V is still transposed by the workers (keys become the K dimension of the second MMA), now from conflict-free
// 16-byte reads of the swizzled tile (the per-row staging buffer had all 32 lanes on one bank).  Rows past the end of
// a problem belong to the next problem (or are zero-filled past the tensor): their scores are masked to -inf before
// the softmax and V rows past the end are zeroed, Q rows past the end are never stored.
constexpr int k2Qhi = 0, k2Qlo = 16384;                    // 128 x 32 (hi = raw TMA tile)
constexpr int k2Kraw = 32768;                              // 64 x 32, two stages (+ 8192): the hi operand
constexpr int k2Klo = 49152;
constexpr int k2Vraw = 57344;                              // 64 x 32, two stages (+ 8192)
constexpr int k2Vhi = 73728, k2Vlo = 81920;                // V^T: 2 sub-tiles of 32 (dims) x 32 (keys)
constexpr int k2SmemBytes = 90112 + 1024;

__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar))
      : "memory");
}
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void __launch_bounds__(kThreadsA, 2)
attention_tma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const int64_t* __restrict__ q_off,
                     const int64_t* __restrict__ k_off, int heads, float scale_mul, float* __restrict__ out,
                     int ld_o) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_tq, bar_tkv[2], bar_kv_ready, bar_s_full, bar_p_ready, bar_o_full;
  __shared__ uint32_t tmem_base_s;
  const int prob = blockIdx.y / heads, head = blockIdx.y % heads;
  const int64_t q0 = q_off[prob] + (int64_t)blockIdx.x * QT, q1 = q_off[prob + 1];
  if (q0 >= q1) return;
  const int64_t k0 = k_off[prob], k1 = k_off[prob + 1];
  const int n_tiles = (int)((k1 - k0 + KT - 1) / KT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    bar_init(&bar_tq, 1);
    bar_init(&bar_tkv[0], 1);
    bar_init(&bar_tkv[1], 1);
    bar_init(&bar_kv_ready, kWorkers / 32);
    bar_init(&bar_s_full, 1);
    bar_init(&bar_p_ready, kWorkers / 32);
    bar_init(&bar_o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)),
                 "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 64, tmem_phi = tmem_base + 96, tmem_plo = tmem_base + 160;
  const float* Qhi = reinterpret_cast<const float*>(base + k2Qhi);
  float* Qlo = reinterpret_cast<float*>(base + k2Qlo);
  float* Klo = reinterpret_cast<float*>(base + k2Klo);
  float* Vhi = reinterpret_cast<float*>(base + k2Vhi);
  float* Vlo = reinterpret_cast<float*>(base + k2Vlo);
  // one thread arms the stage's transaction barrier and issues the two tile copies of key tile j
  auto issue_kv = [&](int j) {
    const int b = j & 1;
    bar_expect_tx(&bar_tkv[b], 2u * KT * HD * 4);
    const int row = (int)(k0 + (int64_t)j * KT);
    tma_tile_2d(smem_addr(base + k2Kraw + b * 8192), &map_k, head * HD, row, &bar_tkv[b]);
    tma_tile_2d(smem_addr(base + k2Vraw + b * 8192), &map_v, head * HD, row, &bar_tkv[b]);
  };

  if (warp < 4) {
    const int nq = (int)min((int64_t)QT, q1 - q0);
    if (tid == 0) {
      bar_expect_tx(&bar_tq, (uint32_t)QT * HD * 4);
      tma_tile_2d(smem_addr(base + k2Qhi), &map_q, head * HD, (int)q0, &bar_tq);
      issue_kv(0);
    }
    bar_wait(&bar_tq, 0);
#pragma unroll
    for (int c = 0; c < HD; c += 4) {        // Q lo (row = query)
      const float4 x = *reinterpret_cast<const float4*>(Qhi + sw_off(tid, c));
      *reinterpret_cast<float4*>(Qlo + sw_off(tid, c)) =
          make_float4(x.x - trunc_tf32(x.x), x.y - trunc_tf32(x.y), x.z - trunc_tf32(x.z), x.w - trunc_tf32(x.w));
    }
    float o_acc[HD];
#pragma unroll
    for (int c = 0; c < HD; c++) o_acc[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = 0; j < n_tiles; j++) {
      const int64_t kb = k0 + (int64_t)j * KT;
      const int nkv = (int)min((int64_t)KT, k1 - kb);
      const uint32_t ph = j & 1;
      const float* Kraw = reinterpret_cast<const float*>(base + k2Kraw + (j & 1) * 8192);
      const float* Vraw = reinterpret_cast<const float*>(base + k2Vraw + (j & 1) * 8192);
      if (tid == 0 && j + 1 < n_tiles) issue_kv(j + 1);
      bar_wait(&bar_tkv[j & 1], (j >> 1) & 1);
      if (tid < KT) {  // K lo (row = key)
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          const float4 x = *reinterpret_cast<const float4*>(Kraw + sw_off(tid, c));
          *reinterpret_cast<float4*>(Klo + sw_off(tid, c)) =
              make_float4(x.x - trunc_tf32(x.x), x.y - trunc_tf32(x.y), x.z - trunc_tf32(x.z), x.w - trunc_tf32(x.w));
        }
      } else {         // V row -> transposed: V^T[dim][key], keys are the K dimension of the second MMA
        const int key = tid - KT, sub = key >> 5, kc = key & 31;
        const bool live = key < nkv;
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          const float4 x4 = *reinterpret_cast<const float4*>(Vraw + sw_off(key, c));
          const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float x = live ? xs[e] : 0.f;
            const float h = trunc_tf32(x);
            Vhi[sub * 1024 + sw_off(c + e, kc)] = h;
            Vlo[sub * 1024 + sw_off(c + e, kc)] = x - h;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(&bar_kv_ready);
      // ---------------------------------------------------------- softmax on the score row
      bar_wait(&bar_s_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t sr[KT];
      ld_tmem32(tmem_s + ((uint32_t)(warp * 32) << 16), sr);
      ld_tmem32(tmem_s + ((uint32_t)(warp * 32) << 16) + 32, sr + 32);
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < KT; c++) {
        const float s_ = c < nkv ? __uint_as_float(sr[c]) * scale_mul : -INFINITY;
        sr[c] = __float_as_uint(s_);
        mx = fmaxf(mx, s_);
      }
      const float m_new = fmaxf(m_run, mx);
      const float corr = fast_exp(m_run - m_new);
      float psum = 0.f;
      const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
      for (int c0 = 0; c0 < KT; c0 += 32) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int e = 0; e < 32; e++) {
          const float p = fast_exp(__uint_as_float(sr[c0 + e]) - m_new);
          psum += p;
          const float h = trunc_tf32(p);       // the tensor core truncates its operands: hi + lo = p exactly
          hi[e] = __float_as_uint(h);
          lo[e] = __float_as_uint(p - h);
        }
        st_tmem32(tmem_phi + lane_base + c0, hi);
        st_tmem32(tmem_plo + lane_base + c0, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      l_run = l_run * corr + psum;
      m_run = m_new;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) bar_arrive(&bar_p_ready);
      // ---------------------------------------------------------- fold this tile's P.V
      bar_wait(&bar_o_full, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t orr[HD];
      ld_tmem32(tmem_o + ((uint32_t)(warp * 32) << 16), orr);
#pragma unroll
      for (int c = 0; c < HD; c++) o_acc[c] = fmaf(o_acc[c], corr, __uint_as_float(orr[c]));
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (tid < nq) {
      float* o = out + (q0 + tid) * ld_o + head * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 4)
        *reinterpret_cast<float4*>(o + c) =
            make_float4(o_acc[c] / l_run, o_acc[c + 1] / l_run, o_acc[c + 2] / l_run, o_acc[c + 3] / l_run);
    }
  } else if (lane == 0) {
    // -------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
    constexpr uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
    const uint32_t qhi = smem_addr(base + k2Qhi), qlo = smem_addr(base + k2Qlo), klo = smem_addr(base + k2Klo),
                   vhi = smem_addr(base + k2Vhi), vlo = smem_addr(base + k2Vlo);
    for (int j = 0; j < n_tiles; j++) {
      const uint32_t ph = j & 1;
      const uint32_t khi = smem_addr(base + k2Kraw + (j & 1) * 8192);
      bar_wait(&bar_kv_ready, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < HD / 8; ks++) {   // S = Q . K^T
        const uint32_t koff = ks * 32;
        umma_tf32(tmem_s, sw128_desc(qhi + koff), sw128_desc(khi + koff), idesc_s, ks != 0);
        umma_tf32(tmem_s, sw128_desc(qhi + koff), sw128_desc(klo + koff), idesc_s, 1);
        umma_tf32(tmem_s, sw128_desc(qlo + koff), sw128_desc(khi + koff), idesc_s, 1);
      }
      umma_commit(&bar_s_full);
      bar_wait(&bar_p_ready, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < KT / 8; ks++) {   // O_j = P . V  (K = keys: 2 sub-tiles of 32)
        const uint32_t sub = ks >> 2, koff = (ks & 3) * 32;
        const uint32_t vb = sub * 4096 + koff, pc = ks * 8;
        umma_tf32_ts(tmem_o, tmem_phi + pc, sw128_desc(vhi + vb), idesc_o, ks != 0);
        umma_tf32_ts(tmem_o, tmem_phi + pc, sw128_desc(vlo + vb), idesc_o, 1);
        umma_tf32_ts(tmem_o, tmem_plo + pc, sw128_desc(vhi + vb), idesc_o, 1);
      }
      umma_commit(&bar_o_full);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

typedef CUresult (*EncodeTiledFnA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
EncodeTiledFnA attn_map_encoder() {
  static EncodeTiledFnA fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFnA)p;
  }
  return fn;
}
// [rows, cols] fp32 operand with row stride ld (floats): boxes of one head (32 floats = one swizzle row) x box_rows
bool attn_make_map(CUtensorMap* m, const float* ptr, int64_t rows, int cols, int ld, int box_rows) {
  EncodeTiledFnA enc = attn_map_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)HD, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

}  // namespace

extern "C" int lcr_attention_tc(const float* q, int ld_q, const float* k, int ld_k, const float* v, int ld_v,
                                const int64_t* q_off, const int64_t* k_off, int n_problems, int64_t max_q_rows,
                                int heads, int head_dim, float* out, int ld_out, double flops_hint, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(head_dim == HD, "attention_tc: head_dim must be 32");
  LCR_REQUIRE(n_problems >= 1 && heads >= 1 && max_q_rows >= 0, "attention_tc: bad sizes");
  LCR_REQUIRE((ld_q % 4) == 0 && (ld_k % 4) == 0 && (ld_v % 4) == 0 && (ld_out % 4) == 0 &&
                  (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0,
              "attention_tc: rows must be 16-byte aligned (TMA bulk copies)");
  if (max_q_rows == 0) return LCR_OK;
  static LcrOncePerDevice attr_done;
  const int attr_done_dev = attr_done.need();
  if (attr_done_dev != -1) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done.done(attr_done_dev);
  }
  LcrProfScope prof("attention_tc", flops_hint, 0.0, stream);
  dim3 grid((unsigned)((max_q_rows + QT - 1) / QT), (unsigned)(n_problems * heads));
  attention_tc_kernel<<<grid, kThreadsA, kSmemBytes, stream>>>(q, ld_q, k, ld_k, v, ld_v, q_off, k_off, heads,
                                                               1.0f / sqrtf((float)head_dim), out, ld_out);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}

// Tensor-map variant of lcr_attention_tc (same arguments + the row counts of the q and k / v operands, which bound
// the tensor maps).  Default attention of the registration path.
extern "C" int lcr_attention_tma(const float* q, int ld_q, int64_t q_rows, const float* k, int ld_k, const float* v,
                                 int ld_v, int64_t k_rows, const int64_t* q_off, const int64_t* k_off, int n_problems,
                                 int64_t max_q_rows, int heads, int head_dim, float* out, int ld_out,
                                 double flops_hint, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(head_dim == HD, "attention_tma: head_dim must be 32");
  LCR_REQUIRE(n_problems >= 1 && heads >= 1 && max_q_rows >= 0 && q_rows >= 0 && k_rows >= 0, "attention_tma: bad sizes");
  LCR_REQUIRE((ld_q % 4) == 0 && (ld_k % 4) == 0 && (ld_v % 4) == 0 && (ld_out % 4) == 0 &&
                  (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0,
              "attention_tma: rows must be 16-byte aligned (TMA)");
  LCR_REQUIRE(q_rows < (1ll << 31) && k_rows < (1ll << 31), "attention_tma: more than 2^31 rows");
  if (max_q_rows == 0 || q_rows == 0) return LCR_OK;
  LCR_REQUIRE(k_rows >= 1, "attention_tma: empty key operand");
  CUtensorMap mq, mk, mv;
  if (!(attn_make_map(&mq, q, q_rows, heads * HD, ld_q, QT) && attn_make_map(&mk, k, k_rows, heads * HD, ld_k, KT) &&
        attn_make_map(&mv, v, k_rows, heads * HD, ld_v, KT)))
    // the driver refused a tensor map (e.g. no cuTensorMapEncodeTiled entry point): per-row bulk-copy variant
    return lcr_attention_tc(q, ld_q, k, ld_k, v, ld_v, q_off, k_off, n_problems, max_q_rows, heads, head_dim, out,
                            ld_out, flops_hint, stream_);
  static LcrOncePerDevice attr_done;
  const int attr_done_dev = attr_done.need();
  if (attr_done_dev != -1) {
    LCR_CUDA_TRY(cudaFuncSetAttribute(attention_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes));
    attr_done.done(attr_done_dev);
  }
  LcrProfScope prof("attention_tc", flops_hint, 0.0, stream);
  dim3 grid((unsigned)((max_q_rows + QT - 1) / QT), (unsigned)(n_problems * heads));
  attention_tma_kernel<<<grid, kThreadsA, k2SmemBytes, stream>>>(mq, mk, mv, q_off, k_off, heads,
                                                                 1.0f / sqrtf((float)head_dim), out, ld_out);
  LCR_LAUNCHED(1);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
