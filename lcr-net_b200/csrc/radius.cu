// radius.cu -- a2: stack-mode radius neighbour search on sm_100a.
//
// Reference semantics (utils/extensions/cpu/radius_neighbors/radius_neighbors_cpu.cpp:3-91,
// extra/nanoflann/nanoflann.hpp:208-256, 423-442): per query, every support of the SAME batch
// element with d2 < r2 (strict), d2 = ((dx*dx + dy*dy) + dz*dz) in fp32 without FMA,
// r2 = radius*radius in fp32; ascending d2; global (stack) indices; pad = total support count.
// Exact-distance ties are returned in ascending support index (the deterministic order of
// cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:125-208); the reference's own tie order is
// std::sort-unstable.
//
// Design (irregular gather, HBM/L2-bound; no tensor cores):
//   build : per-cloud uniform grid with cell edge 1.001*r, stored as an open-addressing hash of
//           (cloud, cx, cy, cz) -> [start, count) into a cell-ordered float4 copy of the
//           supports (xyz + original index) -- counting sort: count, device scan, scatter.
//   query : one warp per query.  27 lanes resolve the 27 stencil cells in parallel, the warp
//           then streams each cell's float4 run with coalesced 512 B loads, tests d2 < r2 and
//           compacts hits with __ballot_sync/__popc into a shared-memory key buffer
//           (key = d2 bits << 32 | support index: d2 >= 0 so integer order == float order and
//           the low word is the tie-break).  A warp bitonic sort orders the keys; the first
//           `width` are written, the rest of the row padded.
//   spill : queries with more hits than the warp buffer are queued and handled by a second
//           kernel, one CTA per query with an 8192-key buffer.
// No host synchronisation anywhere.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr uint64_t kEmpty = 0xFFFFFFFFFFFFFFFFull;
constexpr int kCellBits = 14;  // key field per axis
// Supports may span at most 2048 cells per axis: beyond that the fp32 rounding of (x - o) * inv_cell can exceed the
// 0.1 % cell margin (cell edge 1.001 r) that keeps every neighbour inside the 27-cell stencil.  KITTI-scale clouds
// span ~130 cells at level 0.  (Query cells are clamped to [-1, kCellMax + 1] and keyed with a +1 shift.)
constexpr int kCellMax = 2047;
constexpr int kWarpCap = 512;     // keys per warp buffer
constexpr int kWarpsPerCta = 8;   // query kernel: 256 threads
constexpr int kSpillCap = LCR_RADIUS_MAX_WIDTH;  // keys per CTA in the spill kernel
constexpr int kSpillThreads = 512;

struct GridGeom {
  float ox, oy, oz;
};

__device__ __forceinline__ uint64_t cell_key(int b, int cx, int cy, int cz) {
  return ((uint64_t)b << (3 * kCellBits)) | ((uint64_t)cz << (2 * kCellBits)) | ((uint64_t)cy << kCellBits) |
         (uint64_t)cx;
}

__device__ __forceinline__ int cell_coord(float x, float o, float inv_cell) {
  return (int)floorf((x - o) * inv_cell);
}

__global__ void grid_geom_kernel(const unsigned* __restrict__ bbox, int batch, GridGeom* __restrict__ geom) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  GridGeom g;
  g.ox = lcr_ord2f(bbox[6 * b]);
  g.oy = lcr_ord2f(bbox[6 * b + 1]);
  g.oz = lcr_ord2f(bbox[6 * b + 2]);
  geom[b] = g;
}

__global__ void cell_insert_kernel(const float* __restrict__ s, int64_t ns, const int64_t* __restrict__ s_off,
                                   int batch, const GridGeom* __restrict__ geom, float inv_cell,
                                   unsigned long long* __restrict__ tkeys, uint32_t* __restrict__ tcount,
                                   uint64_t tmask, uint32_t* __restrict__ slot_of, int* __restrict__ err) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const int b = lcr_find_segment(s_off, batch, j);
  const GridGeom g = geom[b];
  const int cx = cell_coord(s[3 * j], g.ox, inv_cell);
  const int cy = cell_coord(s[3 * j + 1], g.oy, inv_cell);
  const int cz = cell_coord(s[3 * j + 2], g.oz, inv_cell);
  if (cx < 0 || cy < 0 || cz < 0 || cx > kCellMax || cy > kCellMax || cz > kCellMax) {
    *err = LCR_ERR_OVERFLOW;  // cloud extent exceeds 2048 cells per axis (or NaN coordinates)
    slot_of[j] = 0xFFFFFFFFu;
    return;
  }
  const uint64_t key = cell_key(b, cx, cy, cz);
  uint64_t h = lcr_mix64(key) & tmask;
  while (true) {
    unsigned long long old = atomicCAS(&tkeys[h], (unsigned long long)kEmpty, (unsigned long long)key);
    if (old == kEmpty || old == key) break;
    h = (h + 1) & tmask;
  }
  slot_of[j] = (uint32_t)h;
  atomicAdd(&tcount[h], 1u);
}

__global__ void cell_scatter_kernel(const float* __restrict__ s, int64_t ns, const uint32_t* __restrict__ slot_of,
                                    const uint32_t* __restrict__ tstart, uint32_t* __restrict__ tcursor,
                                    float4* __restrict__ sorted) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const uint32_t h = slot_of[j];
  if (h == 0xFFFFFFFFu) return;
  const uint32_t pos = tstart[h] + atomicAdd(&tcursor[h], 1u);
  sorted[pos] = make_float4(s[3 * j], s[3 * j + 1], s[3 * j + 2], __uint_as_float((uint32_t)j));
}

__device__ __forceinline__ void lookup_cell(uint64_t key, const unsigned long long* __restrict__ tkeys,
                                            const uint32_t* __restrict__ tstart,
                                            const uint32_t* __restrict__ tcount, uint64_t tmask, uint32_t& start,
                                            uint32_t& count) {
  uint64_t h = lcr_mix64(key) & tmask;
  while (true) {
    const unsigned long long k = tkeys[h];
    if (k == key) {
      start = tstart[h];
      count = tcount[h];
      return;
    }
    if (k == kEmpty) {
      start = 0;
      count = 0;
      return;
    }
    h = (h + 1) & tmask;
  }
}

// d2 exactly as nanoflann's L2_Simple_Adaptor accumulates it: result starts at 0 and adds
// diff*diff per dimension, each operation individually rounded (no contraction).
__device__ __forceinline__ float ref_d2(float qx, float qy, float qz, float sx, float sy, float sz) {
  const float dx = __fsub_rn(qx, sx), dy = __fsub_rn(qy, sy), dz = __fsub_rn(qz, sz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// In-place ascending bitonic sort of n (power of two) keys by `nthreads` cooperating threads.
template <bool kBlock>
__device__ __forceinline__ void bitonic_sort(unsigned long long* keys, int n, int tid, int nthreads) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n; i += nthreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      if (kBlock) __syncthreads(); else __syncwarp();
    }
  }
}

// Ascending bitonic sort of 32 * R keys held in registers, key e = r * 32 + lane in v[r]: exchanges at
// distance < 32 go through warp shuffles, larger distances are register-to-register with compile-time
// directions.  ~2.4x fewer instructions than the shared-memory network above for the common sizes (64 and
// 128 keys: half of the lanes idle in every shared-memory pass, and every pass pays a __syncwarp).
template <int R>
__device__ __forceinline__ void warp_bitonic_regs(unsigned long long (&v)[R], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < R; r++) {
          if ((r & jr) == 0) {
            const bool up = ((r * 32) & k) == 0;
            const unsigned long long a = v[r], b = v[r | jr];
            const bool sw = (a > b) == up;
            v[r] = sw ? b : a;
            v[r | jr] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool lower = (lane & j) == 0;
          const bool up = (((r * 32) | lane) & k) == 0;
          const unsigned long long lo = v[r] < o ? v[r] : o, hi = v[r] < o ? o : v[r];
          v[r] = (lower == up) ? lo : hi;
        }
      }
    }
  }
}

template <int R, typename IdxT>
__device__ __forceinline__ void sort_and_write(const unsigned long long* keys, int total, int lane, IdxT* row, int width,
                                               int64_t ns_total) {
  unsigned long long v[R];
#pragma unroll
  for (int r = 0; r < R; r++) v[r] = (r * 32 + lane < total) ? keys[r * 32 + lane] : kEmpty;
  warp_bitonic_regs<R>(v, lane);
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int t = r * 32 + lane;
    if (t < width) row[t] = t < total ? (IdxT)(uint32_t)(v[r] & 0xFFFFFFFFull) : (IdxT)ns_total;
  }
  for (int t = R * 32 + lane; t < width; t += 32) row[t] = (IdxT)ns_total;
}

// Tail of a query: count / running maximum, then sort the hits and write the first `width` (or queue the query for
// the spill kernel when it has more hits than the warp buffer).
template <typename IdxT>
__device__ __forceinline__ void finish_query(unsigned long long* keys, int total, int lane, int64_t qi, int width,
                                             int64_t ns_total, IdxT* __restrict__ out_idx,
                                             int32_t* __restrict__ out_counts, int32_t* __restrict__ out_max,
                                             uint32_t* __restrict__ spill_list, uint32_t* __restrict__ spill_n,
                                             int cap = kWarpCap) {
  if (lane == 0) {
    if (out_counts) out_counts[qi] = total;
    // widest row so far: a (possibly stale) plain read filters almost every warp, so the warps of a CTA need no
    // block-level reduction and retire independently (ncu: 14 % of the warp samples sat in that final barrier)
    if (total > *reinterpret_cast<volatile int32_t*>(out_max)) atomicMax(out_max, total);
  }
  if (out_idx == nullptr) return;
  if (total <= cap) {
    IdxT* row = out_idx + (size_t)qi * width;
    __syncwarp();
    if (total <= 32) {
      sort_and_write<1>(keys, total, lane, row, width, ns_total);
    } else if (total <= 64) {
      sort_and_write<2>(keys, total, lane, row, width, ns_total);
    } else if (total <= 128) {
      sort_and_write<4>(keys, total, lane, row, width, ns_total);
    } else if (total <= 256) {
      sort_and_write<8>(keys, total, lane, row, width, ns_total);
    } else {
      int n = 512;
      for (int i = total + lane; i < n; i += 32) keys[i] = kEmpty;
      __syncwarp();
      bitonic_sort<false>(keys, n, lane, 32);
      for (int t = lane; t < width; t += 32)
        row[t] = t < total ? (IdxT)(uint32_t)(keys[t] & 0xFFFFFFFFull) : (IdxT)ns_total;
    }
  } else if (lane == 0) {
    spill_list[atomicAdd(spill_n, 1u)] = (uint32_t)qi;
  }
}

// NEAREST: only the closest support inside the radius is wanted (the up-sampling tables: the decoder reads column 0
// only, backbone4.py:333-373) -- every lane keeps the minimum (d2, index) key of its candidates, one warp reduction,
// no key buffer and no sort.
template <typename IdxT, bool NEAREST = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
query_kernel(const float* __restrict__ q, int64_t nq, const int64_t* __restrict__ q_off, int batch,
             const GridGeom* __restrict__ geom, float inv_cell, float r2,
             const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tstart,
             const uint32_t* __restrict__ tcount, uint64_t tmask, const float4* __restrict__ sorted, int width,
             int64_t ns_total, IdxT* __restrict__ out_idx, int32_t* __restrict__ out_counts,
             int32_t* __restrict__ out_max, uint32_t* __restrict__ spill_list, uint32_t* __restrict__ spill_n) {
  __shared__ unsigned long long s_keys[kWarpsPerCta][kWarpCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t qi = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  int total = 0;
  if (qi < nq) {
    const int b = lcr_find_segment(q_off, batch, qi);
    const GridGeom g = geom[b];
    const float qx = q[3 * qi], qy = q[3 * qi + 1], qz = q[3 * qi + 2];
    const int cx = cell_coord(qx, g.ox, inv_cell), cy = cell_coord(qy, g.oy, inv_cell),
              cz = cell_coord(qz, g.oz, inv_cell);
    uint32_t c_start = 0, c_count = 0;
    if (lane < 27) {
      const int nx = cx + (lane % 3) - 1, ny = cy + ((lane / 3) % 3) - 1, nz = cz + (lane / 9) - 1;
      if (nx >= 0 && ny >= 0 && nz >= 0 && nx <= kCellMax && ny <= kCellMax && nz <= kCellMax)
        lookup_cell(cell_key(b, nx, ny, nz), tkeys, tstart, tcount, tmask, c_start, c_count);
    }
    unsigned long long* keys = s_keys[warp];
    const bool want_idx = out_idx != nullptr;
    unsigned long long best = kEmpty;
    // The 27 candidate runs are walked as ONE flat stream of T candidates, 32 per step (a cell holds ~18 points:
    // a per-cell loop left 44 % of the lanes idle and paid 27 loop iterations): an inclusive warp scan of the run
    // lengths, then every lane finds the run of its candidate by a 5-step binary search over the lanes' exclusive
    // prefixes (shuffles).  The hits of a query are sorted afterwards, so the visiting order does not matter.
    uint32_t incl = c_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t excl = incl - c_count;
    const uint32_t n_cand = __shfl_sync(0xffffffffu, incl, 31);
    for (uint32_t base = 0; base < n_cand; base += 32) {
      const uint32_t c = base + lane;
      int run = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const uint32_t e = __shfl_sync(0xffffffffu, excl, run + step);
        if (e <= c) run += step;
      }
      const uint32_t st = __shfl_sync(0xffffffffu, c_start, run), ex = __shfl_sync(0xffffffffu, excl, run);
      bool hit = false;
      unsigned long long key = 0;
      if (c < n_cand) {
        const float4 p = sorted[st + (c - ex)];
        const float d2 = ref_d2(qx, qy, qz, p.x, p.y, p.z);
        hit = d2 < r2;
        key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(p.w);
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (NEAREST) {
        if (hit && key < best) best = key;
      } else if (hit && want_idx) {
        const int pos = total + __popc(m & ((1u << lane) - 1u));
        if (pos < kWarpCap) keys[pos] = key;
      }
      total += __popc(m);
    }
    if (NEAREST) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other < best) best = other;
      }
      if (lane == 0) {
        if (out_counts) out_counts[qi] = total;
        if (total > *reinterpret_cast<volatile int32_t*>(out_max)) atomicMax(out_max, total);
        if (out_idx) out_idx[(size_t)qi * width] = total > 0 ? (IdxT)(uint32_t)(best & 0xFFFFFFFFull) : (IdxT)ns_total;
      }
    } else {
      finish_query<IdxT>(keys, total, lane, qi, width, ns_total, out_idx, out_counts, out_max, spill_list, spill_n);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Self tables (queries == supports: four of the seven tables of a pyramid, 73 % of the query time): one warp per
// occupied CELL instead of per query.  The ~8 queries of a cell share their 27-cell stencil, so the hash lookups,
// the run prefix sums and the candidate loads happen once per cell: the candidates (~230 float4) are staged in
// shared memory and every query of the cell walks them with one LDS.128 per lane and step -- the per-query kernel
// spent half of its instructions on the 5-step binary search that maps a flat candidate index to its run, for
// every query anew.  Cells are handed out through an atomic counter (costs differ by 100x between a sparse cell
// and a dense one near the sensor); a cell with more candidates than the stage falls back to the flat stream.
constexpr int kCellWarps = 4;        // warps per CTA
constexpr int kCellCand = 512;       // staged candidates per warp (8 KB)
constexpr int kCellKeys = 256;       // hit keys per warp (2 KB); more hits -> spill kernel

__global__ void compact_cells_kernel(const unsigned long long* __restrict__ tkeys, uint32_t n_slots,
                                     uint32_t* __restrict__ occ_list, uint32_t* __restrict__ n_occ) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool occ = s < n_slots && tkeys[s] != kEmpty;
  const unsigned m = __ballot_sync(0xffffffffu, occ);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(n_occ, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (occ) occ_list[base + __popc(m & ((1u << lane) - 1u))] = s;
}

template <typename IdxT>
__global__ void __launch_bounds__(kCellWarps * 32)
query_self_kernel(const uint32_t* __restrict__ occ_list, const uint32_t* __restrict__ n_occ,
                  uint32_t* __restrict__ next_cell, float r2, const unsigned long long* __restrict__ tkeys,
                  const uint32_t* __restrict__ tstart, const uint32_t* __restrict__ tcount, uint64_t tmask,
                  const float4* __restrict__ sorted, int width, int64_t ns_total, IdxT* __restrict__ out_idx,
                  int32_t* __restrict__ out_counts, int32_t* __restrict__ out_max,
                  uint32_t* __restrict__ spill_list, uint32_t* __restrict__ spill_n) {
  __shared__ __align__(16) float4 s_cand[kCellWarps][kCellCand];
  __shared__ unsigned long long s_keys[kCellWarps][kCellKeys];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* cand = s_cand[warp];
  unsigned long long* keys = s_keys[warp];
  const uint32_t n_cells = *n_occ;
  const bool want_idx = out_idx != nullptr;
  while (true) {
    uint32_t ci = 0;
    if (lane == 0) ci = atomicAdd(next_cell, 1u);
    ci = __shfl_sync(0xffffffffu, ci, 0);
    if (ci >= n_cells) break;
    const uint32_t slot = occ_list[ci];
    const unsigned long long key = tkeys[slot];
    const int b = (int)(key >> (3 * kCellBits));
    const int cz = (int)((key >> (2 * kCellBits)) & ((1u << kCellBits) - 1u));
    const int cy = (int)((key >> kCellBits) & ((1u << kCellBits) - 1u));
    const int cx = (int)(key & ((1u << kCellBits) - 1u));
    const uint32_t q_start = tstart[slot], q_n = tcount[slot];
    uint32_t c_start = 0, c_count = 0;
    if (lane < 27) {
      const int nx = cx + (lane % 3) - 1, ny = cy + ((lane / 3) % 3) - 1, nz = cz + (lane / 9) - 1;
      if (nx >= 0 && ny >= 0 && nz >= 0 && nx <= kCellMax && ny <= kCellMax && nz <= kCellMax)
        lookup_cell(cell_key(b, nx, ny, nz), tkeys, tstart, tcount, tmask, c_start, c_count);
    }
    uint32_t incl = c_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t excl = incl - c_count;
    const uint32_t n_cand = __shfl_sync(0xffffffffu, incl, 31);
    const bool staged = n_cand <= (uint32_t)kCellCand;
    // candidate c of the flat stream -> its float4 in the cell-ordered support copy
    auto fetch = [&](uint32_t c) -> float4 {
      int run = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const uint32_t e = __shfl_sync(0xffffffffu, excl, run + step);
        if (e <= c) run += step;
      }
      const uint32_t st = __shfl_sync(0xffffffffu, c_start, run), ex = __shfl_sync(0xffffffffu, excl, run);
      return c < n_cand ? sorted[st + (c - ex)] : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    __syncwarp();
    if (staged)
      for (uint32_t base = 0; base < n_cand; base += 32) {
        const float4 p = fetch(base + lane);
        if (base + lane < n_cand) cand[base + lane] = p;
      }
    __syncwarp();
    for (uint32_t qk = 0; qk < q_n; qk++) {
      const float4 qp = sorted[q_start + qk];
      const int64_t qi = (int64_t)__float_as_uint(qp.w);
      int total = 0;
      for (uint32_t base = 0; base < n_cand; base += 32) {
        const uint32_t c = base + lane;
        float4 p;
        if (staged) p = c < n_cand ? cand[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        else p = fetch(c);
        bool hit = false;
        unsigned long long k = 0;
        if (c < n_cand) {
          const float d2 = ref_d2(qp.x, qp.y, qp.z, p.x, p.y, p.z);
          hit = d2 < r2;
          k = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(p.w);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit && want_idx) {
          const int pos = total + __popc(m & ((1u << lane) - 1u));
          if (pos < kCellKeys) keys[pos] = k;
        }
        total += __popc(m);
      }
      finish_query<IdxT>(keys, total, lane, qi, width, ns_total, out_idx, out_counts, out_max, spill_list, spill_n,
                         kCellKeys);
      __syncwarp();
    }
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(kSpillThreads)
spill_kernel(const float* __restrict__ q, const int64_t* __restrict__ q_off, int batch,
             const GridGeom* __restrict__ geom, float inv_cell, float r2,
             const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tstart,
             const uint32_t* __restrict__ tcount, uint64_t tmask, const float4* __restrict__ sorted, int width,
             int64_t ns_total, IdxT* __restrict__ out_idx, const uint32_t* __restrict__ spill_list,
             const uint32_t* __restrict__ spill_n, int* __restrict__ err) {
  extern __shared__ unsigned long long sp_keys[];  // kSpillCap
  __shared__ uint32_t s_start[27], s_count[27];
  __shared__ int s_total;
  const uint32_t n_spill = *spill_n;
  for (uint32_t o = blockIdx.x; o < n_spill; o += gridDim.x) {
    const int64_t qi = spill_list[o];
    const int b = lcr_find_segment(q_off, batch, qi);
    const GridGeom g = geom[b];
    const float qx = q[3 * qi], qy = q[3 * qi + 1], qz = q[3 * qi + 2];
    if (threadIdx.x < 27) {
      const int l = threadIdx.x;
      const int nx = cell_coord(qx, g.ox, inv_cell) + (l % 3) - 1, ny = cell_coord(qy, g.oy, inv_cell) + ((l / 3) % 3) - 1,
                nz = cell_coord(qz, g.oz, inv_cell) + (l / 9) - 1;
      uint32_t st = 0, cn = 0;
      if (nx >= 0 && ny >= 0 && nz >= 0 && nx <= kCellMax && ny <= kCellMax && nz <= kCellMax)
        lookup_cell(cell_key(b, nx, ny, nz), tkeys, tstart, tcount, tmask, st, cn);
      s_start[l] = st;
      s_count[l] = cn;
    }
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    for (int c = 0; c < 27; c++) {
      const uint32_t st = s_start[c], cn = s_count[c];
      for (uint32_t t = threadIdx.x; t < cn; t += blockDim.x) {
        const float4 p = sorted[st + t];
        const float d2 = ref_d2(qx, qy, qz, p.x, p.y, p.z);
        if (d2 < r2) {
          const int pos = atomicAdd(&s_total, 1);
          if (pos < kSpillCap)
            sp_keys[pos] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)__float_as_uint(p.w);
        }
      }
    }
    __syncthreads();
    int total = s_total;
    if (total > kSpillCap) {
      if (threadIdx.x == 0) *err = LCR_ERR_OVERFLOW;  // more than 8192 neighbours for one query
      total = kSpillCap;
    }
    int n = 32;
    while (n < total) n <<= 1;
    for (int i = total + threadIdx.x; i < n; i += blockDim.x) sp_keys[i] = kEmpty;
    __syncthreads();
    bitonic_sort<true>(sp_keys, n, threadIdx.x, blockDim.x);
    IdxT* row = out_idx + (size_t)qi * width;
    for (int t = threadIdx.x; t < width; t += blockDim.x)
      row[t] = t < total ? (IdxT)(uint32_t)(sp_keys[t] & 0xFFFFFFFFull) : (IdxT)ns_total;
    __syncthreads();
  }
}

struct RadiusWs {
  int64_t *q_off, *s_off;
  unsigned* bbox;
  GridGeom* geom;
  unsigned long long* tkeys;
  uint32_t *tcount, *tstart, *tcursor, *partials, *scan_total;
  uint32_t* slot_of;
  float4* sorted;
  uint32_t *occ_list, *n_occ, *next_cell;
  uint32_t *spill_list, *spill_n;
  int* err;
  uint64_t tcap;
};

size_t carve(RadiusWs& w, void* ws, size_t ws_bytes, int64_t nq, int64_t ns, int batch) {
  LcrArena a(ws, ws_bytes);
  uint64_t cap = 1024;
  while (cap < (uint64_t)(2 * ns + 2)) cap <<= 1;
  w.tcap = cap;
  w.q_off = a.take<int64_t>(batch + 1);
  w.s_off = a.take<int64_t>(batch + 1);
  w.bbox = a.take<unsigned>(6 * (size_t)batch);
  w.geom = a.take<GridGeom>(batch);
  w.tkeys = a.take<unsigned long long>(cap);
  w.tcount = a.take<uint32_t>(cap);
  w.tstart = a.take<uint32_t>(cap);
  w.tcursor = a.take<uint32_t>(cap);
  w.partials = a.take<uint32_t>(1024);
  w.scan_total = a.take<uint32_t>(1);
  w.slot_of = a.take<uint32_t>(ns);
  w.sorted = a.take<float4>(ns);
  w.occ_list = a.take<uint32_t>(ns);
  w.n_occ = a.take<uint32_t>(1);
  w.next_cell = a.take<uint32_t>(1);
  w.spill_list = a.take<uint32_t>(nq);
  w.spill_n = a.take<uint32_t>(1);
  w.err = a.take<int>(1);
  return a.used;
}

}  // namespace

extern "C" size_t lcr_radius_neighbors_ws_bytes(int64_t nq_total, int64_t ns_total, int batch) {
  RadiusWs w;
  return carve(w, nullptr, 0, nq_total > 0 ? nq_total : 1, ns_total > 0 ? ns_total : 1, batch > 0 ? batch : 1);
}

extern "C" int lcr_radius_neighbors_ex(const float* q_points, int64_t nq_total, const float* s_points,
                                       int64_t ns_total, const int64_t* q_lengths, const int64_t* s_lengths, int batch,
                                       float radius, int width, void* out_idx, int idx_is64, int32_t* out_counts,
                                       int32_t* out_max_count, int32_t* out_status, void* ws, size_t ws_bytes,
                                       int reuse_grid, void* stream_);

extern "C" int lcr_radius_neighbors(const float* q_points, int64_t nq_total, const float* s_points,
                                    int64_t ns_total, const int64_t* q_lengths, const int64_t* s_lengths, int batch,
                                    float radius, int width, void* out_idx, int idx_is64, int32_t* out_counts,
                                    int32_t* out_max_count, int32_t* out_status, void* ws, size_t ws_bytes,
                                    void* stream_) {
  return lcr_radius_neighbors_ex(q_points, nq_total, s_points, ns_total, q_lengths, s_lengths, batch, radius, width,
                                 out_idx, idx_is64, out_counts, out_max_count, out_status, ws, ws_bytes, 0, stream_);
}

// reuse_grid is a bit set.  Bit 1 (value 2): nearest-only mode, width must be 1 -- column 0 of the full table without
// the sort.  Bit 0: the support grid in `ws` (built by a previous call with the SAME ws pointer, supports, support
// lengths, batch and radius) is reused; only the query side is processed.  The pyramid asks three tables of every
// support level at the same radius (self, subsampling of the next level, upsampling of the previous one,
// data.py:28-66): one grid build instead of three.
extern "C" int lcr_radius_neighbors_ex(const float* q_points, int64_t nq_total, const float* s_points,
                                       int64_t ns_total, const int64_t* q_lengths, const int64_t* s_lengths, int batch,
                                       float radius, int width, void* out_idx, int idx_is64, int32_t* out_counts,
                                       int32_t* out_max_count, int32_t* out_status, void* ws, size_t ws_bytes,
                                       int reuse_grid, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(batch >= 1 && batch < (1 << 20), "radius_neighbors: batch out of range");
  LCR_REQUIRE(nq_total >= 0 && ns_total >= 0 && nq_total < (1ll << 31) && ns_total < (1ll << 31),
              "radius_neighbors: sizes out of range");
  LCR_REQUIRE(radius > 0.f, "radius_neighbors: radius must be positive");
  LCR_REQUIRE(out_idx == nullptr || (width >= 1 && width <= LCR_RADIUS_MAX_WIDTH),
              "radius_neighbors: width out of range");
  LCR_REQUIRE(out_max_count != nullptr, "radius_neighbors: out_max_count is required");
  LCR_CUDA_TRY(cudaMemsetAsync(out_max_count, 0, sizeof(int32_t), stream));
  if (out_status) LCR_CUDA_TRY(cudaMemsetAsync(out_status, 0, sizeof(int32_t), stream));
  if (nq_total == 0) return LCR_OK;
  LCR_REQUIRE(q_points && q_lengths && s_lengths && ws, "radius_neighbors: null pointer");
  LCR_REQUIRE(ns_total == 0 || s_points, "radius_neighbors: null support pointer");
  RadiusWs w;
  const size_t need = carve(w, ws, ws_bytes, nq_total, ns_total > 0 ? ns_total : 1, batch);
  if (need > ws_bytes) {
    lcr_set_error("radius_neighbors: workspace too small", __FILE__, __LINE__);
    return LCR_ERR_WORKSPACE;
  }
  int* err = out_status ? out_status : w.err;
  const int T = 256;
  const float cell = radius * 1.001f;
  const float inv_cell = 1.0f / cell;
  const float r2 = radius * radius;  // radius_neighbors_cpu.cpp:12 (fp32 product)
  const double idx_bytes = out_idx ? (idx_is64 ? 8.0 : 4.0) * (double)nq_total * width : 0.0;
  LcrProfScope prof_all("radius_total", 0.0, 12.0 * (nq_total + ns_total) + idx_bytes, stream);
  lcr_offsets_launch(q_lengths, batch, w.q_off, stream);
  LCR_CUDA_TRY(cudaMemsetAsync(w.spill_n, 0, sizeof(uint32_t), stream));
  if (!out_status) LCR_CUDA_TRY(cudaMemsetAsync(w.err, 0, sizeof(int), stream));
  const bool nearest = (reuse_grid & 2) != 0;
  LCR_REQUIRE(!nearest || (out_idx && width == 1), "radius_neighbors: the nearest-only mode writes one column");
  const bool build = !(reuse_grid & 1);
  if (build) {
    lcr_offsets_launch(s_lengths, batch, w.s_off, stream);
    lcr_bbox_launch(s_points, ns_total, w.s_off, batch, w.bbox, stream);
    grid_geom_kernel<<<(batch + T - 1) / T, T, 0, stream>>>(w.bbox, batch, w.geom);
    LCR_CUDA_TRY(cudaMemsetAsync(w.tkeys, 0xFF, sizeof(unsigned long long) * w.tcap, stream));
    LCR_CUDA_TRY(cudaMemsetAsync(w.tcount, 0, sizeof(uint32_t) * w.tcap, stream));
    LCR_CUDA_TRY(cudaMemsetAsync(w.tcursor, 0, sizeof(uint32_t) * w.tcap, stream));
  }
  if (build && ns_total > 0) {
    const unsigned gridS = (unsigned)((ns_total + T - 1) / T);
    cell_insert_kernel<<<gridS, T, 0, stream>>>(s_points, ns_total, w.s_off, batch, w.geom, inv_cell, w.tkeys,
                                                w.tcount, w.tcap - 1, w.slot_of, err);
    int rc = lcr_scan_u32(w.tcount, w.tstart, (int64_t)w.tcap, w.scan_total, w.partials, stream);
    if (rc != LCR_OK) return rc;
    cell_scatter_kernel<<<gridS, T, 0, stream>>>(s_points, ns_total, w.slot_of, w.tstart, w.tcursor, w.sorted);
    LCR_CUDA_TRY(cudaMemsetAsync(w.n_occ, 0, sizeof(uint32_t), stream));
    compact_cells_kernel<<<(unsigned)((w.tcap + T - 1) / T), T, 0, stream>>>(w.tkeys, (uint32_t)w.tcap, w.occ_list, w.n_occ);
  }
  LcrProfScope prof("radius_query", 0.0, 12.0 * (nq_total + ns_total) + idx_bytes, stream);
  const unsigned gridQ = (unsigned)((nq_total + kWarpsPerCta - 1) / kWarpsPerCta);
  const size_t spill_smem = sizeof(unsigned long long) * kSpillCap;
  // self table: the queries ARE the supports (same arrays), so the support grid also groups the queries by cell
  static const int cell_mode = getenv("LCR_RADIUS_CELLS") ? atoi(getenv("LCR_RADIUS_CELLS")) : 1;
  // (measured, 64 scans: level 0, 909 k points: 1.10 -> 0.90 ms, level 1, 361 k: 0.46 -> 0.43 ms; the small levels
  // have too few cells to fill the persistent grid and stay on the per-query kernel.  What remains is the sort of
  // the hits: ~300 of the ~500 instructions per query.)
  const bool self = cell_mode && !nearest && q_points == s_points && q_lengths == s_lengths && nq_total == ns_total &&
                    (ns_total >= 200000 || cell_mode == 2);
  if (self) {
    LCR_CUDA_TRY(cudaMemsetAsync(w.next_cell, 0, sizeof(uint32_t), stream));
    const unsigned gridC = LCR_SM_COUNT * 5;     // 5 CTAs of 4 warps per SM (40 KB of shared memory each)
    if (idx_is64)
      query_self_kernel<int64_t><<<gridC, kCellWarps * 32, 0, stream>>>(
          w.occ_list, w.n_occ, w.next_cell, r2, w.tkeys, w.tstart, w.tcount, w.tcap - 1, w.sorted, width, ns_total,
          (int64_t*)out_idx, out_counts, out_max_count, w.spill_list, w.spill_n);
    else
      query_self_kernel<int32_t><<<gridC, kCellWarps * 32, 0, stream>>>(
          w.occ_list, w.n_occ, w.next_cell, r2, w.tkeys, w.tstart, w.tcount, w.tcap - 1, w.sorted, width, ns_total,
          (int32_t*)out_idx, out_counts, out_max_count, w.spill_list, w.spill_n);
  }
  if (nearest) {
    if (idx_is64)
      query_kernel<int64_t, true><<<gridQ, kWarpsPerCta * 32, 0, stream>>>(
          q_points, nq_total, w.q_off, batch, w.geom, inv_cell, r2, w.tkeys, w.tstart, w.tcount, w.tcap - 1, w.sorted,
          width, ns_total, (int64_t*)out_idx, out_counts, out_max_count, w.spill_list, w.spill_n);
    else
      query_kernel<int32_t, true><<<gridQ, kWarpsPerCta * 32, 0, stream>>>(
          q_points, nq_total, w.q_off, batch, w.geom, inv_cell, r2, w.tkeys, w.tstart, w.tcount, w.tcap - 1, w.sorted,
          width, ns_total, (int32_t*)out_idx, out_counts, out_max_count, w.spill_list, w.spill_n);
  } else if (!self) {
    if (idx_is64)
      query_kernel<int64_t><<<gridQ, kWarpsPerCta * 32, 0, stream>>>(
          q_points, nq_total, w.q_off, batch, w.geom, inv_cell, r2, w.tkeys, w.tstart, w.tcount, w.tcap - 1, w.sorted,
          width, ns_total, (int64_t*)out_idx, out_counts, out_max_count, w.spill_list, w.spill_n);
    else
      query_kernel<int32_t><<<gridQ, kWarpsPerCta * 32, 0, stream>>>(
          q_points, nq_total, w.q_off, batch, w.geom, inv_cell, r2, w.tkeys, w.tstart, w.tcount, w.tcap - 1, w.sorted,
          width, ns_total, (int32_t*)out_idx, out_counts, out_max_count, w.spill_list, w.spill_n);
  }
  if (out_idx && !nearest) {   // rows with more hits than a warp buffer: one CTA per queued query
    if (idx_is64) {
      static LcrOncePerDevice attr_set64;
      const int attr_set64_dev = attr_set64.need();
      if (attr_set64_dev != -1) {
        LCR_CUDA_TRY(cudaFuncSetAttribute(spill_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)spill_smem));
        attr_set64.done(attr_set64_dev);
      }
      spill_kernel<int64_t><<<LCR_SM_COUNT, kSpillThreads, spill_smem, stream>>>(
          q_points, w.q_off, batch, w.geom, inv_cell, r2, w.tkeys, w.tstart, w.tcount, w.tcap - 1, w.sorted, width,
          ns_total, (int64_t*)out_idx, w.spill_list, w.spill_n, err);
    } else {
      static LcrOncePerDevice attr_set32;
      const int attr_set32_dev = attr_set32.need();
      if (attr_set32_dev != -1) {
        LCR_CUDA_TRY(cudaFuncSetAttribute(spill_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)spill_smem));
        attr_set32.done(attr_set32_dev);
      }
      spill_kernel<int32_t><<<LCR_SM_COUNT, kSpillThreads, spill_smem, stream>>>(
          q_points, w.q_off, batch, w.geom, inv_cell, r2, w.tkeys, w.tstart, w.tcount, w.tcap - 1, w.sorted, width,
          ns_total, (int32_t*)out_idx, w.spill_list, w.spill_n, err);
    }
  }
  LCR_LAUNCHED(1 + (build ? 1 : 0) + (build && ns_total > 0 ? 3 : 0) + (out_idx && !nearest ? 1 : 0));  // query, geom, insert, scatter, compact, spill
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
