// subsample.cu -- a1: voxel-grid subsampling (stack mode) on sm_100a.
//
// Reference semantics reproduced bit-for-bit (utils/extensions/cpu/grid_subsampling/
// grid_subsampling_cpu.cpp:3-75, grid_subsampling_cpu.h:7-21, extra/cloud/cloud.cpp:4-37):
//   origin = floor(min * (float)(1.0/voxel)) * voxel ; nx,ny from (max-origin)/voxel ;
//   key = ix + nx*iy + nx*ny*iz with i* = floor((p-origin)/voxel) (true fp32 division) ;
//   centroid = (sequential fp32 sum in input order) * (float)(1.0/count) ;
//   output order = iteration order of libstdc++ std::unordered_map<size_t,...>.
//
// Design (HBM-bound integer/byte work, no tensor cores):
//   1. bbox per cloud (warp-aggregated float atomics).
//   2. open-addressing hash insert of (cloud,key): per slot first point index (atomicMin),
//      count and an atomicExch-built chain of member points.
//   3. ONE device-wide scan over points of (is_first ? (1,count) : (0,0)) gives both the voxel
//      rank in first-seen order and the start of its member list.
//   4. every point ranks itself inside its voxel by walking the (short) chain -> members are
//      laid out in ascending input index -> one thread per voxel adds them sequentially, which
//      is exactly the reference's accumulation order (no float atomics anywhere).
//   5. reference order: the unordered_map's final list order is a pure function of the
//      first-seen key sequence.  Re-hashing is equivalent to re-inserting the current list, so
//      the order is obtained by ~log2(m) parallel "group by bucket" rounds (one CTA per cloud),
//      see emulate_order_kernel.  Verified bit-exact against the compiled reference.
// No host synchronisation: all sizes stay on the device; outputs are capacity-sized.
#include "common.cuh"

namespace {

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kKeyBits = 44;  // packed = cloud << 44 | key
constexpr uint64_t kKeyMask = (1ull << kKeyBits) - 1;
constexpr int kCntBits = 36;  // scan word = voxel_count << 36 | member_count
constexpr uint64_t kCntMask = (1ull << kCntBits) - 1;

// libstdc++ (GCC 13) bucket-count sequence for single inserts at max_load_factor 1.
__constant__ uint32_t c_bucket_seq[24] = {1u,       13u,       29u,       59u,       127u,     257u,
                                          541u,     1109u,     2357u,     5087u,     10273u,   20753u,
                                          42043u,   85229u,    172933u,   351061u,   712697u,  1447153u,
                                          2938679u, 5967347u,  12117689u, 24607243u, 49969847u, 101473717u};

struct CloudGeom {
  float ox, oy, oz;
  float voxel;
  uint64_t nx, nxy;
};

__global__ void geom_kernel(const unsigned* __restrict__ bbox, int batch, float voxel, CloudGeom* __restrict__ geom) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float mnx = lcr_ord2f(bbox[6 * b]), mny = lcr_ord2f(bbox[6 * b + 1]), mnz = lcr_ord2f(bbox[6 * b + 2]);
  const float mxx = lcr_ord2f(bbox[6 * b + 3]), mxy = lcr_ord2f(bbox[6 * b + 4]);
  // PointXYZ * (1. / voxel_size): double reciprocal narrowed to the float operand (cloud.h:84).
  const float inv = __double2float_rn(1.0 / (double)voxel);
  CloudGeom g;
  g.voxel = voxel;
  g.ox = __fmul_rn(floorf(__fmul_rn(mnx, inv)), voxel);
  g.oy = __fmul_rn(floorf(__fmul_rn(mny, inv)), voxel);
  g.oz = __fmul_rn(floorf(__fmul_rn(mnz, inv)), voxel);
  g.nx = (uint64_t)(int64_t)(floorf(__fdiv_rn(__fsub_rn(mxx, g.ox), voxel)) + 1.0f);
  uint64_t ny = (uint64_t)(int64_t)(floorf(__fdiv_rn(__fsub_rn(mxy, g.oy), voxel)) + 1.0f);
  g.nxy = g.nx * ny;
  geom[b] = g;
}

__global__ void insert_kernel(const float* __restrict__ pts, int64_t n, const int64_t* __restrict__ off, int batch,
                              const CloudGeom* __restrict__ geom, unsigned long long* __restrict__ tkeys,
                              uint32_t* __restrict__ tfirst, uint32_t* __restrict__ tcount,
                              uint32_t* __restrict__ thead, uint64_t tmask, uint32_t* __restrict__ slot_of,
                              uint32_t* __restrict__ next, int* __restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = lcr_find_segment(off, batch, i);
  const CloudGeom g = geom[b];
  const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  const uint64_t ix = (uint64_t)(int64_t)floorf(__fdiv_rn(__fsub_rn(x, g.ox), g.voxel));
  const uint64_t iy = (uint64_t)(int64_t)floorf(__fdiv_rn(__fsub_rn(y, g.oy), g.voxel));
  const uint64_t iz = (uint64_t)(int64_t)floorf(__fdiv_rn(__fsub_rn(z, g.oz), g.voxel));
  uint64_t key = ix + g.nx * iy + g.nxy * iz;
  if (key > kKeyMask) {  // result is invalid but stays memory-safe; reported through *err
    *err = LCR_ERR_OVERFLOW;
    key &= kKeyMask;
  }
  const uint64_t packed = ((uint64_t)b << kKeyBits) | key;
  uint64_t h = lcr_mix64(packed) & tmask;
  while (true) {
    unsigned long long old = atomicCAS(&tkeys[h], (unsigned long long)kEmptyKey, (unsigned long long)packed);
    if (old == kEmptyKey || old == packed) break;
    h = (h + 1) & tmask;
  }
  slot_of[i] = (uint32_t)h;
  atomicMin(&tfirst[h], (uint32_t)i);
  atomicAdd(&tcount[h], 1u);
  next[i] = atomicExch(&thead[h], (uint32_t)i);
}

__global__ void flag_kernel(int64_t n, const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ tfirst,
                            const uint32_t* __restrict__ tcount, uint64_t* __restrict__ scan_in) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = slot_of[i];
  scan_in[i] = (tfirst[s] == (uint32_t)i) ? ((1ull << kCntBits) | (uint64_t)tcount[s]) : 0ull;
}

// Each point: rank among the members of its voxel (ascending input index) -> member list.
__global__ void rank_kernel(int64_t n, const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ next,
                            const unsigned long long* __restrict__ tkeys, const uint32_t* __restrict__ tfirst,
                            const uint32_t* __restrict__ tcount, const uint32_t* __restrict__ thead,
                            const uint64_t* __restrict__ scan_ex, uint32_t* __restrict__ members,
                            uint64_t* __restrict__ vkey, uint32_t* __restrict__ vstart, uint32_t* __restrict__ vcnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = slot_of[i];
  const uint32_t f = tfirst[s];
  const uint64_t ex = scan_ex[f];
  const uint32_t v = (uint32_t)(ex >> kCntBits);
  const uint32_t start = (uint32_t)(ex & kCntMask);
  uint32_t rank = 0;
  for (uint32_t j = thead[s]; j != kNone; j = next[j]) rank += (j < (uint32_t)i);
  members[start + rank] = (uint32_t)i;
  if (f == (uint32_t)i) {
    vkey[v] = tkeys[s] & kKeyMask;
    vstart[v] = start;
    vcnt[v] = tcount[s];
  }
}

// Per-cloud voxel base (first-seen rank of the first voxel of each cloud) and output lengths.
__global__ void lengths_kernel(const int64_t* __restrict__ off, int batch, int64_t n,
                               const uint64_t* __restrict__ scan_ex, const uint64_t* __restrict__ scan_total,
                               uint32_t* __restrict__ vbase, int64_t* __restrict__ out_lengths,
                               int64_t* __restrict__ out_total) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > batch) return;
  const uint32_t total = (uint32_t)(*scan_total >> kCntBits);
  uint32_t lo = off[b] < n ? (uint32_t)(scan_ex[off[b]] >> kCntBits) : total;
  vbase[b] = lo;
  if (b < batch) {
    uint32_t hi = off[b + 1] < n ? (uint32_t)(scan_ex[off[b + 1]] >> kCntBits) : total;
    out_lengths[b] = (int64_t)hi - (int64_t)lo;
  } else if (out_total) {
    *out_total = (int64_t)total;
  }
}

// ---- reference (libstdc++ unordered_map) iteration order, one CTA per cloud ------------------
// State after inserting keys k_0..k_{t-1} (distinct, in first-seen order) with bucket count nb:
// a singly linked list grouped by bucket; a key landing in an empty bucket goes to the list
// head, otherwise to the front of its bucket's group (_M_insert_bucket_begin); a rehash walks
// the list and re-inserts every node by the same rule (_M_rehash_aux), i.e. it equals inserting
// the current list order from scratch.  Hence
//     list_k = order_{nb_k}( list_{k-1} ++ [keys inserted while the bucket count is nb_k] )
// where order_nb(seq) sorts by (first position of the element's bucket, own position), reversed.
// Each round is data-parallel: chains per bucket via atomicExch (load factor <= 1 so they are
// short), one block-wide scan for the group starts.
constexpr int kEmuThreads = 1024;

__device__ __forceinline__ uint32_t emu_block_scan(uint32_t v, uint32_t* total, uint32_t* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sm[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = sm[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    sm[lane] = wi - w;
    if (lane == 31) sm[32] = wi;
  }
  __syncthreads();
  uint32_t r = incl - v + sm[warp];
  *total = sm[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kEmuThreads)
emulate_order_kernel(const uint32_t* __restrict__ vbase, const uint64_t* __restrict__ vkey,
                     uint32_t* __restrict__ cur_all, uint32_t* __restrict__ nxt_all,
                     uint32_t* __restrict__ chain_all, uint32_t* __restrict__ bkt_all,
                     uint32_t* __restrict__ rank_all, uint32_t* __restrict__ first_all,
                     uint32_t* __restrict__ gsum_all, uint32_t* __restrict__ head_all,
                     uint32_t* __restrict__ pos_out) {
  __shared__ uint32_t sm[33];
  const int b = blockIdx.x;
  const uint32_t base = vbase[b];
  const uint32_t m = vbase[b + 1] - base;
  if (m == 0) return;
  const uint64_t* key = vkey + base;
  uint32_t* cur = cur_all + base;
  uint32_t* nxt = nxt_all + base;
  uint32_t* chain = chain_all + base;
  uint32_t* bkt = bkt_all + base;
  uint32_t* rnk = rank_all + base;
  uint32_t* fst = first_all + base;
  uint32_t* gsum = gsum_all + base;
  // bucket heads: sum over earlier clouds of (2.2*m + 14) >= their bucket counts
  uint32_t* head = head_all + (size_t)((double)base * 2.2) + (size_t)14 * b;

  uint32_t len = 0;
  for (int k = 1; k < 24; k++) {
    const uint32_t nb = c_bucket_seq[k];
    const uint32_t mk = min(nb, m);
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) head[i] = kNone;
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < mk; p += blockDim.x) {
      const uint32_t e = p < len ? cur[p] : p;
      if (p >= len) cur[p] = e;
      const uint32_t bk = (uint32_t)(key[e] % (uint64_t)nb);
      bkt[p] = bk;
      chain[p] = atomicExch(&head[bk], p);
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < mk; p += blockDim.x) {
      uint32_t first = p, size = 0, r = 0;
      for (uint32_t j = head[bkt[p]]; j != kNone; j = chain[j]) {
        first = min(first, j);
        size++;
        r += (j < p);
      }
      rnk[p] = r;
      fst[p] = first;
      gsum[p] = (first == p) ? size : 0u;
    }
    __syncthreads();
    uint32_t carry = 0;
    for (uint32_t basep = 0; basep < mk; basep += blockDim.x) {
      const uint32_t p = basep + threadIdx.x;
      const uint32_t v = p < mk ? gsum[p] : 0u;
      uint32_t tot;
      const uint32_t ex = emu_block_scan(v, &tot, sm);
      if (p < mk) gsum[p] = ex + carry;
      carry += tot;
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < mk; p += blockDim.x) {
      const uint32_t asc = gsum[fst[p]] + rnk[p];
      nxt[mk - 1 - asc] = cur[p];
    }
    __syncthreads();
    uint32_t* t = cur;
    cur = nxt;
    nxt = t;
    len = mk;
    if (nb >= m) break;
  }
  for (uint32_t p = threadIdx.x; p < m; p += blockDim.x) pos_out[base + cur[p]] = p;
}

// One thread per voxel: sequential fp32 accumulation in ascending input index.
__global__ void centroid_kernel(const float* __restrict__ pts, const uint64_t* __restrict__ scan_total,
                                const uint32_t* __restrict__ members, const uint32_t* __restrict__ vstart,
                                const uint32_t* __restrict__ vcnt, const uint32_t* __restrict__ vbase, int batch,
                                const uint32_t* __restrict__ pos /* nullable */, float* __restrict__ out) {
  const uint32_t total = (uint32_t)(*scan_total >> kCntBits);
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total) return;
  const uint32_t start = vstart[v], cnt = vcnt[v];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (uint32_t k = 0; k < cnt; k++) {
    const uint32_t i = members[start + k];
    sx = __fadd_rn(sx, pts[3 * (size_t)i]);
    sy = __fadd_rn(sy, pts[3 * (size_t)i + 1]);
    sz = __fadd_rn(sz, pts[3 * (size_t)i + 2]);
  }
  const float r = __double2float_rn(1.0 / (double)cnt);
  uint32_t dst = v;
  if (pos) {
    // cloud of voxel v: vbase is monotone with batch+1 entries
    int lo = 0, hi = batch;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (vbase[mid] <= v) lo = mid; else hi = mid;
    }
    dst = vbase[lo] + pos[v];
  }
  out[3 * (size_t)dst] = __fmul_rn(sx, r);
  out[3 * (size_t)dst + 1] = __fmul_rn(sy, r);
  out[3 * (size_t)dst + 2] = __fmul_rn(sz, r);
}

struct SubsampleWs {
  int64_t* off;
  unsigned* bbox;
  CloudGeom* geom;
  unsigned long long* tkeys;
  uint32_t *tfirst, *tcount, *thead;
  uint32_t *slot_of, *next, *members;
  uint64_t *scan_in, *scan_ex, *scan_total, *partials;
  uint64_t* vkey;
  uint32_t *vstart, *vcnt, *vbase;
  uint32_t *e_cur, *e_nxt, *e_chain, *e_bkt, *e_rank, *e_first, *e_gsum, *e_head, *pos;
  int* err;
  uint64_t tcap;
  size_t head_cap;
};

size_t carve(SubsampleWs& w, void* ws, size_t ws_bytes, int64_t n, int batch) {
  LcrArena a(ws, ws_bytes);
  uint64_t cap = 1024;
  while (cap < (uint64_t)(2 * n + 2)) cap <<= 1;
  w.tcap = cap;
  w.head_cap = (size_t)(2.2 * (double)n) + (size_t)14 * (batch + 1) + 64;
  w.off = a.take<int64_t>(batch + 1);
  w.bbox = a.take<unsigned>(6 * (size_t)batch);
  w.geom = a.take<CloudGeom>(batch);
  w.tkeys = a.take<unsigned long long>(cap);
  w.tfirst = a.take<uint32_t>(cap);
  w.tcount = a.take<uint32_t>(cap);
  w.thead = a.take<uint32_t>(cap);
  w.slot_of = a.take<uint32_t>(n);
  w.next = a.take<uint32_t>(n);
  w.members = a.take<uint32_t>(n);
  w.scan_in = a.take<uint64_t>(n);
  w.scan_ex = a.take<uint64_t>(n);
  w.scan_total = a.take<uint64_t>(1);
  w.partials = a.take<uint64_t>(1024);
  w.vkey = a.take<uint64_t>(n);
  w.vstart = a.take<uint32_t>(n);
  w.vcnt = a.take<uint32_t>(n);
  w.vbase = a.take<uint32_t>(batch + 2);
  w.e_cur = a.take<uint32_t>(n);
  w.e_nxt = a.take<uint32_t>(n);
  w.e_chain = a.take<uint32_t>(n);
  w.e_bkt = a.take<uint32_t>(n);
  w.e_rank = a.take<uint32_t>(n);
  w.e_first = a.take<uint32_t>(n);
  w.e_gsum = a.take<uint32_t>(n);
  w.e_head = a.take<uint32_t>(w.head_cap);
  w.pos = a.take<uint32_t>(n);
  w.err = a.take<int>(1);
  return a.used;
}

}  // namespace

extern "C" size_t lcr_grid_subsample_ws_bytes(int64_t n_total, int batch) {
  SubsampleWs w;
  return carve(w, nullptr, 0, n_total > 0 ? n_total : 1, batch > 0 ? batch : 1);
}

extern "C" int lcr_grid_subsample(const float* points, int64_t n_total, const int64_t* lengths, int batch,
                                  float voxel_size, int order_mode, float* out_points, int64_t* out_lengths,
                                  int64_t* out_total, int32_t* out_status, void* ws, size_t ws_bytes,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LCR_REQUIRE(batch >= 1 && batch < (1 << 19), "grid_subsample: batch out of range");
  LCR_REQUIRE(n_total >= 0 && n_total < (1ll << 28), "grid_subsample: n_total out of range");
  LCR_REQUIRE(voxel_size > 0.f, "grid_subsample: voxel_size must be positive");
  LCR_REQUIRE(order_mode == 0 || order_mode == 1, "grid_subsample: order_mode must be 0 or 1");
  if (out_status) LCR_CUDA_TRY(cudaMemsetAsync(out_status, 0, sizeof(int32_t), stream));
  if (n_total == 0) {
    LCR_CUDA_TRY(cudaMemsetAsync(out_lengths, 0, sizeof(int64_t) * batch, stream));
    if (out_total) LCR_CUDA_TRY(cudaMemsetAsync(out_total, 0, sizeof(int64_t), stream));
    return LCR_OK;
  }
  LCR_REQUIRE(points && lengths && out_points && out_lengths && ws, "grid_subsample: null pointer");
  SubsampleWs w;
  size_t need = carve(w, ws, ws_bytes, n_total, batch);
  if (need > ws_bytes) {
    lcr_set_error("grid_subsample: workspace too small", __FILE__, __LINE__);
    return LCR_ERR_WORKSPACE;
  }
  LcrProfScope prof("grid_subsample", 0.0, 24.0 * n_total + 8.0 * batch, stream);
  const int T = 256;
  const unsigned gridN = (unsigned)((n_total + T - 1) / T);
  lcr_offsets_launch(lengths, batch, w.off, stream);
  lcr_bbox_launch(points, n_total, w.off, batch, w.bbox, stream);
  geom_kernel<<<(batch + T - 1) / T, T, 0, stream>>>(w.bbox, batch, voxel_size, w.geom);
  LCR_CUDA_TRY(cudaMemsetAsync(w.tkeys, 0xFF, sizeof(unsigned long long) * w.tcap, stream));
  LCR_CUDA_TRY(cudaMemsetAsync(w.tfirst, 0xFF, sizeof(uint32_t) * w.tcap, stream));
  LCR_CUDA_TRY(cudaMemsetAsync(w.tcount, 0, sizeof(uint32_t) * w.tcap, stream));
  LCR_CUDA_TRY(cudaMemsetAsync(w.thead, 0xFF, sizeof(uint32_t) * w.tcap, stream));
  LCR_CUDA_TRY(cudaMemsetAsync(w.err, 0, sizeof(int), stream));
  insert_kernel<<<gridN, T, 0, stream>>>(points, n_total, w.off, batch, w.geom, w.tkeys, w.tfirst, w.tcount, w.thead,
                                         w.tcap - 1, w.slot_of, w.next, out_status ? out_status : w.err);
  flag_kernel<<<gridN, T, 0, stream>>>(n_total, w.slot_of, w.tfirst, w.tcount, w.scan_in);
  int rc = lcr_scan_u64(w.scan_in, w.scan_ex, n_total, w.scan_total, w.partials, stream);
  if (rc != LCR_OK) return rc;
  rank_kernel<<<gridN, T, 0, stream>>>(n_total, w.slot_of, w.next, w.tkeys, w.tfirst, w.tcount, w.thead, w.scan_ex,
                                       w.members, w.vkey, w.vstart, w.vcnt);
  lengths_kernel<<<(batch + 1 + T - 1) / T, T, 0, stream>>>(w.off, batch, n_total, w.scan_ex, w.scan_total, w.vbase,
                                                            out_lengths, out_total);
  if (order_mode == 1) {
    emulate_order_kernel<<<batch, kEmuThreads, 0, stream>>>(w.vbase, w.vkey, w.e_cur, w.e_nxt, w.e_chain, w.e_bkt,
                                                            w.e_rank, w.e_first, w.e_gsum, w.e_head, w.pos);
  }
  centroid_kernel<<<gridN, T, 0, stream>>>(points, w.scan_total, w.members, w.vstart, w.vcnt, w.vbase, batch,
                                           order_mode == 1 ? w.pos : nullptr, out_points);
  LCR_LAUNCHED(order_mode == 1 ? 7 : 6);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
