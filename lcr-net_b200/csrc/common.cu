// common.cu -- error reporting + device-wide exclusive scans used by several kernels.
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void lcr_set_error(const char* msg, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s (%s:%d)", msg, file, line);
}

extern "C" const char* lcr_last_error(void) { return g_err; }
extern "C" int lcr_abi_version(void) { return 1; }

#include <atomic>
static std::atomic<long long> g_launches{0};
void lcr_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t lcr_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------ per-kernel-group profiling
#include <mutex>
#include <vector>
namespace {
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
  double flops, bytes;
};
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<float> g_prof_ms;
std::mutex g_prof_mu;
}  // namespace

LcrProfScope::LcrProfScope(const char* name, double flops, double bytes, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.name = name;
  r.flops = flops;
  r.bytes = bytes;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
LcrProfScope::~LcrProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].b, stream);
}

extern "C" void lcr_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_ms.clear();
  g_prof_on = true;
}
// Stops recording, waits for the recorded work and returns the number of records.
extern "C" int lcr_profile_end(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  g_prof_ms.assign(g_prof.size(), 0.f);
  for (size_t i = 0; i < g_prof.size(); i++) {
    cudaEventSynchronize(g_prof[i].b);
    cudaEventElapsedTime(&g_prof_ms[i], g_prof[i].a, g_prof[i].b);
  }
  return (int)g_prof.size();
}
extern "C" int lcr_profile_get(int i, char* name, int name_cap, double* ms, double* flops, double* bytes) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (i < 0 || i >= (int)g_prof.size() || name_cap < 1) return LCR_ERR_INVALID;
  snprintf(name, (size_t)name_cap, "%s", g_prof[i].name);
  *ms = (double)g_prof_ms[i];
  *flops = g_prof[i].flops;
  *bytes = g_prof[i].bytes;
  return LCR_OK;
}

// ------------------------------------------------------------------ scan
// Layout: up to 1024 blocks, each block owns a contiguous chunk of `chunk` elements
// (chunk is a multiple of the block size); phase 1 reduces chunks, phase 2 scans the <=1024
// partials in one block, phase 3 rescans each chunk with its carry-in.
namespace {
constexpr int kScanThreads = 256;
constexpr int kScanMaxBlocks = 1024;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /* >= 32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = lane < (blockDim.x >> 5) ? smem[lane] : T(0);
    T wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T n = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += n;
    }
    smem[lane] = wi - w;  // exclusive prefix of warp totals
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  T res = incl - v + smem[warp];
  *total = smem[32];
  __syncthreads();
  return res;
}

template <typename T>
__global__ void scan_reduce_kernel(const T* __restrict__ in, int64_t n, int64_t chunk, T* __restrict__ partials) {
  __shared__ T sm[33];
  const int64_t beg = (int64_t)blockIdx.x * chunk;
  const int64_t end = min(beg + chunk, n);
  T acc = 0;
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) acc += in[i];
  T tot;
  block_exclusive_scan<T>(acc, &tot, sm);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

template <typename T>
__global__ void scan_partials_kernel(T* __restrict__ partials, int nblocks, T* __restrict__ total_out) {
  __shared__ T sm[33];
  T carry = 0;
  for (int base = 0; base < nblocks; base += blockDim.x) {
    int i = base + threadIdx.x;
    T v = i < nblocks ? partials[i] : T(0);
    T tot;
    T ex = block_exclusive_scan<T>(v, &tot, sm);
    if (i < nblocks) partials[i] = ex + carry;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename T>
__global__ void scan_down_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n, int64_t chunk,
                                 const T* __restrict__ partials) {
  __shared__ T sm[33];
  const int64_t beg = (int64_t)blockIdx.x * chunk;
  const int64_t end = min(beg + chunk, n);
  T carry = partials[blockIdx.x];
  for (int64_t base = beg; base < end; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    T v = i < end ? in[i] : T(0);
    T tot;
    T ex = block_exclusive_scan<T>(v, &tot, sm);
    if (i < end) out[i] = ex + carry;
    carry += tot;
  }
}

template <typename T>
int scan_impl(const T* in, T* out, int64_t n, T* total_out, T* partials, cudaStream_t stream) {
  if (n <= 0) {
    if (total_out) cudaMemsetAsync(total_out, 0, sizeof(T), stream);
    return LCR_OK;
  }
  int64_t chunk = (n + kScanMaxBlocks - 1) / kScanMaxBlocks;
  chunk = (chunk + kScanThreads - 1) / kScanThreads * kScanThreads;
  int nblocks = (int)((n + chunk - 1) / chunk);
  scan_reduce_kernel<T><<<nblocks, kScanThreads, 0, stream>>>(in, n, chunk, partials);
  scan_partials_kernel<T><<<1, kScanThreads, 0, stream>>>(partials, nblocks, total_out);
  scan_down_kernel<T><<<nblocks, kScanThreads, 0, stream>>>(in, out, n, chunk, partials);
  LCR_LAUNCHED(3);
  LCR_CUDA_CHECK_LAUNCH();
  return LCR_OK;
}
}  // namespace

int lcr_scan_u64(const uint64_t* in, uint64_t* out, int64_t n, uint64_t* total_out, uint64_t* partials,
                 cudaStream_t stream) {
  return scan_impl<unsigned long long>((const unsigned long long*)in, (unsigned long long*)out, n,
                                       (unsigned long long*)total_out, (unsigned long long*)partials, stream);
}
int lcr_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total_out, uint32_t* partials,
                 cudaStream_t stream) {
  return scan_impl<uint32_t>(in, out, n, total_out, partials, stream);
}


// ------------------------------------------------------------------ offsets / bounding boxes
namespace {
__global__ void offsets_kernel(const int64_t* __restrict__ lengths, int batch, int64_t* __restrict__ off) {
  // one warp: chunked warp scan (batch is the number of clouds per launch, at most a few thousand)
  const int lane = threadIdx.x;
  int64_t carry = 0;
  for (int base = 0; base < batch; base += 32) {
    const int b = base + lane;
    const int64_t v = b < batch ? lengths[b] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (b < batch) off[b] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) off[batch] = carry;
}

__global__ void bbox_init_kernel(unsigned* __restrict__ bbox, int batch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch * 6) bbox[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;
}

// Every thread folds kBboxPer points (strided by the block size: coalesced) into a running box and
// flushes it when the cloud changes; at the end a warp whose lanes all sit in the same cloud reduces by
// shuffle and issues one set of atomics.  ~30 same-address atomics per cloud instead of one per warp of
// points (which serialised in L2: 22 us per call for 470 k points).
constexpr int kBboxPer = 16;
__device__ __forceinline__ void bbox_flush(unsigned* __restrict__ bbox, int b, const float* mn, const float* mx) {
  unsigned* bb = bbox + 6 * b;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    atomicMin(bb + d, lcr_f2ord(mn[d]));
    atomicMax(bb + 3 + d, lcr_f2ord(mx[d]));
  }
}

__global__ void __launch_bounds__(256)
bbox_kernel(const float* __restrict__ pts, int64_t n, const int64_t* __restrict__ off, int batch,
            unsigned* __restrict__ bbox) {
  const int64_t base = (int64_t)blockIdx.x * (blockDim.x * kBboxPer) + threadIdx.x;
  int b = -1;
  int64_t next = 0;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = 0; k < kBboxPer; k++) {
    const int64_t i = base + (int64_t)k * blockDim.x;
    if (i >= n) break;
    if (b < 0 || i >= next) {
      if (b >= 0) bbox_flush(bbox, b, mn, mx);
      b = lcr_find_segment(off, batch, i);
      next = off[b + 1];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        mn[d] = INFINITY;
        mx[d] = -INFINITY;
      }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const float v = pts[3 * i + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
    }
  }
  const int b0 = __shfl_sync(0xffffffffu, b, 0);
  if (__all_sync(0xffffffffu, b == b0)) {
    if (b0 < 0) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int d = 0; d < 3; d++) {
        mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
        mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
      }
    if (lcr_lane() == 0) bbox_flush(bbox, b0, mn, mx);
  } else if (b >= 0) {
    bbox_flush(bbox, b, mn, mx);
  }
}
}  // namespace

void lcr_offsets_launch(const int64_t* lengths, int batch, int64_t* off, cudaStream_t stream) {
  offsets_kernel<<<1, 32, 0, stream>>>(lengths, batch, off);
  LCR_LAUNCHED(1);
}

void lcr_bbox_launch(const float* pts, int64_t n, const int64_t* off, int batch, unsigned* bbox, cudaStream_t stream) {
  const int T = 256;
  bbox_init_kernel<<<(batch * 6 + T - 1) / T, T, 0, stream>>>(bbox, batch);
  if (n > 0) bbox_kernel<<<(unsigned)((n + T * kBboxPer - 1) / (T * kBboxPer)), T, 0, stream>>>(pts, n, off, batch, bbox);
  LCR_LAUNCHED(n > 0 ? 2 : 1);
}
