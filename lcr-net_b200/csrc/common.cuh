// common.cuh -- shared device/host helpers for the lcr_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lcr_b200.h"

#define LCR_SM_COUNT 148  // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

#define LCR_CUDA_CHECK_LAUNCH()                                  \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) {                                    \
      lcr_set_error(cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LCR_ERR_CUDA;                                       \
    }                                                            \
  } while (0)

#define LCR_CUDA_TRY(x)                                          \
  do {                                                           \
    cudaError_t e__ = (x);                                       \
    if (e__ != cudaSuccess) {                                    \
      lcr_set_error(cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LCR_ERR_CUDA;                                       \
    }                                                            \
  } while (0)

#define LCR_REQUIRE(cond, msg)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      lcr_set_error(msg, __FILE__, __LINE__);          \
      return LCR_ERR_INVALID;                          \
    }                                                  \
  } while (0)

void lcr_set_error(const char* msg, const char* file, int line);
// Kernel-launch accounting (bench.py reports it as "gpu_launches").
void lcr_count_launches(int n);
#define LCR_LAUNCHED(n) lcr_count_launches(n)

// Optional per-kernel-group timing (lcr_profile_begin/end/get in the C ABI): when enabled, a
// scope records CUDA events on the launching stream around the launches it brackets, together
// with the group's ALGORITHMIC flops and bytes (the compulsory work).
struct LcrProfScope {
  int slot;
  cudaStream_t stream;
  LcrProfScope(const char* name, double flops, double bytes, cudaStream_t s);
  ~LcrProfScope();
};

// Opt-in dynamic shared memory is a PER-DEVICE function attribute: set it once per (kernel, device).
// Usage: static LcrOncePerDevice once; int d = once.need(); if (d != -1) { cudaFuncSetAttribute(...); once.done(d); }
struct LcrOncePerDevice {
  unsigned long long mask[2] = {0ull, 0ull};  // 128 devices; racing threads at worst set the attribute twice
  // current device id if the attribute still has to be set on it, -1 if already done; -2: unknown device (always set)
  int need() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return -2;
    return ((__atomic_load_n(&mask[dev >> 6], __ATOMIC_ACQUIRE) >> (dev & 63)) & 1ull) ? -1 : dev;
  }
  void done(int dev) {
    if (dev >= 0 && dev < 128) __atomic_fetch_or(&mask[dev >> 6], 1ull << (dev & 63), __ATOMIC_RELEASE);
  }
};

static inline size_t lcr_align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct LcrArena {
  char* base;
  size_t cap;
  size_t used;
  LcrArena(void* p, size_t n) : base((char*)p), cap(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = lcr_align_up(count * sizeof(T));
    char* r = base ? base + used : nullptr;
    used += bytes;
    return (T*)r;
  }
  bool ok() const { return base != nullptr && used <= cap; }
};

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ int lcr_lane() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T lcr_warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float lcr_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Largest b such that off[b] <= i  (off has nb+1 monotone entries, off[0] = 0).
__device__ __forceinline__ int lcr_find_segment(const int64_t* __restrict__ off, int nb, int64_t i) {
  int lo = 0, hi = nb;  // invariant: off[lo] <= i < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// Order-preserving float <-> uint mapping for atomicMin/Max on floats.
__device__ __forceinline__ unsigned lcr_f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float lcr_ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---------------------------------------------------------------- device-wide exclusive scan (u64)
// Three launches (block reduce -> single-block scan of partials -> block downsweep); no host sync.
// n is read from a device pointer OR given directly (n_dev == nullptr).
int lcr_scan_u64(const uint64_t* in, uint64_t* out, int64_t n, uint64_t* total_out, uint64_t* partials /*>=1024*/,
                 cudaStream_t stream);
int lcr_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total_out, uint32_t* partials /*>=1024*/,
                 cudaStream_t stream);

// lengths[batch] -> off[batch+1] (exclusive cumulative sum), on the device.
void lcr_offsets_launch(const int64_t* lengths, int batch, int64_t* off, cudaStream_t stream);
// Per-cloud bounding boxes of stacked points; bbox[6*b + {0,1,2}] = min xyz, {3,4,5} = max xyz,
// stored in the order-preserving uint encoding (lcr_f2ord / lcr_ord2f).
void lcr_bbox_launch(const float* pts, int64_t n, const int64_t* off, int batch, unsigned* bbox, cudaStream_t stream);

__device__ __forceinline__ uint64_t lcr_mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}
