// Microbenchmark: legacy mma.sync.m16n8k8 TF32 issue rate on sm_100a (is the warp-level MMA path worth
// using for the KPConv influence contraction, whose M = 15 kernel points fits m16 exactly?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_tf32 mma_sync_tf32.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ILP>
__global__ void k(float* out, int iters) {
  float d[ILP][4];
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) d[i][j] = 0.f;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) mma_tf32(d[i], a, b);
  }
  float s = 0.f;
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void kf(float* out, int iters) {   // FFMA reference: same loop with 16 * ILP independent FFMA
  float d[ILP][16];
  const float a = threadIdx.x * 1e-9f, b = 1.0000001f;
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 16; j++) d[i][j] = (float)j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
#pragma unroll
      for (int j = 0; j < 16; j++) d[i][j] = fmaf(d[i][j], b, a);
  }
  float s = 0.f;
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 16; j++) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int blocks = 148 * 2, threads = warps * 32 / 2;
    k<4><<<blocks, threads>>>(out, 10);
    cudaEventRecord(e0);
    k<4><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)blocks * (threads / 32) * iters * 4;
    const double flops = mmas * 16 * 8 * 8 * 2;
    printf("mma.sync m16n8k8 tf32: %2d warps/SM  %.3f ms  %.1f TFLOP/s  %.2f MMA/clk/SM (1.965 GHz)\n", warps, ms,
           flops / ms / 1e9, mmas / 148 / (ms * 1e-3 * 1.965e9));
    kf<2><<<blocks, threads>>>(out, 10);
    cudaEventRecord(e0);
    kf<2><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double ffma = (double)blocks * threads * iters * 32;
    printf("FFMA                 : %2d warps/SM  %.3f ms  %.1f TFLOP/s\n", warps, ms, ffma * 2 / ms / 1e9);
  }
  return 0;
}
