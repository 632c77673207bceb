// Micro-benchmark: issue rate of mma.sync.m16n8k8 tf32 vs FFMA on this GPU (sizing input for the table update).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int ILP>
__global__ void k_mma(float* out, int iters) {
  float c[ILP][4] = {};
  unsigned a[4] = {threadIdx.x, 2, 3, 4}, b[2] = {5, threadIdx.x};
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < ILP; ++j) mma_tf32(c[j], a, b);
  float s = 0;
  for (int j = 0; j < ILP; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_ffma(float* out, int iters) {
  float c[ILP];
  for (int j = 0; j < ILP; ++j) c[j] = threadIdx.x * 0.001f + j;
  float a = 1.0001f, b = 0.5f;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < ILP; ++j) c[j] = fmaf(c[j], a, b);
  float s = 0;
  for (int j = 0; j < ILP; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_mma<8><<<148, warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double n = 148.0 * warps * iters * 8;
      if (rep) printf("mma.m16n8k8.tf32 warps/SM=%2d: %.1f Gmma/s  = %.1f TFLOP/s  (%.2f cycles/mma/SM @1.9GHz)\n", warps, n / ms / 1e6,
                      n * 2048 / ms / 1e9, ms * 1e-3 * 1.9e9 / (iters * 8.0 * warps));
    }
  }
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_ffma<8><<<148, warps * 32>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double n = 148.0 * warps * 32 * iters * 8;
      if (rep) printf("ffma warps/SM=%2d: %.1f TFLOP/s\n", warps, n * 2 / ms / 1e9);
    }
  }
  return 0;
}
