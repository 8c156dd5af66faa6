// micro-benchmark: legacy mma.sync.m16n8k8 tf32 issue rate per SM
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NACC>
__global__ void k(float* out, int iters) {
  float d[NACC][4];
  unsigned a[4] = {threadIdx.x, threadIdx.x + 1, 3u, 4u}, b[2] = {5u, threadIdx.x};
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) mma_tf32(d[i], a, b);
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 8);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<8><<<pr.multiProcessorCount, warps * 32>>>(out, 100);
    cudaEventRecord(e0);
    k<8><<<pr.multiProcessorCount, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)pr.multiProcessorCount * warps * iters * 8;
    const double macs = mmas * 16 * 8 * 8;
    printf("warps/SM %2d: %.3f ms  %.1f TFLOP/s tf32 (mma.sync)  %.1f MAC/clk/SM at %d MHz nominal\n", warps, ms, 2 * macs / ms / 1e9,
           macs / pr.multiProcessorCount / (ms * 1e-3) / (clk * 1e3), clk / 1000);
  }
  return 0;
}
