// Microbenchmark: FP64 tensor-core (mma.sync m8n8k4 f64) throughput and dependent latency on this GPU,
// next to the DFMA figures of fp64_probe.cu.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP>
__global__ void tput(double *out, int iters, double a, double b) {
  double c[ILP][2];
  for (int k = 0; k < ILP; ++k) { c[k][0] = threadIdx.x * 1e-9 + k; c[k][1] = k; }
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k) dmma(c[k][0], c[k][1], a, b);
  double s = 0; for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void chain(double *out, int iters, double a, double b) {
  double c0 = threadIdx.x * 1e-9, c1 = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = c0 + c1; out[1] = (double)(t1 - t0) / (4.0 * iters); }
}
int main() {
  double *d; cudaMalloc(&d, 1 << 24);
  chain<<<1, 32>>>(d, 10000, 1e-3, 1e-3); cudaDeviceSynchronize();
  double h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("DMMA m8n8k4 dependent latency: %.2f cycles\n", h[1]);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  for (int warps = 1; warps <= 32; warps *= 2) {
    tput<8><<<148, 32 * warps>>>(d, iters, 1e-3, 1e-3);
    cudaEventRecord(e0); tput<8><<<148, 32 * warps>>>(d, iters, 1e-3, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 148 * warps * 8.0 * iters * 256.0;   // 8x8x4 FMAs per warp instruction
    printf("warps/SM=%2d ILP=8: %.2f TFLOP/s FP64 (DMMA)\n", warps, flops / ms / 1e9);
  }
  return 0;
}
