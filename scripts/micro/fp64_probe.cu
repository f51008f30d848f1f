// Microbenchmarks: FP64 FMA dependent-issue latency and throughput on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(double *out, int iters, double a, double b) {
  double x = threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = x; out[1] = (double)(t1 - t0) / (4.0 * iters); }
}
template <int ILP>
__global__ void tput(double *out, int iters, double a, double b) {
  double x[ILP];
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-9 + k;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  double s = 0; for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *d; cudaMalloc(&d, 1 << 24);
  chain<<<1, 32>>>(d, 10000, 0.999, 1e-3); cudaDeviceSynchronize();
  double h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency: %.2f cycles\n", h[1]);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000;
  for (int warps = 1; warps <= 32; warps *= 2) {
    tput<8><<<148, 32 * warps>>>(d, iters, 0.999, 1e-3);
    cudaEventRecord(e0); tput<8><<<148, 32 * warps>>>(d, iters, 0.999, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 148 * 32 * warps * 8.0 * iters;
    printf("warps/SM=%2d ILP=8: %.2f TFLOP/s FP64\n", warps, flops / ms / 1e9);
  }
  return 0;
}
