// FP64 FMA microbenchmark ("FP64 vector peak ... the builder must measure with an FMA
// microbenchmark"): the denominator for the element-stiffness assembly roofline.  Every thread runs 8 independent
// DFMA chains; prints achieved TFLOP/s (2 flops per FMA) for a few occupancies.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peak.cu -o tools/fp64_peak && tools/fp64_peak
#include <cuda_runtime.h>

#include <cstdio>

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b, int iters)
{
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc[c] = double(threadIdx.x + c) * 1e-3;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) acc[c] = fma(acc[c], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += acc[c];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keeps the chains alive, never taken in practice
}

int main()
{
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out = nullptr;
  cudaMalloc(&out, sizeof(double) * size_t(sms) * 8 * 256);
  constexpr int CHAINS = 8;
  const int     iters  = 4096;
  for (int ctas_per_sm : {1, 2, 4, 8}) {
    const int   grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_kernel<CHAINS><<<grid, 256>>>(out, 0.999999, 1e-9, iters);  // warm-up
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; r++) dfma_kernel<CHAINS><<<grid, 256>>>(out, 0.999999, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * CHAINS * 16.0 * iters * 256.0 * grid * reps;
    std::printf("{\"sms\": %d, \"ctas_per_sm\": %d, \"threads_per_sm\": %d, \"fp64_tflops\": %.2f}\n", sms, ctas_per_sm,
                ctas_per_sm * 256, flops / (ms * 1e-3) / 1e12);
  }
  if (cudaGetLastError() != cudaSuccess) {
    std::printf("CUDA error\n");
    return 1;
  }
  return 0;
}
