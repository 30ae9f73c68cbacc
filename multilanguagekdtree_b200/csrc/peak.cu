// peak.cu — FP64 pipe microbenchmark: the measured denominator of the walk's roofline
// (MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only).  Eight independent DFMA chains per thread.
#include "ctx.cuh"

namespace kdnb {

__global__ void __launch_bounds__(256) dfma_chain(double* out, int iters, double a, double b) {
  pdl_sync();
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int measure_fp64_peak(Ctx* c, double* tflops) {
  int sms = 0;
  KDNB_CUDA_TRY(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  KDNB_CUDA_TRY(c, cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, c->stream);
    KDNB_LAUNCH(c, dfma_chain, blocks, threads, 0, out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, c->stream);
    KDNB_CUDA_TRY(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 64.0 * (double)iters * blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return 0;
}

}  // namespace kdnb
