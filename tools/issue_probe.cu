// issue_probe.cu — development aid (NOT product, NOT test): does an FP64 warp instruction block its scheduler's
// issue port for both of the cycles the half-rate FP64 pipe needs, or can instructions for other pipes issue in its
// shadow?  Each loop iteration runs 16 independent DFMAs plus K filler instructions (ALU xor/add, FSEL, or shared-memory
// broadcast loads); the output is cycles per iteration per scheduler.  If K fillers are free up to K=16 the walk's drain
// loop is bound by the FP64 pipe alone; if every filler adds a cycle it is bound by issue slots (FP64 counted twice).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/issue_probe tools/issue_probe.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

template <int K, int MODE>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double a, double b, unsigned ka, unsigned kb) {
  __shared__ double4 sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = make_double4(a, b, a, b);
  __syncthreads();
  double x[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) x[u] = threadIdx.x + u;
  unsigned v[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
  float f[4] = {1.f, 2.f, 3.f, 4.f};
  double acc = 0.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = fma(x[u], a, b);
#pragma unroll
      for (int k = 0; k < K / 2; ++k) {
        if (MODE == 0) {  // ALU: xor / add on four independent chains
          if (((k >> 2) + r * ((K / 2 + 3) / 4)) & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(v[k & 3]) : "r"(ka));
          else asm volatile("add.u32 %0, %0, %1;" : "+r"(v[k & 3]) : "r"(kb));
        } else if (MODE == 1) {  // IMAD (FMA-lite/heavy pipe)
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[k & 3]) : "r"(ka), "r"(kb));
        } else {  // LDS.128 broadcast
          double t0, t1;
          const unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[(i + k) & 63]);
          asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t0), "=d"(t1) : "r"(addr));
          asm volatile("" ::"d"(t0), "d"(t1));
        }
      }
    }
  }
  double s = acc;
#pragma unroll
  for (int u = 0; u < 8; ++u) s += x[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(v[0] ^ v[1] ^ v[2] ^ v[3]) + f[0] + f[1] + f[2] + f[3];
}

template <int K, int MODE>
static void run(double* out, int sms, int warps_per_smsp, int clock_khz) {
  const int threads = 256, blocks = sms * warps_per_smsp / 2, iters = 8192;  // 8 warps per CTA = 2 per scheduler
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<K, MODE><<<blocks, threads>>>(out, iters, 0.999999, 1e-9, 12345u, 77u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  // cycles per iteration per scheduler: each scheduler runs warps_per_smsp warps
  const double cyc = best * 1e-3 * clock_khz * 1e3 / iters;
  printf("mode=%d K=%2d warps/smsp=%d: %.3f ms  %.1f cycles per iteration per scheduler  (%.2f per warp-iteration; 16 DFMA + %d fillers)\n",
         MODE, K, warps_per_smsp, best, cyc, cyc / warps_per_smsp, K);
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * 256 * sms * 16);
  printf("SMs=%d clock=%d kHz\n", sms, khz);
  for (int w : {8, 4, 2}) {
    run<0, 0>(out, sms, w, khz);
    run<4, 0>(out, sms, w, khz);
    run<8, 0>(out, sms, w, khz);
    run<16, 0>(out, sms, w, khz);
    run<24, 0>(out, sms, w, khz);
    run<32, 0>(out, sms, w, khz);
    run<48, 0>(out, sms, w, khz);
    run<8, 1>(out, sms, w, khz);
    run<16, 1>(out, sms, w, khz);
    run<32, 1>(out, sms, w, khz);
    run<8, 2>(out, sms, w, khz);
    run<16, 2>(out, sms, w, khz);
    run<32, 2>(out, sms, w, khz);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
