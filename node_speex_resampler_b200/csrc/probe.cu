// probe.cu -- FP32 peak probe: the measured denominator of the fp32-FMA roofline bench.py
// reports (MEASURED_PEAKS.json carries no fp32 figure). 8 independent FMA chains per thread,
// register resident, 8 blocks of 256 threads per SM.
#include <cuda_runtime.h>

#include "launch.h"

namespace spxb {
namespace {
__global__ void __launch_bounds__(256) ffma_probe_kernel(float *sink, int iters, float seed) {
  float x0 = seed + threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  const float m = 0.999999f, c = 1e-7f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, m, c); x1 = fmaf(x1, m, c); x2 = fmaf(x2, m, c); x3 = fmaf(x3, m, c);
      x4 = fmaf(x4, m, c); x5 = fmaf(x5, m, c); x6 = fmaf(x6, m, c); x7 = fmaf(x7, m, c);
    }
  }
  const float r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (r == 123.456f) sink[0] = r;  // keep the chains alive
}
}  // namespace

double measure_fp32_peak_flops(int iters) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float *sink = nullptr;
  if (cudaMalloc(&sink, 4) != cudaSuccess) return 0.0;
  const int blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  ffma_probe_kernel<<<blocks, threads>>>(sink, 64, 1.f);  // warm-up
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  ffma_probe_kernel<<<blocks, threads>>>(sink, iters, 1.f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (ms <= 0.f) return 0.0;
  const double fmas = static_cast<double>(blocks) * threads * static_cast<double>(iters) * 16.0 * 8.0;
  return 2.0 * fmas / (ms * 1e-3);
}
}  // namespace spxb

extern "C" __attribute__((visibility("default"))) double spxb_measure_fp32_peak(int iters) {
  return spxb::measure_fp32_peak_flops(iters > 0 ? iters : 4096);
}
