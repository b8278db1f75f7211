// xchg_probe.cu -- can a few SMs move a C3-sized step over PCIe as fast as the copy engines do,
// without the per-operation cost of DMA + stream events? One kernel pulls
// `pull_bytes` from mapped pinned host memory into HBM and pushes `push_bytes` from HBM to pinned
// host memory at the same time; timed back to back against cudaMemcpyAsync on two streams.
// Build: make xchg    Run on a B200: ./xchg_probe
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e__ = (x);                                                             \
    if (e__ != cudaSuccess) {                                                          \
      printf("%s: %s\n", #x, cudaGetErrorString(e__));                                 \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

// CTAs [0, pull_ctas) pull, the rest push; 16-byte accesses, UNROLL independent ones per thread
template <int UNROLL>
__global__ void __launch_bounds__(256) xchg_kernel(const uint4 *__restrict__ pull_src, uint4 *__restrict__ pull_dst,
                                                   size_t pull_n, const uint4 *__restrict__ push_src,
                                                   uint4 *__restrict__ push_dst, size_t push_n, int pull_ctas) {
  const bool pull = static_cast<int>(blockIdx.x) < pull_ctas;
  const uint4 *src = pull ? pull_src : push_src;
  uint4 *dst = pull ? pull_dst : push_dst;
  const size_t n = pull ? pull_n : push_n;
  const size_t ctas = pull ? pull_ctas : gridDim.x - pull_ctas;
  const size_t rank = pull ? blockIdx.x : blockIdx.x - pull_ctas;
  const size_t stride = ctas * blockDim.x;
  size_t i = rank * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = src[i + u * stride];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) dst[i + u * stride] = v[u];
  }
  for (; i < n; i += stride) dst[i] = src[i];
}

int main() {
  const size_t in_bytes = 1024ull * 882 * 2 * 2, out_bytes = 1024ull * 960 * 2 * 2;  // C3 step
  const int HR = 6, STEPS = 200;
  std::vector<void *> h_in(HR), h_out(HR);
  for (int k = 0; k < HR; ++k) {
    CK(cudaHostAlloc(&h_in[k], in_bytes, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h_out[k], out_bytes, cudaHostAllocDefault));
  }
  void *d_in[4], *d_out[4];
  for (int k = 0; k < 4; ++k) {
    CK(cudaMalloc(&d_in[k], in_bytes));
    CK(cudaMalloc(&d_out[k], out_bytes));
  }
  cudaStream_t s1, s2;
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto report = [&](const char *label, float ms) {
    printf("%-52s %7.1f us/step\n", label, ms * 1e3f / STEPS);
  };
  // DMA reference: both directions free-running on two streams (wall time over both)
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, s1));
    for (int k = 0; k < STEPS; ++k) {
      CK(cudaMemcpyAsync(d_in[k % 4], h_in[k % HR], in_bytes, cudaMemcpyHostToDevice, s1));
      CK(cudaMemcpyAsync(h_out[k % HR], d_out[k % 4], out_bytes, cudaMemcpyDeviceToHost, s2));
    }
    CK(cudaStreamSynchronize(s2));
    CK(cudaEventRecord(e1, s1));
    CK(cudaDeviceSynchronize());
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  report("DMA: H2D || D2H free-running (two streams)", ms);

  const int grids[] = {8, 16, 32, 64, 128};
  for (int mode = 0; mode < 3; ++mode) {  // 0 pull only, 1 push only, 2 both
    for (int g : grids) {
      const int pull_ctas = mode == 0 ? g : mode == 1 ? 0 : g / 2;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s1));
        for (int k = 0; k < STEPS; ++k)
          xchg_kernel<4><<<g, 256, 0, s1>>>(static_cast<const uint4 *>(h_in[k % HR]), static_cast<uint4 *>(d_in[k % 4]),
                                            mode == 1 ? 0 : in_bytes / 16, static_cast<const uint4 *>(d_out[k % 4]),
                                            static_cast<uint4 *>(h_out[k % HR]), mode == 0 ? 0 : out_bytes / 16,
                                            pull_ctas);
        CK(cudaEventRecord(e1, s1));
        CK(cudaDeviceSynchronize());
      }
      CK(cudaGetLastError());
      CK(cudaEventElapsedTime(&ms, e0, e1));
      char label[96];
      snprintf(label, sizeof label, "SM copy: %s, %d CTAs x 256 threads, 4 x 16 B in flight",
               mode == 0 ? "pull only" : mode == 1 ? "push only" : "pull || push", g);
      report(label, ms);
    }
  }
  // deeper unroll at the best-looking grid
  for (int g : {32, 64}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, s1));
      for (int k = 0; k < STEPS; ++k)
        xchg_kernel<8><<<g, 256, 0, s1>>>(static_cast<const uint4 *>(h_in[k % HR]), static_cast<uint4 *>(d_in[k % 4]),
                                          in_bytes / 16, static_cast<const uint4 *>(d_out[k % 4]),
                                          static_cast<uint4 *>(h_out[k % HR]), out_bytes / 16, g / 2);
      CK(cudaEventRecord(e1, s1));
      CK(cudaDeviceSynchronize());
    }
    CK(cudaEventElapsedTime(&ms, e0, e1));
    char label[96];
    snprintf(label, sizeof label, "SM copy: pull || push, %d CTAs, 8 x 16 B in flight", g);
    report(label, ms);
  }
  return 0;
}
