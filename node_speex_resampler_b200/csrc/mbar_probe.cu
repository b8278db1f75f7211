// mbar_probe.cu -- what does initialising mbarriers cost? (bring-up tool; the tensor kernels' prologues
// spend ~1700 cycles in it). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_probe mbar_probe.cu
#include <cuda_runtime.h>

#include <cstdio>

#include "umma_ptx.cuh"

using namespace spxb::ptx;

__global__ void probe(unsigned long long *out, uint32_t n, uint32_t count_arg) {
  __shared__ uint64_t bars[64];
  const uint32_t lane = threadIdx.x & 31;
  long long t0, t1;
  // (a) one lane, n barriers one after the other, run-time count
  __syncthreads();
  t0 = clock64();
  if (threadIdx.x == 0)
    for (uint32_t i = 0; i < n; ++i) mbar_init(&bars[i], count_arg);
  __syncwarp();
  t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  // (b) one lane, compile-time counts
  __syncthreads();
  t0 = clock64();
  if (threadIdx.x == 0) {
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
  }
  __syncwarp();
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = t1 - t0;
  // (c) 32 lanes, one barrier each, same instruction, run-time count per lane
  __syncthreads();
  t0 = clock64();
  if (threadIdx.x < 32) mbar_init(&bars[lane], count_arg + (lane & 3));
  __syncwarp();
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = t1 - t0;
  // (d) the fence
  __syncthreads();
  t0 = clock64();
  if (threadIdx.x < 32) fence_mbar_init();
  __syncwarp();
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = t1 - t0;
  // (e) divergent: each lane its own branch (8-way)
  __syncthreads();
  t0 = clock64();
  if (threadIdx.x < 32) {
    const uint32_t i = lane;
    if (i < 4) mbar_init(&bars[i], 4);
    else if (i < 8) mbar_init(&bars[i], 1);
    else if (i < 12) mbar_init(&bars[i], 1);
    else if (i < 20) mbar_init(&bars[i], 1);
    else if (i == 20) mbar_init(&bars[i], 1);
    else if (i <= 22) mbar_init(&bars[i], 1);
    else if (i <= 24) mbar_init(&bars[i], count_arg);
    else mbar_init(&bars[i], 1);
  }
  __syncwarp();
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = t1 - t0;
  // (f) clock overhead
  t0 = clock64();
  t1 = clock64();
  if (threadIdx.x == 0) out[5] = t1 - t0;
}

int main() {
  unsigned long long *d, h[6];
  cudaMalloc(&d, sizeof(h));
  for (int rep = 0; rep < 3; ++rep) {
    probe<<<1, 128>>>(d, 13, 4);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("rep %d: 13 inits by one lane %llu cycles | 16 unrolled constant-count inits %llu | 32 lanes one init each %llu | fence.mbarrier_init %llu | 8-way divergent chain %llu | clock pair %llu\n",
           rep, h[0], h[1], h[2], h[3], h[4], h[5]);
  }
  return 0;
}
