// atmem_probe.cu -- bring-up check: tcgen05.mma.kind::i8 with the A operand in TENSOR MEMORY, written there
// by tcgen05.st from registers (what a converter that bypasses shared memory for the byte planes needs).
// One CTA: warps 0-3 write A (128 rows x 32 bytes of K = 8 columns) into TMEM, B (N x 32 bytes) sits in shared
// memory in the kernels' canonical K-major layout, D = A x B is read back and compared with the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o atmem_probe atmem_probe.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma_ptx.cuh"

using namespace spxb::ptx;

struct Args {
  const int8_t *a;   // [128][32] row-major bytes (s8 or u8 bit patterns)
  const int8_t *b;   // [N][32] row-major s8
  int32_t *d;        // [128][N]
  int n;
  int a_signed;
  int pack;          // hypothesis for the byte order inside a TMEM column: 0 little-endian k = 4j + byte
  int *status;
};

__global__ void __launch_bounds__(160, 1) probe(const Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  // B: chunk c (16 bytes of K) of row n at c * N * 16 + n * 16
  for (int i = tid; i < p.n * 2; i += blockDim.x) {
    const int n = i % p.n, c = i / p.n;
    reinterpret_cast<uint4 *>(smem)[c * p.n + n] = *reinterpret_cast<const uint4 *>(p.b + n * 32 + c * 16);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t a_col = 256;  // A lives in columns [256, 264)
  if (warp < 4) {
    const int row = warp * 32 + lane;
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) {
      uint32_t w = 0;
      for (int e = 0; e < 4; ++e) {
        const int k = p.pack == 0 ? 4 * j + e : 4 * j + (3 - e);
        w |= static_cast<uint32_t>(static_cast<uint8_t>(p.a[row * 32 + k])) << (8 * e);
      }
      r[j] = w;
    }
    const uint32_t taddr = tmem + a_col + (static_cast<uint32_t>(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 4 && lane == 0) {
    const uint64_t db = umma_smem_desc(smem_u32(smem), p.n * 16, 128);
    const uint32_t idesc = umma_idesc_i8(128, p.n, p.a_signed, true);
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, q;\n\t}" ::"r"(tmem),
        "r"(tmem + a_col), "l"(db), "r"(idesc), "r"(0)
        : "memory");
    umma_commit(&bar);
  }
  if (warp < 4) {
    bool ok = false;
    for (int i = 0; i < (1 << 22); ++i)
      if (mbar_try_wait(&bar, 0)) {
        ok = true;
        break;
      }
    if (!ok && lane == 0) *p.status = 1;
    tc_fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < p.n; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + c0 + (static_cast<uint32_t>(warp * 32) << 16), v);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) p.d[row * p.n + c0 + j] = static_cast<int32_t>(v[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(slot, 512);
}

int main() {
  const int N = 64;
  std::vector<int8_t> a(128 * 32), b(N * 32);
  srand(7);
  for (auto &v : a) v = static_cast<int8_t>(rand() % 256 - 128);
  for (auto &v : b) v = static_cast<int8_t>(rand() % 256 - 128);
  int8_t *da, *dbb;
  int32_t *dd;
  int *st;
  cudaMalloc(&da, a.size());
  cudaMalloc(&dbb, b.size());
  cudaMalloc(&dd, 128 * N * 4);
  cudaMalloc(&st, 4);
  cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dbb, b.data(), b.size(), cudaMemcpyHostToDevice);
  for (int a_signed : {1, 0})
    for (int pack : {0, 1}) {
      cudaMemset(dd, 0xff, 128 * N * 4);
      cudaMemset(st, 0, 4);
      Args p{da, dbb, dd, N, a_signed, pack, st};
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      probe<<<1, 160, 64 * 1024>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("a_signed %d pack %d: CUDA error %s\n", a_signed, pack, cudaGetErrorString(e));
        return 2;
      }
      std::vector<int32_t> d(128 * N);
      int status = 0;
      cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(&status, st, 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          long ref = 0;
          for (int k = 0; k < 32; ++k) {
            const int av = a_signed ? a[m * 32 + k] : static_cast<uint8_t>(a[m * 32 + k]);
            ref += static_cast<long>(av) * b[n * 32 + k];
          }
          if (d[m * N + n] != ref) ++bad;
        }
      printf("A in TMEM, a %s, byte order hypothesis %d: %d of %d accumulators differ (status %d)%s\n", a_signed ? "s8" : "u8", pack, bad,
             128 * N, status, bad == 0 ? "  <-- exact" : "");
    }
  return 0;
}
