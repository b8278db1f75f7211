// tma_rate.cu -- per-SM global->shared copy throughput of the async copy engines (bring-up tool):
// 1-D cp.async.bulk (UBLKCP) vs 2-D tensor-map TMA (UTMALDG) for the tap-stage stream of
// kernels_umma.cu (one ~21 KB copy per stage into a ring of 6 slots, data resident in L2).
// Build: make -C node_speex_resampler_b200/csrc tmarate ; run on a B200: ./tma_rate
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma_ptx.cuh"

using namespace spxb::ptx;

struct Args {
  const uint8_t *src;
  uint32_t stage_bytes, stages, iters, mode;  // mode 0: 1-D bulk, 1: 2-D tensor map, 2: 1-D bulk split in 4
  uint32_t tile_stride_rows;                  // rows of 256 B between the tiles of consecutive CTAs
  unsigned long long *cycles;
};

__device__ __forceinline__ void tma_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(128, 1) tma_kernel(const Args p, const __grid_constant__ CUtensorMap map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < p.stages; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t rows = p.stage_bytes / 256;
    const uint32_t tile_row0 = (blockIdx.x % 8) * p.tile_stride_rows;  // 8 distinct tiles, like the FIR grid
    const long long t0 = clock64();
    for (uint32_t it = 0; it < p.iters + p.stages; ++it) {
      const uint32_t slot = it % p.stages, par = ((it / p.stages) & 1u) ^ 1u;
      if (it >= p.stages) mbar_wait(&bar[slot], par);  // previous copy into this slot has landed
      if (it < p.iters) {
        uint8_t *dst = smem + slot * p.stage_bytes;
        const uint32_t row = tile_row0 + (it % 13) * rows;
        mbar_arrive_expect_tx(&bar[slot], p.stage_bytes);
        if (p.mode == 0) {
          bulk_g2s(dst, p.src + static_cast<size_t>(row) * 256, p.stage_bytes, &bar[slot]);
        } else if (p.mode == 2) {
          const uint32_t q = p.stage_bytes / 4;
          for (int k = 0; k < 4; ++k) bulk_g2s(dst + k * q, p.src + static_cast<size_t>(row) * 256 + k * q, q, &bar[slot]);
        } else {
          tma_2d(dst, &map, 0, static_cast<int>(row), &bar[slot]);
        }
      }
    }
    p.cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
  }
}

// The FIR kernel's three-party stage handshake without any math: producer lane (bulk copy, waits
// `empty`), `conv` converter threads (wait `empty`, optional proxy fence, arrive on `full`), consumer
// lane (waits `full`, releases the slot with tcgen05.commit or a plain arrive).
struct HsArgs {
  const uint8_t *src;
  uint32_t stage_bytes, stages, iters;
  uint32_t conv_threads;   // 0 or 256
  uint32_t fence;          // converters execute fence.proxy.async
  uint32_t commit;         // 1: release via tcgen05.commit, 0: mbarrier.arrive
  unsigned long long *cycles;
};

__global__ void __launch_bounds__(352, 1) hs_kernel(const HsArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[8], empty_bar[8];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (uint32_t s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], p.conv_threads + 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(&tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const long long t0 = clock64();
  if (warp < 8) {
    if (static_cast<uint32_t>(tid) < p.conv_threads) {
      for (uint32_t it = 0; it < p.iters; ++it) {
        const uint32_t slot = it % p.stages, par = (it / p.stages) & 1u;
        mbar_wait(&empty_bar[slot], par ^ 1u);
        if (p.fence) fence_proxy_async_smem();
        mbar_arrive(&full_bar[slot]);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      for (uint32_t it = 0; it < p.iters; ++it) {
        const uint32_t slot = it % p.stages, par = (it / p.stages) & 1u;
        mbar_wait(&empty_bar[slot], par ^ 1u);
        mbar_arrive_expect_tx(&full_bar[slot], p.stage_bytes);
        bulk_g2s(smem + slot * p.stage_bytes, p.src + static_cast<size_t>((blockIdx.x % 8) * 13 + it % 13) * p.stage_bytes,
                 p.stage_bytes, &full_bar[slot]);
      }
    }
  } else if (warp == 9) {
    for (uint32_t it = 0; it < p.iters; ++it) {
      const uint32_t slot = it % p.stages, par = (it / p.stages) & 1u;
      mbar_wait(&full_bar[slot], par);
      tc_fence_after_sync();
      if (lane == 0) {
        if (p.commit) umma_commit(&empty_bar[slot]);
        else mbar_arrive(&empty_bar[slot]);
      }
      __syncwarp();
    }
    if (lane == 0) p.cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_slot, 32);
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const uint32_t nt = 112, stage_bytes = 4 * 3 * nt * 16, stages = 6, iters = 2000;
  const uint32_t rows_per_stage = stage_bytes / 256, rows_per_tile = rows_per_stage * 13;
  const size_t total_rows = static_cast<size_t>(rows_per_tile) * 8 + rows_per_stage;
  uint8_t *d;
  cudaMalloc(&d, total_rows * 256);
  cudaMemset(d, 1, total_rows * 256);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  CUtensorMap map;
  cuuint64_t gdim[2] = {256, total_rows};
  cuuint64_t gstride[1] = {256};
  cuuint32_t box[2] = {256, rows_per_stage};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeFn>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstride, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d (rows per stage %u, stage %u B)\n", static_cast<int>(r), rows_per_stage, stage_bytes);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int same_tile : {0, 1})
  for (int grid : {1, 16, 148}) {
    for (uint32_t mode : {0u, 1u}) {
      unsigned long long *c;
      cudaMalloc(&c, grid * 8);
      Args a{d, stage_bytes, stages, iters, mode, same_tile ? 0u : rows_per_tile, c};
      tma_kernel<<<grid, 128, stages * stage_bytes>>>(a, map);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("mode %u grid %d: CUDA error %s\n", mode, grid, cudaGetErrorString(e));
        return 2;
      }
      std::vector<unsigned long long> h(grid);
      cudaMemcpy(h.data(), c, grid * 8, cudaMemcpyDeviceToHost);
      double mx = 0;
      for (auto v : h) mx = v > mx ? v : mx;
      printf("%s %-26s grid %3d: %7.1f cycles per %u-byte stage = %5.1f B/clk/SM\n",
             same_tile ? "[all CTAs read ONE tile]" : "[8 tiles]", mode == 0 ? "1-D bulk" : mode == 2 ? "1-D bulk, 4 pieces" : "2-D tensor map", grid, mx / iters, stage_bytes,
             stage_bytes / (mx / iters));
      cudaFree(c);
    }
  }
  for (uint32_t conv : {0u, 256u}) {
    for (uint32_t fence : {0u, 1u}) {
      for (uint32_t commit : {0u, 1u}) {
        if (conv == 0 && fence) continue;
        unsigned long long *c;
        const int grid = 148;
        cudaMalloc(&c, grid * 8);
        HsArgs a{d, stage_bytes, stages, iters, conv, fence, commit, c};
        cudaFuncSetAttribute(hs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        hs_kernel<<<grid, 352, stages * stage_bytes>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("handshake: CUDA error %s\n", cudaGetErrorString(e));
          return 2;
        }
        std::vector<unsigned long long> h(grid);
        cudaMemcpy(h.data(), c, grid * 8, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (auto v : h) mx = v > mx ? v : mx;
        printf("handshake conv=%3u fence=%u release=%s: %7.1f cycles per stage\n", conv, fence,
               commit ? "tcgen05.commit" : "mbarrier.arrive", mx / iters);
        cudaFree(c);
      }
    }
  }
  return 0;
}
