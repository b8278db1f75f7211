// ldg_rate.cu -- per-SM throughput of the three ways PCM can reach an SM, in the access pattern of
// the tensor-core FIR's converters (bring-up tool): every CTA streams 64-frame stages (256 bytes per
// stream) of its own 64 stereo streams, rows `row_bytes` apart, one CTA per SM, all SMs at once.
//   mode 0: LDG.128 into registers, two register sets per thread in flight (what the kernels do)
//   mode 1: cp.async 16 B (LDGSTS, L1 bypass) into a shared-memory ring
//   mode 2: cp.async.bulk, one 256-byte copy per stream and stage, into a shared-memory ring
// The dynamic shared memory asked for sets the L1 carve-out (what is left of 256 KB).
// Build: make -C node_speex_resampler_b200/csrc ldgrate ; run on a B200: ./ldg_rate
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma_ptx.cuh"

using namespace spxb::ptx;

struct Args {
  const uint8_t *src;
  uint32_t row_bytes;   // bytes between consecutive streams
  uint32_t row_stages;  // 64-frame stages per row before the CTA wraps to the row start
  uint32_t iters;       // stages each CTA streams
  uint32_t warps;       // loading warps
  uint32_t depth;       // ring depth (modes 1, 2), stages
  uint32_t span;        // mode 0: stages per register set (2 sets per thread: 32 KB x span in flight per SM)
  unsigned long long *cycles;
  uint32_t *sink;
};

constexpr uint32_t kStageBytes = 64 * 256;  // 64 streams x 64 stereo frames

__global__ void __launch_bounds__(512, 1) ldg_kernel(const Args p) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= p.warps) return;
  // a stage = 1024 items of 16 bytes; lanes walk along a stream's 256 bytes (16 items), two streams per
  // warp instruction; a thread's items are 2 * warps streams apart
  const uint32_t ips = 32 / p.warps;        // per thread per stage (warps = 4: 8, warps = 8: 4)
  const uint32_t items = ips * p.span;      // per register set
  const uint8_t *base = p.src + static_cast<size_t>(blockIdx.x) * 64 * p.row_bytes +
                        static_cast<size_t>(warp * 2 + lane / 16) * p.row_bytes + (lane % 16) * 16;
  const size_t item_stride = static_cast<size_t>(2 * p.warps) * p.row_bytes;
  uint4 r0[8], r1[8];
  uint32_t acc = 0;
  auto fetch = [&](uint4 (&r)[8], uint32_t it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (static_cast<uint32_t>(i) < items) {
        const uint8_t *s = base + static_cast<size_t>((it * p.span + i / ips) % p.row_stages) * 256;
        r[i] = __ldg(reinterpret_cast<const uint4 *>(s + (i % ips) * item_stride));
      }
  };
  auto eat = [&](const uint4 (&r)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (static_cast<uint32_t>(i) < items) acc ^= r[i].x ^ r[i].y ^ r[i].z ^ r[i].w;
  };
  const long long t0 = clock64();
  fetch(r0, 0);
  fetch(r1, 1);
  for (uint32_t it = 0; it < p.iters / p.span; it += 2) {
    eat(r0);
    fetch(r0, it + 2);
    eat(r1);
    fetch(r1, it + 3);
  }
  eat(r0);
  eat(r1);
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
  if (acc == 0x12345u) p.sink[0] = acc;
}

__global__ void __launch_bounds__(512, 1) ldgsts_kernel(const Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= p.warps) return;
  const uint32_t items = 32 / p.warps;
  const uint8_t *base = p.src + static_cast<size_t>(blockIdx.x) * 64 * p.row_bytes +
                        static_cast<size_t>(warp * 2 + lane / 16) * p.row_bytes + (lane % 16) * 16;
  const size_t item_stride = static_cast<size_t>(2 * p.warps) * p.row_bytes;
  const uint32_t my_off = (warp * 2 + lane / 16) * 256 + (lane % 16) * 16;
  uint32_t acc = 0;
  auto issue = [&](uint32_t it) {
    if (it < p.iters) {
      const uint8_t *s = base + static_cast<size_t>(it % p.row_stages) * 256;
      const uint32_t dst = smem_u32(smem + (it % p.depth) * kStageBytes + my_off);
      for (uint32_t i = 0; i < items; ++i)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 2 * p.warps * 256), "l"(s + i * item_stride)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const long long t0 = clock64();
  for (uint32_t d = 0; d + 1 < p.depth; ++d) issue(d);
  for (uint32_t it = 0; it < p.iters; ++it) {
    issue(it + p.depth - 1);
    // the oldest group (stage `it`) has landed when at most depth - 1 groups are pending
    if (p.depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else if (p.depth == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else if (p.depth == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
    else if (p.depth == 6) asm volatile("cp.async.wait_group 5;" ::: "memory");
    else asm volatile("cp.async.wait_group 7;" ::: "memory");
    const uint4 v = *reinterpret_cast<const uint4 *>(smem + (it % p.depth) * kStageBytes + my_off);  // own data only
    acc ^= v.x ^ v.w;
  }
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
  if (acc == 0x12345u) p.sink[0] = acc;
}

__global__ void __launch_bounds__(512, 1) bulk_kernel(const Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < p.depth; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (warp != 0) return;
  const uint8_t *base = p.src + static_cast<size_t>(blockIdx.x) * 64 * p.row_bytes;
  const long long t0 = clock64();
  for (uint32_t it = 0; it < p.iters + p.depth; ++it) {
    const uint32_t slot = it % p.depth, par = ((it / p.depth) & 1u) ^ 1u;
    if (it >= p.depth) mbar_wait(&bar[slot], par);
    if (it < p.iters) {
      // span = 256-byte units per copy: a "stage" here is 64 rows x 256*span bytes (span stages of the FIR at once)
      const uint32_t seg = 256 * p.span;
      if (lane == 0) mbar_arrive_expect_tx(&bar[slot], 64 * seg);
      __syncwarp();
      const uint8_t *s = base + static_cast<size_t>((it * p.span) % p.row_stages) * 256;
      for (uint32_t r = lane; r < 64; r += 32)
        bulk_g2s(smem + slot * 64 * seg + r * seg, s + static_cast<size_t>(r) * p.row_bytes, seg, &bar[slot]);
    }
  }
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
}

// mode 3: one 2-D tensor-map box per stage (64 rows x 256 bytes), or `span` boxes of 256/span bytes
__global__ void __launch_bounds__(512, 1) tma2d_kernel(const Args p, const __grid_constant__ CUtensorMap map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < p.depth; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const long long t0 = clock64();
  const uint32_t w = 256 / p.span;
  for (uint32_t it = 0; it < p.iters + p.depth; ++it) {
    const uint32_t slot = it % p.depth, par = ((it / p.depth) & 1u) ^ 1u;
    if (it >= p.depth) mbar_wait(&bar[slot], par);
    if (it < p.iters) {
      mbar_arrive_expect_tx(&bar[slot], kStageBytes);
      for (uint32_t k = 0; k < p.span; ++k)
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
                "r"(smem_u32(smem + slot * kStageBytes + k * 64 * w)),
            "l"(&map), "r"(static_cast<int>((it % p.row_stages) * 256 + k * w)), "r"(static_cast<int>(blockIdx.x * 64)),
            "r"(smem_u32(&bar[slot]))
            : "memory");
    }
  }
  p.cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (argc > 2) sms = atoi(argv[2]);  // CTAs (one per SM)
  const uint32_t row_bytes = argc > 3 ? static_cast<uint32_t>(atoi(argv[3])) : 7680;  // C5: 1920 stereo frames
  const uint32_t row_stages = argc > 1 ? static_cast<uint32_t>(atoi(argv[1])) : row_bytes / 256;  // stages before a CTA wraps
  printf("CTAs %d, rows %u bytes apart, footprint %.1f MB\n", sms, row_bytes, sms * 64.0 * row_stages * 256 / 1e6);
  const size_t bytes = static_cast<size_t>(sms) * 64 * row_bytes;
  uint8_t *src;
  unsigned long long *cyc;
  uint32_t *sink;
  cudaMalloc(&src, bytes);
  cudaMemset(src, 1, bytes);
  cudaMalloc(&cyc, sms * sizeof(unsigned long long));
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(ldg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(ldgsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<unsigned long long> h(sms);
  auto report = [&](const char *what, uint32_t iters) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", what, cudaGetErrorString(e));
      return;
    }
    cudaMemcpy(h.data(), cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemset(cyc, 0, sms * sizeof(unsigned long long));
    double sum = 0, mx = 0;
    for (int i = 0; i < sms; ++i) {
      sum += static_cast<double>(h[i]);
      mx = h[i] > mx ? static_cast<double>(h[i]) : mx;
    }
    printf("%-58s %7.1f cycles/stage  %5.1f B/clk/SM (slowest SM %5.1f)\n", what, sum / sms / iters,
           static_cast<double>(kStageBytes) * iters * sms / sum, static_cast<double>(kStageBytes) * iters / mx);
  };
  const uint32_t iters = 600;  // 20 passes over the CTA's rows: L2 hits after the first
  char what[128];
  for (uint32_t smem_kb : {200u}) {
    for (uint32_t warps : {4u, 8u, 16u})
      for (uint32_t span : {1u, 2u, 4u}) {
        if (32 * span / warps > 8 || 32 * span / warps == 0) continue;
        Args p{src, row_bytes, row_stages, iters, warps, 2, span, cyc, sink};
        for (int rep = 0; rep < 2; ++rep) ldg_kernel<<<sms, 512, smem_kb * 1024>>>(p);
        snprintf(what, sizeof(what), "LDG.128  %2u warps x 2 sets (%3u KB in flight) smem %3u KB", warps, 32u * span, smem_kb);
        report(what, iters);
      }
  }
  for (uint32_t depth : {2u, 4u}) {
    for (uint32_t warps : {4u}) {
      Args p{src, row_bytes, row_stages, iters, warps, depth, 1, cyc, sink};
      for (int rep = 0; rep < 2; ++rep) ldgsts_kernel<<<sms, 512, 200 * 1024>>>(p);
      snprintf(what, sizeof(what), "cp.async 16 B  %u warps ring of %u stages (smem 200 KB)", warps, depth);
      report(what, iters);
    }
  }
  for (uint32_t span : {1u, 2u, 4u, 8u})
    for (uint32_t depth : {2u, 4u}) {
      if (depth * span * kStageBytes > 200 * 1024 || row_stages % span) continue;
      Args p{src, row_bytes, row_stages, iters / span, 1, depth, span, cyc, sink};
      for (int rep = 0; rep < 2; ++rep) bulk_kernel<<<sms, 512, 200 * 1024>>>(p);
      snprintf(what, sizeof(what), "cp.async.bulk %4u B x 64  ring of %u (cycles per 16 KB)", 256 * span, depth);
      report(what, iters);
    }
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  cudaFuncSetAttribute(tma2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (uint32_t span : {1u, 4u})
    for (uint32_t depth : {1u, 2u, 4u, 8u}) {
      CUtensorMap map;
      cuuint64_t gdim[2] = {row_bytes, static_cast<cuuint64_t>(sms) * 64};
      cuuint64_t gstride[1] = {row_bytes};
      cuuint32_t box[2] = {256 / span, 64};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = reinterpret_cast<EncodeFn>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, gdim, gstride, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        printf("encode failed: %d\n", static_cast<int>(r));
        continue;
      }
      Args p{src, row_bytes, row_stages, iters, 1, depth, span, cyc, sink};
      for (int rep = 0; rep < 2; ++rep) tma2d_kernel<<<sms, 512, 200 * 1024>>>(p, map);
      snprintf(what, sizeof(what), "2-D tensor map, %u box(es) of %3u B x 64 rows, ring of %u", span, 256 / span, depth);
      report(what, iters);
    }
  return 0;
}
