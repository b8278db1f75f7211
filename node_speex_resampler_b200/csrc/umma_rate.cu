// umma_rate.cu -- tcgen05.mma issue-rate microbenchmark (bring-up tool, run on a B200):
// cycles per MMA for kind::i8 / kind::f16, N, and the shared-memory operand layout
// (no swizzle "interleaved" vs 32/64/128-byte swizzle, K-major). Decides the operand layout of
// kernels_umma.cu. Build: make -C node_speex_resampler_b200/csrc rate ; run: ./umma_rate
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma_ptx.cuh"

using namespace spxb::ptx;

struct RateArgs {
  int i8;         // 1: kind::i8, 0: kind::f16
  int n;          // MMA N
  int layout;     // descriptor layout type field (0 none, 6 sw32, 4 sw64, 2 sw128)
  uint32_t lbo, sbo;
  uint32_t kadv;  // bytes between the K slices of consecutive MMAs (cycled over `kslices`)
  int kslices;
  int reps;
  int pairs;      // 1: alternate two A operands (hi / lo planes) against the same B
  unsigned long long *cycles;  // [grid]
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(layout & 7) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) rate_kernel(const RateArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // A0 [0,32K) A1 [32K,64K) B [64K,128K)
  for (uint32_t i = tid; i < 128 * 1024 / 16; i += blockDim.x)
    reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x01020304u * (i & 3), 0x01010101u, i * 2654435761u, 0x7f80ff01u);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(smem), a1 = a0 + 32 * 1024, b = a0 + 64 * 1024;
    const uint32_t idesc = p.i8 ? umma_idesc_i8(128, p.n, true, true) : umma_idesc_f16(128, p.n);
    // warm-up
    for (int r = 0; r < 8; ++r) {
      const uint64_t da = make_desc(a0, p.lbo, p.sbo, p.layout), db = make_desc(b, p.lbo, p.sbo, p.layout);
      if (p.i8) umma_i8(tmem, da, db, idesc, r > 0);
      else umma_f16(tmem, da, db, idesc, r > 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    // descriptors precomputed: the timed loop is MMA issue only (a single thread's scalar
    // work per MMA must stay far below the MMA's own N/2 cycles)
    uint64_t da[2][4], db[4];
    for (int k = 0; k < 4; ++k) {
      const uint32_t ko = (k % p.kslices) * p.kadv;
      db[k] = make_desc(b + ko, p.lbo, p.sbo, p.layout);
      da[0][k] = make_desc(a0 + ko, p.lbo, p.sbo, p.layout);
      da[1][k] = make_desc((p.pairs ? a1 : a0) + ko, p.lbo, p.sbo, p.layout);
    }
    const uint32_t col1 = p.pairs ? 128u : 0u;
    const long long t0 = clock64();
    if (p.i8) {
      for (int r = 0; r < p.reps; r += 4) {
        umma_i8(tmem, da[0][0], db[0], idesc, 1);
        umma_i8(tmem + col1, da[1][1], db[1], idesc, 1);
        umma_i8(tmem, da[0][2], db[2], idesc, 1);
        umma_i8(tmem + col1, da[1][3], db[3], idesc, 1);
      }
    } else {
      for (int r = 0; r < p.reps; r += 4) {
        umma_f16(tmem, da[0][0], db[0], idesc, 1);
        umma_f16(tmem + col1, da[1][1], db[1], idesc, 1);
        umma_f16(tmem, da[0][2], db[2], idesc, 1);
        umma_f16(tmem + col1, da[1][3], db[3], idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t1 = clock64();
    p.cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// The FIR kernel's own issue pattern: per K step four MMAs (hi plane x [0,3nt), lo plane x the
// same B into columns [nt,4nt)), N = 3nt split in two pieces, operands walking through a ring of
// stages exactly as kernels_umma.cu lays them out. No loads, no converters: the tensor pipe alone.
struct FirArgs {
  uint32_t nt, a_lbo, b_lbo, piece0;  // piece0: columns of the first piece (rest in the second)
  uint32_t stages, ksteps;
  unsigned long long *cycles;
};

__global__ void __launch_bounds__(128, 1) fir_rate_kernel(const FirArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < 200 * 1024 / 16; i += blockDim.x)
    reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x01020304u * (i & 3), 0x01010101u, i * 2654435761u, 0x7f80ff01u);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t nt = p.nt, n3 = 3 * nt, np0 = p.piece0, np1 = n3 - np0;
    const uint32_t a_chunk = p.a_lbo, x_plane = 4 * a_chunk, x_stage = 2 * x_plane;
    const uint32_t tap_chunk = p.b_lbo, stage_bytes = x_stage + 4 * tap_chunk;
    const uint32_t id_hi0 = umma_idesc_i8(128, np0, true, true), id_hi1 = umma_idesc_i8(128, np1 ? np1 : 16, true, true);
    const uint32_t id_lo0 = umma_idesc_i8(128, np0, false, true), id_lo1 = umma_idesc_i8(128, np1 ? np1 : 16, false, true);
    const uint64_t a_base = make_desc(smem_u32(smem), a_chunk, 128, 0);
    const uint64_t b_base = make_desc(smem_u32(smem) + x_stage, tap_chunk, 128, 0);
    const uint32_t st16 = stage_bytes >> 4, a_ks16 = (2 * a_chunk) >> 4, a_lo16 = x_plane >> 4, b_ks16 = (2 * tap_chunk) >> 4;
    long long t0 = 0;
    for (uint32_t k = 0; k < p.ksteps + 4; ++k) {
      if (k == 4) {
        // warm-up done
        if ((tid & 31) == 0) {
          umma_commit(&bar);
          mbar_wait(&bar, 0);
        }
        __syncwarp();
        t0 = clock64();
      }
      const uint32_t slot = (k / 2) % p.stages, ks = k & 1;
      const uint64_t a_hi = a_base + slot * st16 + ks * a_ks16, a_lo = a_hi + a_lo16, b = b_base + slot * st16 + ks * b_ks16;
      uint32_t pred;
      asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(pred));
      if (pred) {
        umma_i8(tmem, a_hi, b, id_hi0, 1u);
        if (np1) umma_i8(tmem + np0, a_hi, b + np0, id_hi1, 1u);
        umma_i8(tmem + nt, a_lo, b, id_lo0, 1u);
        if (np1) umma_i8(tmem + nt + np0, a_lo, b + np0, id_lo1, 1u);
      }
      __syncwarp();
    }
    if ((tid & 31) == 0) {
      umma_commit(&bar);
      mbar_wait(&bar, 1);
      p.cycles[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static void run_fir(const char *name, FirArgs p, int grid) {
  unsigned long long *d;
  cudaMalloc(&d, grid * sizeof(unsigned long long));
  p.cycles = d;
  cudaFuncSetAttribute(fir_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  fir_rate_kernel<<<grid, 128, 204 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-60s CUDA error %s\n", name, cudaGetErrorString(e));
    exit(2);
  }
  std::vector<unsigned long long> h(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  double mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  printf("%-60s grid %3d: %7.1f cyc per K step (MMA floor 2 x 3nt/2 = %u)\n", name, grid, mx / p.ksteps, 3 * p.nt);
  fflush(stdout);
}

static double run(const char *name, RateArgs p, int grid) {
  unsigned long long *d;
  cudaMalloc(&d, grid * sizeof(unsigned long long));
  p.cycles = d;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  rate_kernel<<<grid, 128, 130 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-52s CUDA error %s\n", name, cudaGetErrorString(e));
    exit(2);
  }
  std::vector<unsigned long long> h(grid);
  cudaMemcpy(h.data(), d, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  double mx = 0, mn = 1e30;
  for (auto v : h) {
    mx = v > mx ? v : mx;
    mn = v < mn ? v : mn;
  }
  const double per = mx / p.reps;
  printf("%-52s grid %3d: %7.1f cyc/MMA (min-CTA %7.1f)  floor N/2 = %5.1f  -> %4.0f%% of floor rate\n", name, grid,
         per, mn / p.reps, p.n / 2.0, 100.0 * (p.n / 2.0) / per);
  fflush(stdout);
  return per;
}

int main(int argc, char **argv) {
  const int R = 512;
  if (argc > 1) {
    // FIR issue pattern (kernels_umma.cu): tile widths, chunk strides, piece splits
    for (int grid : {1, 148}) {
      run_fir("fir nt=112 a_lbo=2080 pieces 256+80", FirArgs{112, 2080, 3 * 112 * 16, 256, 4, 512, nullptr}, grid);
      run_fir("fir nt=112 a_lbo=2048 pieces 256+80", FirArgs{112, 2048, 3 * 112 * 16, 256, 4, 512, nullptr}, grid);
      run_fir("fir nt=112 a_lbo=2080 pieces 176+160", FirArgs{112, 2080, 3 * 112 * 16, 176, 4, 512, nullptr}, grid);
      run_fir("fir nt=112 a_lbo=2080 b_lbo+32 pieces 176+160", FirArgs{112, 2080, 3 * 112 * 16 + 32, 176, 4, 512, nullptr}, grid);
      run_fir("fir nt=80  a_lbo=2080 one piece 240", FirArgs{80, 2080, 3 * 80 * 16, 240, 6, 512, nullptr}, grid);
      run_fir("fir nt=80  a_lbo=2048 one piece 240", FirArgs{80, 2048, 3 * 80 * 16, 240, 6, 512, nullptr}, grid);
      run_fir("fir nt=64  a_lbo=2080 one piece 192", FirArgs{64, 2080, 3 * 64 * 16, 192, 6, 512, nullptr}, grid);
      run_fir("fir nt=128 a_lbo=2080 pieces 256+128", FirArgs{128, 2080, 3 * 128 * 16, 256, 3, 512, nullptr}, grid);
      run_fir("fir nt=128 a_lbo=2080 pieces 192+192", FirArgs{128, 2080, 3 * 128 * 16, 192, 3, 512, nullptr}, grid);
      run_fir("fir nt=128 a_lbo=2048 pieces 192+192", FirArgs{128, 2048, 3 * 128 * 16, 192, 3, 512, nullptr}, grid);
    }
    return 0;
  }
  for (int grid : {1, 148}) {
    for (int i8 : {1, 0}) {
      for (int n : {64, 128, 256}) {
        char nm[128];
        // no swizzle: chunk-major [K chunk][row][16 B]: LBO = rows*16 (use 256 rows -> 4096), SBO = 128
        snprintf(nm, sizeof nm, "%s N=%3d none   lbo=4096 sbo=128", i8 ? "i8 " : "f16", n);
        run(nm, RateArgs{i8, n, 0, 4096, 128, 8192, 4, R, 0, nullptr}, grid);
        // no swizzle, row-group-major: [8-row group][K chunk][8 rows][16 B]: LBO = 128, SBO = 256
        snprintf(nm, sizeof nm, "%s N=%3d none   lbo=128 sbo=256", i8 ? "i8 " : "f16", n);
        run(nm, RateArgs{i8, n, 0, 128, 256, 0, 1, R, 0, nullptr}, grid);
        snprintf(nm, sizeof nm, "%s N=%3d sw32   sbo=256", i8 ? "i8 " : "f16", n);
        run(nm, RateArgs{i8, n, 6, 16, 256, 8192, 4, R, 0, nullptr}, grid);
        snprintf(nm, sizeof nm, "%s N=%3d sw64   sbo=512 (2 K slices)", i8 ? "i8 " : "f16", n);
        run(nm, RateArgs{i8, n, 4, 16, 512, 32, 2, R, 0, nullptr}, grid);
        snprintf(nm, sizeof nm, "%s N=%3d sw128  sbo=1024 (4 K slices)", i8 ? "i8 " : "f16", n);
        run(nm, RateArgs{i8, n, 2, 16, 1024, 32, 4, R, 0, nullptr}, grid);
      }
      char nm[128];
      snprintf(nm, sizeof nm, "%s N=256 sw128 alternating A planes", i8 ? "i8 " : "f16");
      run(nm, RateArgs{i8, 256, 2, 16, 1024, 32, 4, R, 1, nullptr}, grid);
      snprintf(nm, sizeof nm, "%s N=256 none  alternating A planes", i8 ? "i8 " : "f16");
      run(nm, RateArgs{i8, 256, 0, 4096, 128, 8192, 4, R, 1, nullptr}, grid);
      snprintf(nm, sizeof nm, "%s N=80  none", i8 ? "i8 " : "f16");
      run(nm, RateArgs{i8, 80, 0, 4096, 128, 8192, 4, R, 0, nullptr}, grid);
      snprintf(nm, sizeof nm, "%s N=16  none", i8 ? "i8 " : "f16");
      run(nm, RateArgs{i8, 16, 0, 4096, 128, 8192, 4, R, 0, nullptr}, grid);
    }
  }
  return 0;
}
