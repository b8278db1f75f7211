// umma_probe.cu -- standalone bring-up check of the tcgen05 building blocks used by
// kernels_umma.cu, against an exact integer reference computed on the host:
//   * K-major no-swizzle shared-memory descriptors (which of LBO / SBO is which),
//   * kind::f16 with integer-valued fp16 operands and fp32 accumulators (exact below 2^24),
//   * the "shifted accumulator" trick: two MMAs with the same B operand writing D column
//     ranges [0,2Nt) and [Nt,3Nt), with the first K-step of the second one split so that
//     only the fresh columns are overwritten,
//   * kind::i8 (s8 x s8 / s8 x u8, s32 accumulators),
//   * the 1-D bulk copy + mbarrier complete_tx path for the B operand,
//   * tcgen05.ld 32x32b lane/column mapping.
// Build: make -C node_speex_resampler_b200/csrc probe ; run on a B200: ./umma_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "umma_ptx.cuh"

using namespace spxb::ptx;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

struct ProbeArgs {
  const uint8_t *a0;  // [chunks][128][16 B]  chunk-major operand images
  const uint8_t *a1;  // second A operand (shifted-accumulator test) or nullptr
  const uint8_t *b;   // [chunks][nb][16 B]
  uint32_t *d;        // [128][dcols] raw 32-bit accumulators
  int nb;             // rows of B (= N of the MMA)
  int nt;             // shift of the second accumulator range (concat) -- nb == 2*nt
  int dcols;          // columns read back
  int ksteps;         // MMAs along K (each consumes 2 chunks = 32 bytes of K)
  int swap;           // 1: exchange the LBO / SBO fields (hypothesis test)
  int use_bulk;       // 1: B arrives by cp.async.bulk + complete_tx
  int i8;             // 1: kind::i8
  int a_signed, b_signed;
  int *status;        // 0 ok, else where a bounded wait gave up
};

__device__ bool bounded_wait(uint64_t *bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

__global__ void __launch_bounds__(160, 1) probe_kernel(const ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int chunks = 2 * p.ksteps;
  const uint32_t a_chunk = 128 * 16, b_chunk = p.nb * 16;
  uint8_t *sA0 = smem;
  uint8_t *sA1 = sA0 + chunks * a_chunk;
  uint8_t *sB = sA1 + chunks * a_chunk;
  const uint32_t b_bytes = chunks * b_chunk;

  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(&tmem_base_slot, 512);
    tmem_relinquish();
  }
  // operands: generic-proxy stores, then the proxy fence
  for (uint32_t i = tid; i < chunks * a_chunk / 16; i += blockDim.x) {
    reinterpret_cast<uint4 *>(sA0)[i] = reinterpret_cast<const uint4 *>(p.a0)[i];
    if (p.a1) reinterpret_cast<uint4 *>(sA1)[i] = reinterpret_cast<const uint4 *>(p.a1)[i];
  }
  if (!p.use_bulk)
    for (uint32_t i = tid; i < b_bytes / 16; i += blockDim.x)
      reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(p.b)[i];
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 4 && (tid & 31) == 0) {
    if (p.use_bulk) {
      mbar_arrive_expect_tx(&bar_b, b_bytes);
      bulk_g2s(sB, p.b, b_bytes, &bar_b);
      if (!bounded_wait(&bar_b, 0)) {
        *p.status = 1;
      }
    }
    const uint32_t lbo_a = p.swap ? 128u : a_chunk, sbo_a = p.swap ? a_chunk : 128u;
    const uint32_t lbo_b = p.swap ? 128u : b_chunk, sbo_b = p.swap ? b_chunk : 128u;
    const bool concat = p.a1 != nullptr;
    for (int ks = 0; ks < p.ksteps; ++ks) {
      const uint64_t da0 = umma_smem_desc(smem_u32(sA0) + 2 * ks * a_chunk, lbo_a, sbo_a);
      const uint64_t da1 = umma_smem_desc(smem_u32(sA1) + 2 * ks * a_chunk, lbo_a, sbo_a);
      const uint64_t db = umma_smem_desc(smem_u32(sB) + 2 * ks * b_chunk, lbo_b, sbo_b);
      const uint32_t acc = ks > 0;
      if (!p.i8) {
        umma_f16(tmem_base, da0, db, umma_idesc_f16(128, p.nb), acc);
        if (concat) {
          if (ks == 0) {
            // columns [nt,2nt) already hold A0*B: accumulate; columns [2nt,3nt) are fresh
            const uint64_t db_hi = umma_smem_desc(smem_u32(sB) + p.nt * 16, lbo_b, sbo_b);
            umma_f16(tmem_base + p.nt, da1, db, umma_idesc_f16(128, p.nt), 1);
            umma_f16(tmem_base + 2 * p.nt, da1, db_hi, umma_idesc_f16(128, p.nt), 0);
          } else {
            umma_f16(tmem_base + p.nt, da1, db, umma_idesc_f16(128, p.nb), 1);
          }
        }
      } else {
        umma_i8(tmem_base, da0, db, umma_idesc_i8(128, p.nb, p.a_signed, p.b_signed), acc);
      }
    }
    umma_commit(&bar_mma);
  }
  if (warp < 4) {
    if (!bounded_wait(&bar_mma, 0)) {
      if (tid == 0) *p.status = 2;
    } else {
      tc_fence_after_sync();
      const int row = tid;  // TMEM lane
      for (int c0 = 0; c0 < p.dcols; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) p.d[row * p.dcols + c0 + i] = v[i];
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// chunk-major image of a row-major [rows][kbytes] byte matrix
static std::vector<uint8_t> chunk_major(const std::vector<uint8_t> &m, int rows, int kbytes) {
  std::vector<uint8_t> out(m.size());
  const int chunks = kbytes / 16;
  for (int c = 0; c < chunks; ++c)
    for (int r = 0; r < rows; ++r)
      memcpy(&out[(static_cast<size_t>(c) * rows + r) * 16], &m[static_cast<size_t>(r) * kbytes + c * 16], 16);
  return out;
}

static uint32_t rng_state = 12345u;
static int rnd(int lo, int hi) {  // inclusive
  rng_state = rng_state * 1664525u + 1013904223u;
  return lo + static_cast<int>((rng_state >> 8) % static_cast<uint32_t>(hi - lo + 1));
}

template <typename T>
static T *to_dev(const std::vector<T> &v) {
  T *d;
  CK(cudaMalloc(&d, v.size() * sizeof(T)));
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

static int run_case(const char *name, bool i8, bool concat, int nt, int ksteps, int swap, int use_bulk,
                    bool a_signed = true, bool b_signed = true) {
  const int nb = concat ? 2 * nt : nt;
  const int epc = i8 ? 16 : 8;  // elements per 16-byte chunk
  const int K = ksteps * 2 * epc;
  const int dcols = concat ? 3 * nt : nb;
  std::vector<int> A0(128 * K), A1(128 * K), B(nb * K);
  for (auto &v : A0) v = i8 ? (a_signed ? rnd(-128, 127) : rnd(0, 255)) : rnd(-128, 127);
  for (auto &v : A1) v = i8 ? rnd(-128, 127) : rnd(0, 255);
  for (auto &v : B) v = i8 ? (b_signed ? rnd(-128, 127) : rnd(0, 255)) : rnd(-2048, 2047);
  const int esz = i8 ? 1 : 2;
  auto pack = [&](const std::vector<int> &m, int rows) {
    std::vector<uint8_t> raw(static_cast<size_t>(rows) * K * esz);
    for (size_t i = 0; i < m.size(); ++i) {
      if (i8) raw[i] = static_cast<uint8_t>(m[i]);
      else {
        __half h = __float2half(static_cast<float>(m[i]));
        memcpy(&raw[i * 2], &h, 2);
      }
    }
    return chunk_major(raw, rows, K * esz);
  };
  uint8_t *dA0 = to_dev(pack(A0, 128)), *dA1 = to_dev(pack(A1, 128)), *dB = to_dev(pack(B, nb));
  uint32_t *dD;
  int *dStatus;
  CK(cudaMalloc(&dD, 128 * dcols * 4));
  CK(cudaMemset(dD, 0xff, 128 * dcols * 4));
  CK(cudaMalloc(&dStatus, 4));
  CK(cudaMemset(dStatus, 0, 4));
  ProbeArgs p{dA0, concat ? dA1 : nullptr, dB, dD, nb, nt, dcols, ksteps, swap, use_bulk, i8 ? 1 : 0,
              a_signed ? 1 : 0, b_signed ? 1 : 0, dStatus};
  const size_t smem = static_cast<size_t>(2 * ksteps) * (2 * 128 * 16 + nb * 16);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_kernel<<<1, 160, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-44s CUDA error: %s\n", name, cudaGetErrorString(e));
    exit(3);  // context is gone
  }
  int status = 0;
  CK(cudaMemcpy(&status, dStatus, 4, cudaMemcpyDeviceToHost));
  std::vector<uint32_t> D(128 * dcols);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  // exact reference
  long long bad = 0;
  double maxerr = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < dcols; ++c) {
      long long ref = 0;
      auto dot = [&](const std::vector<int> &A, int brow) {
        long long s = 0;
        for (int k = 0; k < K; ++k) s += static_cast<long long>(A[r * K + k]) * B[brow * K + k];
        return s;
      };
      if (!concat) ref = dot(A0, c);
      else {
        if (c < 2 * nt) ref += dot(A0, c);
        if (c >= nt) ref += dot(A1, c - nt);
      }
      double got;
      if (i8) got = static_cast<double>(static_cast<int32_t>(D[r * dcols + c]));
      else {
        float f;
        memcpy(&f, &D[r * dcols + c], 4);
        got = f;
      }
      const double err = std::fabs(got - static_cast<double>(ref));
      if (err > maxerr) maxerr = err;
      if (err != 0.0) ++bad;
    }
  printf("%-44s status=%d mismatches=%lld/%d max|err|=%g %s\n", name, status, bad, 128 * dcols, maxerr,
         (status == 0 && bad == 0) ? "PASS" : "FAIL");
  fflush(stdout);
  cudaFree(dA0);
  cudaFree(dA1);
  cudaFree(dB);
  cudaFree(dD);
  cudaFree(dStatus);
  return (status == 0 && bad == 0) ? 0 : 1;
}

int main(int argc, char **argv) {
  (void)argc;
  (void)argv;
  int fails = 0;
  // 1. descriptor convention (one K-step would not distinguish the K stride: use 2 chunks/MMA)
  fails += run_case("f16 N=64 ks=1 lbo=K-stride", false, false, 64, 1, 0, 0);
  fails += run_case("f16 N=160 ks=4", false, false, 160, 4, 0, 0);
  fails += run_case("f16 N=160 ks=4 bulk B", false, false, 160, 4, 0, 1);
  fails += run_case("f16 N=256 ks=8 bulk B", false, false, 256, 8, 0, 1);
  // 2. shifted accumulators
  fails += run_case("f16 concat Nt=80 ks=1", false, true, 80, 1, 0, 0);
  fails += run_case("f16 concat Nt=80 ks=6 bulk", false, true, 80, 6, 0, 1);
  fails += run_case("f16 concat Nt=112 ks=6 bulk", false, true, 112, 6, 0, 1);
  fails += run_case("f16 concat Nt=128 ks=6 bulk", false, true, 128, 6, 0, 1);
  fails += run_case("f16 concat Nt=16 ks=3", false, true, 16, 3, 0, 0);
  // 3. integer tensor cores
  int i8f = 0;
  i8f += run_case("i8 s8xs8 N=64 ks=2", true, false, 64, 2, 0, 0, true, true);
  i8f += run_case("i8 s8xu8 N=240 ks=4 bulk", true, false, 240, 4, 0, 1, true, false);
  i8f += run_case("i8 u8xs8 N=240 ks=4 bulk", true, false, 240, 4, 0, 1, false, true);
  printf("probe: %d f16 failures, %d i8 failures\n", fails, i8f);
  return fails ? 1 : 0;
}
