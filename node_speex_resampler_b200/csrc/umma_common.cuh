// umma_common.cuh -- device helpers and layout constants shared by the two tensor-core FIR kernels
// (kernels_umma.cu: one tile per CTA, tap tiles streamed; kernels_umma2.cu: persistent CTAs, packed
// tap tiles resident in shared memory).
#pragma once

#include <cstdint>

#include "kernels_common.cuh"
#include "umma_plan.h"
#include "umma_ptx.cuh"

namespace spxb {
namespace ummac {

using namespace ptx;

constexpr int kConvWarps = 8;
constexpr int kConvThreads = kConvWarps * 32;
constexpr int kTmaWarp = 8, kMmaWarp = 9;  // warp 10 slides the history
constexpr int kThreads = 352;
constexpr int kStageChunks = 4;                                  // 64 frames, 2 K steps
// One byte plane of a stage = four K chunks of 128 rows x 16 B. Chunk c sits at
// (c >> 1) * x_kstep + (c & 1) * x_lbo: the pair of chunks of one MMA is LBO apart (a free multiple
// of 16 B), and the paddings are chosen so that one converter store instruction hits 32 distinct
// banks -- mono: a warp stores 4 rows x 8 positions (8 B each), chunks must start 8 banks apart;
// stereo: a warp stores 2 streams x 16 positions into rows 2s (left) or 2s+1 (right), chunk
// starts must be {0, 16, 4, 20} banks (checked exhaustively in tests/test_tensor_plan.py).
__host__ __device__ constexpr uint32_t x_lbo(int ch) { return kUmmaRows * 16 + (ch == 2 ? 64 : 32); }
__host__ __device__ constexpr uint32_t x_kstep(int ch) { return 2 * x_lbo(ch) + (ch == 2 ? 16 : 0); }
__host__ __device__ constexpr uint32_t x_plane(int ch) { return 2 * x_kstep(ch); }
__host__ __device__ constexpr uint32_t x_stage(int ch) { return 2 * x_plane(ch); }  // hi + lo planes
// ring slot of the TMA-fed kernel: a converted stage, rounded up so that every slot can take a TMA box
__host__ __device__ constexpr uint32_t x_slot(int ch) { return (x_stage(ch) + 127u) & ~127u; }
constexpr uint32_t kMaxSmem = 227u * 1024u - 2048u;              // dynamic part; barriers are static

// 16 bytes of one stream's input, sample by sample: the first `n` int16 samples at p, zeros after
// (kept out of line: only the item holding the end of a row, or 2-byte aligned rows, come here)
static __device__ __noinline__ uint4 fetch_item_slow(const int16_t *p, int n) {
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < n) w[i >> 1] |= static_cast<uint32_t>(static_cast<uint16_t>(p[i])) << (16 * (i & 1));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// 16 bytes of one stream's input whatever the alignment of the row: the first `n` int16 samples at p
// (p aligned to `align` bytes: 8, 4 or 2), zeros after. Out of line: the unaligned / ragged paths.
static __device__ __noinline__ uint4 fetch_item_any(const int16_t *p, int n, int align) {
  if (n >= 8 && align >= 8) {
    const uint2 lo = __ldg(reinterpret_cast<const uint2 *>(p)), hi = __ldg(reinterpret_cast<const uint2 *>(p) + 1);
    return make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
  if (n >= 8 && align >= 4) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
    return make_uint4(__ldg(w), __ldg(w + 1), __ldg(w + 2), __ldg(w + 3));
  }
  return fetch_item_slow(p, n);
}

// exactly one lane of a converged warp (elect.sync): the lane that issues tcgen05 / bulk copies
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// two int32 -> saturated int16 pair: (hi << 16) | lo  (the saturation of WORD2INT, arch.h:208-209)
__device__ __forceinline__ uint32_t pack_sat_s16x2(int hi, int lo) {
  uint32_t d;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
  return d;
}

// the 16 accumulator columns of one output group -> rounded (not yet saturated) integers
// y = (p0*2^24 + p1*2^16 + p2*2^8 + p3) * 2^-shift, result floor(y + 1/2) (arch.h:208-209)
__device__ __forceinline__ void combine16(const uint32_t (&p0)[16], const uint32_t (&p1)[16],
                                          const uint32_t (&p2)[16], const uint32_t (&p3)[16], int shift,
                                          int (&r16)[16]) {
  if (shift >= 16 && shift <= 30) {
    // nested floor division by 256 is exact: floor((a*256 + b) / 256) = a + floor(b / 256). The
    // running value after two steps is floor((v + half) / 2^16); it fits int32 (|y| < 2^18 for any
    // windowed-sinc filter, so |v| * 2^-16 < 2^(shift+2)), hence wrapping arithmetic is exact.
    const int half = 1 << (shift - 1), s2 = shift - 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int w = static_cast<int>(p3[i]) + half;
      w = (w >> 8) + static_cast<int>(p2[i]);
      w = (w >> 8) + static_cast<int>(p1[i]) + static_cast<int>(p0[i] << 8);
      r16[i] = w >> s2;
    }
  } else {
    const long long half = 1ll << (shift - 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      long long v = static_cast<long long>(static_cast<int>(p3[i]));
      v += static_cast<long long>(static_cast<int>(p2[i])) << 8;
      v += static_cast<long long>(static_cast<int>(p1[i])) << 16;
      v += static_cast<long long>(static_cast<int>(p0[i])) << 24;
      const long long r = (v + half) >> shift;
      r16[i] = static_cast<int>(max(-40000ll, min(40000ll, r)));
    }
  }
}

// the 8 accumulator columns (persistent kernel: 16 epilogue warps, smaller groups) of one output group -> rounded (not yet saturated) integers
// y = (p0*2^24 + p1*2^16 + p2*2^8 + p3) * 2^-shift, result floor(y + 1/2) (arch.h:208-209)
__device__ __forceinline__ void combine8(const uint32_t (&p0)[8], const uint32_t (&p1)[8],
                                          const uint32_t (&p2)[8], const uint32_t (&p3)[8], int shift,
                                          int (&r8)[8]) {
  if (shift >= 16 && shift <= 30) {
    // nested floor division by 256 is exact: floor((a*256 + b) / 256) = a + floor(b / 256). The
    // running value after two steps is floor((v + half) / 2^16); it fits int32 (|y| < 2^18 for any
    // windowed-sinc filter, so |v| * 2^-16 < 2^(shift+2)), hence wrapping arithmetic is exact.
    const int half = 1 << (shift - 1), s2 = shift - 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int w = static_cast<int>(p3[i]) + half;
      w = (w >> 8) + static_cast<int>(p2[i]);
      w = (w >> 8) + static_cast<int>(p1[i]) + static_cast<int>(p0[i] << 8);
      r8[i] = w >> s2;
    }
  } else {
    const long long half = 1ll << (shift - 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      long long v = static_cast<long long>(static_cast<int>(p3[i]));
      v += static_cast<long long>(static_cast<int>(p2[i])) << 8;
      v += static_cast<long long>(static_cast<int>(p1[i])) << 16;
      v += static_cast<long long>(static_cast<int>(p0[i])) << 24;
      const long long r = (v + half) >> shift;
      r8[i] = static_cast<int>(max(-40000ll, min(40000ll, r)));
    }
  }
}

}  // namespace ummac
}  // namespace spxb
