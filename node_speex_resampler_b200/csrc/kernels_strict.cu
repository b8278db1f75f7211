// kernels_strict.cu -- "strict" FIR kernel: one thread per output sample, performing the
// reference's own operations in the reference's own order, so results are bit-identical to
// the scalar C / WASM build:
//   resampler_basic_direct_single       deps/speex/resample.c:331-384  f32 mul, f32 add, j ascending
//   resampler_basic_direct_double       :389-435  f32 product into 4 strided f64 sums
//   resampler_basic_interpolate_single  :438-496  4 f32 sums over the oversampled prototype,
//                                                 cubic blend (:318-328) afterwards
//   resampler_basic_interpolate_double  :501-558  same with f64 sums and an f64 blend
// followed by WORD2INT (arch.h:208-209) and the interleaved store (:1018-1022). The same kernels
// serve the float entry (speex_resampler_process_interleaved_float, :1038-1059 over :927-963):
// FMT selects float history and float in/out, and the result is stored unrounded.
// Every multiply/add goes through __fmul_rn/__fadd_rn/__dadd_rn/__dmul_rn so ptxas can never
// contract them into FMAs. Works for any ratio, any per-stream position; it is also the
// fallback when a batch does not qualify for the tiled kernel.
#include <type_traits>

#include "kernels_common.cuh"
#include "launch.h"

namespace spxb {

namespace {

constexpr int kStrictThreads = 128;

template <bool kDirect, bool kWide, int FMT>
__device__ __forceinline__ float strict_output(const CallArgs &a, uint32_t s, uint32_t c,
                                               const StreamCall &sc, uint32_t m) {
  const FilterDev &F = a.filt;
  const int N = static_cast<int>(F.taps);
  const unsigned long long t = static_cast<unsigned long long>(sc.frac0) +
                               static_cast<unsigned long long>(m) * F.num;
  const uint32_t phase = static_cast<uint32_t>(t % F.den);
  // first frame of the window in X~ coordinates (history is f < 0)
  const int q = sc.ls0 - (N - 1) + static_cast<int>(t / F.den);

  if (kDirect) {
    const float *h = F.table + static_cast<size_t>(phase) * N;
    if (!kWide) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) {
        const float x = fetch_sample_f<FMT>(a, s, q + j, c, sc.n_in);
        acc = __fadd_rn(acc, __fmul_rn(__ldg(h + j), x));
      }
      return acc;
    } else {
      double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
      for (int j = 0; j < N; j += 4) {
        const float x0 = fetch_sample_f<FMT>(a, s, q + j, c, sc.n_in);
        const float x1 = fetch_sample_f<FMT>(a, s, q + j + 1, c, sc.n_in);
        const float x2 = fetch_sample_f<FMT>(a, s, q + j + 2, c, sc.n_in);
        const float x3 = fetch_sample_f<FMT>(a, s, q + j + 3, c, sc.n_in);
        a0 = __dadd_rn(a0, static_cast<double>(__fmul_rn(__ldg(h + j), x0)));
        a1 = __dadd_rn(a1, static_cast<double>(__fmul_rn(__ldg(h + j + 1), x1)));
        a2 = __dadd_rn(a2, static_cast<double>(__fmul_rn(__ldg(h + j + 2), x2)));
        a3 = __dadd_rn(a3, static_cast<double>(__fmul_rn(__ldg(h + j + 3), x3)));
      }
      return static_cast<float>(__dadd_rn(__dadd_rn(__dadd_rn(a0, a1), a2), a3));
    }
  } else {
    const uint32_t os = F.oversample;
    // resample.c:454: prototype cell of this phase; tap k of input j is tp[j*os + k]
    const uint32_t cell = static_cast<uint32_t>((static_cast<unsigned long long>(phase) * os) / F.den);
    const float *tp = F.table + 4 + os - cell - 2;
    const float4 w = __ldg(reinterpret_cast<const float4 *>(F.blend) + phase);
    if (!kWide) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int j = 0; j < N; ++j) {
        const float x = fetch_sample_f<FMT>(a, s, q + j, c, sc.n_in);
        const float *cf = tp + static_cast<size_t>(j) * os;
        a0 = __fadd_rn(a0, __fmul_rn(x, __ldg(cf)));
        a1 = __fadd_rn(a1, __fmul_rn(x, __ldg(cf + 1)));
        a2 = __fadd_rn(a2, __fmul_rn(x, __ldg(cf + 2)));
        a3 = __fadd_rn(a3, __fmul_rn(x, __ldg(cf + 3)));
      }
      // resample.c:476
      return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w.x, a0), __fmul_rn(w.y, a1)),
                                 __fmul_rn(w.z, a2)),
                       __fmul_rn(w.w, a3));
    } else {
      double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
      for (int j = 0; j < N; ++j) {
        const float x = fetch_sample_f<FMT>(a, s, q + j, c, sc.n_in);
        const float *cf = tp + static_cast<size_t>(j) * os;
        a0 = __dadd_rn(a0, static_cast<double>(__fmul_rn(x, __ldg(cf))));
        a1 = __dadd_rn(a1, static_cast<double>(__fmul_rn(x, __ldg(cf + 1))));
        a2 = __dadd_rn(a2, static_cast<double>(__fmul_rn(x, __ldg(cf + 2))));
        a3 = __dadd_rn(a3, static_cast<double>(__fmul_rn(x, __ldg(cf + 3))));
      }
      // resample.c:539: f32 weight * f64 sum, summed in f64, demoted once
      const double r = __dadd_rn(
          __dadd_rn(__dadd_rn(__dmul_rn(static_cast<double>(w.x), a0),
                              __dmul_rn(static_cast<double>(w.y), a1)),
                    __dmul_rn(static_cast<double>(w.z), a2)),
          __dmul_rn(static_cast<double>(w.w), a3));
      return static_cast<float>(r);
    }
  }
}

template <bool kDirect, bool kWide, int FMT>
__global__ void __launch_bounds__(kStrictThreads)
    strict_fir_kernel(const CallArgs a, const uint32_t blocks_per_stream,
                      const uint32_t fir_blocks) {
  if (blockIdx.x >= fir_blocks) {
    history_block_f<FMT>(a, blockIdx.x - fir_blocks);
    return;
  }
  const uint32_t s = launch_stream(a, blockIdx.x / blocks_per_stream);
  const uint32_t e = (blockIdx.x % blocks_per_stream) * kStrictThreads + threadIdx.x;
  const StreamCall sc = load_call(a, s);
  const uint32_t m = e / a.channels;
  const uint32_t c = e % a.channels;
  if (m >= sc.n_out) return;
  const float y = strict_output<kDirect, kWide, FMT>(a, s, c, sc, m);
  if (FMT == 2)  // the float entry stores the kernel's result as is (resample.c:927-963)
    reinterpret_cast<float *>(a.out + static_cast<size_t>(s) * a.out_stride)[e] = y;
  else
    a.out[static_cast<size_t>(s) * a.out_stride + e] = word2int_exact(y);
}

}  // namespace

uint32_t hist_blocks(const CallArgs &a, uint32_t threads) {
  const uint32_t hist_elems = a.hist_frames * a.channels;
  return (a.ids ? a.n_ids : a.n_streams) * ((hist_elems + threads - 1) / threads);
}

cudaError_t launch_strict(const CallArgs &a, cudaStream_t stream, uint32_t *launches) {
  const uint64_t elems = static_cast<uint64_t>(a.max_n_out) * a.channels;
  const uint32_t bps = static_cast<uint32_t>((elems + kStrictThreads - 1) / kStrictThreads);
  const uint64_t fir_blocks64 = static_cast<uint64_t>(a.ids ? a.n_ids : a.n_streams) * bps;
  const uint64_t total = fir_blocks64 + hist_blocks(a, kStrictThreads);
  if (total == 0) return cudaSuccess;
  if (total > 0x7fffffffull) return cudaErrorInvalidConfiguration;
  const uint32_t fir_blocks = static_cast<uint32_t>(fir_blocks64);
  const dim3 grid(static_cast<uint32_t>(total)), block(kStrictThreads);
  const uint32_t bps_arg = bps ? bps : 1;
  auto go = [&](auto kernel) { kernel<<<grid, block, 0, stream>>>(a, bps_arg, fir_blocks); };
  auto by_filter = [&](auto fmt) {
    constexpr int F = decltype(fmt)::value;
    if (a.filt.direct) {
      if (a.filt.wide_accum) go(strict_fir_kernel<true, true, F>);
      else go(strict_fir_kernel<true, false, F>);
    } else {
      if (a.filt.wide_accum) go(strict_fir_kernel<false, true, F>);
      else go(strict_fir_kernel<false, false, F>);
    }
  };
  if (a.fmt == 2) by_filter(std::integral_constant<int, 2>{});
  else if (a.fmt == 1) by_filter(std::integral_constant<int, 1>{});
  else by_filter(std::integral_constant<int, 0>{});
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace spxb
