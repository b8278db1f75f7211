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

// first frame of output m's window in X~ coordinates (history is f < 0), and its phase
__device__ __forceinline__ int window_start(const FilterDev &F, const StreamCall &sc, uint32_t m, uint32_t *phase) {
  const unsigned long long t = static_cast<unsigned long long>(sc.frac0) +
                               static_cast<unsigned long long>(m) * F.num;
  *phase = static_cast<uint32_t>(t % F.den);
  return sc.ls0 - (static_cast<int>(F.taps) - 1) + static_cast<int>(t / F.den);
}

// X(j) = sample j of the output's window, as the f32 the reference holds in `mem`
// `table` = the reference-layout sinc table, in global or (staged) shared memory
template <bool kDirect, bool kWide, typename Window>
__device__ __forceinline__ float strict_output(const FilterDev &F, const float *table, uint32_t phase, const Window &X) {
  const int N = static_cast<int>(F.taps);

  if (kDirect) {
    const float *h = table + static_cast<size_t>(phase) * N;
    if (!kWide) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) {
        const float x = X(j);
        acc = __fadd_rn(acc, __fmul_rn(h[j], x));
      }
      return acc;
    } else {
      double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
      for (int j = 0; j < N; j += 4) {
        const float x0 = X(j);
        const float x1 = X(j + 1);
        const float x2 = X(j + 2);
        const float x3 = X(j + 3);
        a0 = __dadd_rn(a0, static_cast<double>(__fmul_rn(h[j], x0)));
        a1 = __dadd_rn(a1, static_cast<double>(__fmul_rn(h[j + 1], x1)));
        a2 = __dadd_rn(a2, static_cast<double>(__fmul_rn(h[j + 2], x2)));
        a3 = __dadd_rn(a3, static_cast<double>(__fmul_rn(h[j + 3], x3)));
      }
      return static_cast<float>(__dadd_rn(__dadd_rn(__dadd_rn(a0, a1), a2), a3));
    }
  } else {
    const uint32_t os = F.oversample;
    // resample.c:454: prototype cell of this phase; tap k of input j is tp[j*os + k]
    const uint32_t cell = static_cast<uint32_t>((static_cast<unsigned long long>(phase) * os) / F.den);
    const float *tp = table + 4 + os - cell - 2;
    const float4 w = __ldg(reinterpret_cast<const float4 *>(F.blend) + phase);
    if (!kWide) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int j = 0; j < N; ++j) {
        const float x = X(j);
        const float *cf = tp + static_cast<size_t>(j) * os;
        a0 = __fadd_rn(a0, __fmul_rn(x, cf[0]));
        a1 = __fadd_rn(a1, __fmul_rn(x, cf[1]));
        a2 = __fadd_rn(a2, __fmul_rn(x, cf[2]));
        a3 = __fadd_rn(a3, __fmul_rn(x, cf[3]));
      }
      // resample.c:476
      return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w.x, a0), __fmul_rn(w.y, a1)),
                                 __fmul_rn(w.z, a2)),
                       __fmul_rn(w.w, a3));
    } else {
      double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
      for (int j = 0; j < N; ++j) {
        const float x = X(j);
        const float *cf = tp + static_cast<size_t>(j) * os;
        a0 = __dadd_rn(a0, static_cast<double>(__fmul_rn(x, cf[0])));
        a1 = __dadd_rn(a1, static_cast<double>(__fmul_rn(x, cf[1])));
        a2 = __dadd_rn(a2, static_cast<double>(__fmul_rn(x, cf[2])));
        a3 = __dadd_rn(a3, static_cast<double>(__fmul_rn(x, cf[3])));
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

// STAGED: the block first copies the window its 128 outputs read (history || input, as f32) into
// shared memory, coalesced; the tap loops then run without the per-sample history / input / bounds
// branches and unroll. Same values, same operations, same order -- only where the samples are read
// from changes. 3.5x on a single stream's 20 ms call (the path is latency, not throughput, there).
template <bool kDirect, bool kWide, int FMT, bool STAGED>
__global__ void __launch_bounds__(kStrictThreads)
    strict_fir_kernel(const CallArgs a, const uint32_t blocks_per_stream,
                      const uint32_t fir_blocks, const uint32_t win_cap, const uint32_t tab_floats) {
  extern __shared__ float win[];
  if (blockIdx.x >= fir_blocks) {
    history_block_f<FMT>(a, blockIdx.x - fir_blocks);
    return;
  }
  const uint32_t s = launch_stream(a, blockIdx.x / blocks_per_stream);
  const uint32_t e0 = (blockIdx.x % blocks_per_stream) * kStrictThreads;
  const uint32_t e = e0 + threadIdx.x;
  const StreamCall sc = load_call(a, s);
  const uint32_t ch = a.channels;
  const uint32_t m = e / ch;
  const uint32_t c = e % ch;
  uint32_t phase = 0;
  float y = 0.f;
  if (STAGED) {
    // frames [q_lo, q_hi + N) serve every output of the block (q is non-decreasing in m)
    const uint32_t m_lo = e0 / ch;
    if (m_lo >= sc.n_out) return;  // whole block past the end (block-uniform)
    const uint32_t m_hi = min((e0 + kStrictThreads - 1) / ch, sc.n_out - 1);
    uint32_t ph;
    const int q_lo = window_start(a.filt, sc, m_lo, &ph);
    const int q_hi = window_start(a.filt, sc, m_hi, &ph);
    const uint32_t elems = (static_cast<uint32_t>(q_hi - q_lo) + a.filt.taps) * ch;
    if (elems <= win_cap) {
      for (uint32_t i = threadIdx.x; i < elems; i += kStrictThreads)
        win[i] = fetch_sample_f<FMT>(a, s, q_lo + static_cast<int>(i / ch), i % ch, sc.n_in);
      // the sinc table too when it fits beside the window (tab_floats != 0): every tap of the
      // interpolating kernels reads four of its entries
      const float *table = a.filt.table;
      if (tab_floats) {
        float *tab = win + win_cap;
        for (uint32_t i = threadIdx.x; i < tab_floats; i += kStrictThreads) tab[i] = __ldg(a.filt.table + i);
        table = tab;
      }
      __syncthreads();
      if (m >= sc.n_out) return;
      const int q = window_start(a.filt, sc, m, &phase);
      const float *xw = win + static_cast<uint32_t>(q - q_lo) * ch + c;
      y = strict_output<kDirect, kWide>(a.filt, table, phase, [&](int j) { return xw[static_cast<uint32_t>(j) * ch]; });
    } else {
      // (a window wider than the shared memory asked for: straight from global memory)
      if (m >= sc.n_out) return;
      const int q = window_start(a.filt, sc, m, &phase);
      y = strict_output<kDirect, kWide>(a.filt, a.filt.table, phase,
                                        [&](int j) { return fetch_sample_f<FMT>(a, s, q + j, c, sc.n_in); });
    }
  } else {
    if (m >= sc.n_out) return;
    const int q = window_start(a.filt, sc, m, &phase);
    y = strict_output<kDirect, kWide>(a.filt, a.filt.table, phase,
                                      [&](int j) { return fetch_sample_f<FMT>(a, s, q + j, c, sc.n_in); });
  }
  const size_t oe = static_cast<size_t>(m) * a.out_step + c;  // == e unless the call is strided
  if (FMT == 2)  // the float entry stores the kernel's result as is (resample.c:927-963)
    reinterpret_cast<float *>(a.out + static_cast<size_t>(s) * a.out_stride)[oe] = y;
  else if (FMT == 3)  // scaled float PCM out: the int16 result (WORD2INT, exact) over full scale
    reinterpret_cast<float *>(a.out + static_cast<size_t>(s) * a.out_stride)[oe] = static_cast<float>(word2int_exact(y)) * (1.f / 32768.f);
  else
    a.out[static_cast<size_t>(s) * a.out_stride + oe] = word2int_exact(y);
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
  // shared-memory window of a block: its 128 outputs advance by at most ceil(128/ch * num/den) + 1
  // frames, plus the filter length; staged when that fits the default 48 KB
  const uint64_t out_frames = (kStrictThreads + a.channels - 1) / a.channels + 1;
  const uint64_t win_frames = (out_frames * a.filt.num + a.filt.den - 1) / a.filt.den + 2 + a.filt.taps;
  const uint64_t win_bytes = win_frames * a.channels * sizeof(float);
  const bool staged = win_bytes <= 48u * 1024u && a.channels <= kStrictThreads;
  const uint32_t win_cap = staged ? static_cast<uint32_t>(win_bytes / sizeof(float)) : 0u;
  // the reference-layout table beside it when both fit the default 48 KB
  const uint64_t table_floats = a.filt.direct ? static_cast<uint64_t>(a.filt.den) * a.filt.taps
                                              : static_cast<uint64_t>(a.filt.oversample) * a.filt.taps + 8;
  const uint32_t tab_floats = staged && win_bytes + table_floats * sizeof(float) <= 48u * 1024u
                                  ? static_cast<uint32_t>(table_floats) : 0u;
  const size_t smem_bytes = static_cast<size_t>(win_bytes) + static_cast<size_t>(tab_floats) * sizeof(float);
  auto go = [&](auto kernel_staged, auto kernel_plain) {
    if (staged) kernel_staged<<<grid, block, smem_bytes, stream>>>(a, bps_arg, fir_blocks, win_cap, tab_floats);
    else kernel_plain<<<grid, block, 0, stream>>>(a, bps_arg, fir_blocks, 0u, 0u);
  };
  auto by_filter = [&](auto fmt) {
    constexpr int F = decltype(fmt)::value;
    if (a.filt.direct) {
      if (a.filt.wide_accum) go(strict_fir_kernel<true, true, F, true>, strict_fir_kernel<true, true, F, false>);
      else go(strict_fir_kernel<true, false, F, true>, strict_fir_kernel<true, false, F, false>);
    } else {
      if (a.filt.wide_accum) go(strict_fir_kernel<false, true, F, true>, strict_fir_kernel<false, true, F, false>);
      else go(strict_fir_kernel<false, false, F, true>, strict_fir_kernel<false, false, F, false>);
    }
  };
  if (a.fmt == 3) by_filter(std::integral_constant<int, 3>{});
  else if (a.fmt == 2) by_filter(std::integral_constant<int, 2>{});
  else if (a.fmt == 1) by_filter(std::integral_constant<int, 1>{});
  else by_filter(std::integral_constant<int, 0>{});
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace spxb
