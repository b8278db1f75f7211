// launch.h -- host-callable launchers of the CUDA kernels.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "device_types.h"

namespace spxb {

// blocks appended to a FIR grid for the fused history slide
uint32_t hist_blocks(const CallArgs &a, uint32_t threads);

// strict kernel: any ratio, any per-stream state; bit-exact against the scalar reference
cudaError_t launch_strict(const CallArgs &a, cudaStream_t stream, uint32_t *launches);

// tiled kernel: register-tiled per-phase FIR. `tiled_qualifies` says whether this call's
// shape is covered (uniform stream positions, window fits shared memory, ...).
struct TiledConfig {
  int variant = 0;         // index into the compiled tile shapes
  uint32_t smem_bytes = 0;
  uint32_t grid = 0;
};
bool tiled_qualifies(const CallArgs &a, int sm_count, TiledConfig *cfg);
cudaError_t launch_tiled(const CallArgs &a, const TiledConfig &cfg, cudaStream_t stream,
                         uint32_t *launches);
// one-time per-device attribute setup (opt-in shared memory sizes)
cudaError_t tiled_prepare_device();

// tensor kernel (kernels_umma.cu): exact integer banded GEMM on the int8 tensor cores.
// A context holds the batch's fixed-point taps and tap-tile pool. umma_prepare plans the call
// (tiles, tap tiles it still has to build -- queued on `stream`) and says whether the call is
// covered (uniform stream positions, mono/stereo, ...); launch_umma then runs the FIR.
struct UmmaContext;
struct FilterSpec;
UmmaContext *umma_create(const FilterSpec &spec, const std::vector<float> &ref_table, uint32_t channels,
                         int sm_count);
void umma_destroy(UmmaContext *c);
int umma_shift(const UmmaContext *c);
// last planned call: {outputs per tile, K steps per tile, tiles, series groups, smem stages, smem bytes}
void umma_geometry(const UmmaContext *c, uint32_t out[6]);
bool umma_prepare(UmmaContext *c, const CallArgs &a, cudaStream_t stream, cudaError_t *err);
cudaError_t launch_umma(UmmaContext *c, const CallArgs &a, cudaStream_t stream, uint32_t *launches);
// CUDA-graph support: while frozen, umma_prepare refuses (cudaErrorNotSupported) any call that
// would need the stream or an allocation; stream_ops counts such operations since creation;
// pool_generation changes whenever earlier launches' tap-tile pointers go stale.
void umma_set_frozen(UmmaContext *c, bool frozen);
uint64_t umma_stream_ops(const UmmaContext *c);
uint64_t umma_pool_generation(const UmmaContext *c);
// persistent kernel with resident packed tap tiles (kernels_umma2.cu), chosen by umma_prepare when
// the packed tile of the geometry fits shared memory
void umma2_configure_device();
bool umma2_planes_in_tmem();  // SPXB_UMMA_ATMEM (default on)
bool umma2_covers(const CallArgs &a);  // whole batch, rows on 16-byte boundaries, an input to read
cudaError_t umma2_upload_plan(UmmaContext *c, cudaStream_t stream);
uint32_t umma2_x_stages(uint32_t channels, uint32_t tile_bytes, uint32_t ksteps);
cudaError_t umma2_build_tiles(UmmaContext *c, const uint32_t *d_jobs, size_t n_jobs, cudaStream_t stream);
cudaError_t launch_umma2(UmmaContext *c, const CallArgs &a, cudaStream_t stream, uint32_t *launches);
// SPXB_UMMA_TRACE=1: clock64 timeline of the last launch, 32 words per CTA (debug only)
long umma_read_trace(const UmmaContext *c, unsigned long long *dst, size_t cap_words);

// register-resident FFMA microbenchmark: returns achieved FP32 FLOP/s (2 flops per FMA)
// on the current device; the measured denominator of the fp32 roofline (bench.py)
double measure_fp32_peak_flops(int iters);

}  // namespace spxb
