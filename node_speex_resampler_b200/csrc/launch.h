// launch.h -- host-callable launchers of the CUDA kernels.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

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

// register-resident FFMA microbenchmark: returns achieved FP32 FLOP/s (2 flops per FMA)
// on the current device; the measured denominator of the fp32 roofline (bench.py)
double measure_fp32_peak_flops(int iters);

}  // namespace spxb
