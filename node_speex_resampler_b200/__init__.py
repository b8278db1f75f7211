"""node_speex_resampler_b200 -- B200-native (sm_100a) drop-in for the hot path of
geekuillaume/node-speex-resampler: SpeexResampler.processChunk / processChunks over
hand-written CUDA kernels behind a C ABI (include/speexb200.h). No CPU path."""
from ._lib import KERNEL_AUTO, KERNEL_STRICT, KERNEL_TENSOR, KERNEL_TILED, LIB_PATH, lib  # noqa: F401
from .resampler import (SpeexResampler, SpeexResamplerBatchTransform, SpeexResamplerTransform,  # noqa: F401
                        StreamBatch)
from .formats import WavPcm, wav_pcm  # noqa: F401
from .signals import synth_pcm  # noqa: F401

__all__ = ["SpeexResampler", "SpeexResamplerTransform", "SpeexResamplerBatchTransform", "StreamBatch", "synth_pcm", "wav_pcm", "WavPcm", "lib",
           "KERNEL_AUTO", "KERNEL_STRICT", "KERNEL_TILED", "KERNEL_TENSOR", "LIB_PATH"]
default = SpeexResampler  # `export default SpeexResampler` (src/index.ts:164)
