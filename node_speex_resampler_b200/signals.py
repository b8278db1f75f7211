"""Seeded synthetic multi-stream PCM: per stream two sines plus uniform
noise, RMS ~6-7k, every 64th stream carries full-scale square bursts so saturation is
exercised. Used by the parity tests and bench.py; no arithmetic of the hot path here."""
from __future__ import annotations

import numpy as np


def synth_pcm(n_streams: int, channels: int, frames: int, rate: int, seed: int = 0xB200,
              start_frame: int = 0) -> np.ndarray:
    """int16 array [n_streams, frames*channels], interleaved [frame][channel]."""
    rng = np.random.default_rng(seed + 7919 * (start_frame // max(frames, 1)))
    s = np.arange(n_streams, dtype=np.float64)[:, None, None]
    c = np.arange(channels, dtype=np.float64)[None, None, :]
    t = (start_frame + np.arange(frames, dtype=np.float64))[None, :, None] / float(rate)
    f = 100.0 + 37.0 * np.mod(s, 199.0)
    x = 8000.0 * np.sin(2 * np.pi * f * t) + 4000.0 * np.sin(2 * np.pi * 3.1 * f * t + c)
    x = x + rng.uniform(-2000.0, 2000.0, size=(n_streams, frames, channels))
    burst = (np.mod(s, 64.0) == 63.0) & (np.mod(np.floor(t * 50.0), 5.0) == 0.0)
    sq = np.where(np.mod(np.floor(t * 2000.0), 2.0) == 0.0, 32767.0, -32768.0)
    x = np.where(burst, sq + 0.0 * c, x)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16).reshape(n_streams, frames * channels)
