"""Host logic of the tensor-core kernel (csrc/umma_plan.cpp), checked on CPU: the fixed-point
taps, the tile plan and the int8 tap tiles are pulled through the C ABI's introspection entries
and the kernel's integer pipeline (byte planes x digit rows -> four int32 accumulators ->
64-bit recombination -> round-half-up + saturate) is emulated with numpy, then compared with
the oracle. No GPU, no compute call into the library."""
import ctypes as C

import numpy as np
import pytest

from node_speex_resampler_b200 import lib, synth_pcm
from oracle import oracle as O

CASES = [
    # ch, in, out, q, frames of the measured call, nt
    (2, 44100, 48000, 7, 882, 112),   # C3 (interpolate_single)
    (1, 48000, 16000, 10, 480, 80),   # C4 (direct_double, N = 768)
    (2, 96000, 44100, 10, 640, 128),  # C5 (interpolate_double, N = 560)
    (1, 24000, 48000, 5, 333, 16),    # direct_single, odd length
    (1, 24000, 44100, 9, 500, 48),    # interpolate_double
    (2, 44100, 48000, 0, 200, 96),    # shortest filter (N = 8)
    (1, 8000, 96000, 2, 100, 64),     # x12 up-sampler
]


def fixed_taps(i, o, q):
    L = lib()
    info = __import__("node_speex_resampler_b200")._lib.FilterInfo()
    assert L.spxb_filter_describe(i, o, q, C.byref(info)) == 0
    n = info.den * info.filt_len
    h = np.zeros(n, np.int32)
    sh = C.c_int(0)
    assert L.spxb_filter_fixed_taps(i, o, q, h.ctypes.data, n, C.byref(sh)) == n
    taps = np.zeros(n, np.float32)
    assert L.spxb_filter_phase_taps(i, o, q, taps.ctypes.data, n) == n
    return info, h.reshape(info.den, info.filt_len), sh.value, taps.reshape(info.den, info.filt_len)


@pytest.mark.parametrize("c", CASES, ids=lambda c: f"{c[0]}ch_{c[1]}to{c[2]}_q{c[3]}_nt{c[5]}")
def test_fixed_taps_match_float_taps(c):
    _, i, o, q, _, _ = c
    info, h, sh, taps = fixed_taps(i, o, q)
    assert 1 <= sh <= 30
    assert np.abs(h).max() <= 127 * 65536 + 127 * 256 + 127
    assert np.abs(h).max() > 2 ** 21  # the scale uses the available 24 bits
    # within half a quantisation step (+ the f32 rounding of the float taps)
    err = np.abs(h.astype(np.float64) - taps.astype(np.float64) * 2.0 ** sh)
    assert err.max() <= 0.5 + np.abs(taps).max() * 2.0 ** (sh - 24) + 1e-9


@pytest.mark.parametrize("c", CASES, ids=lambda c: f"{c[0]}ch_{c[1]}to{c[2]}_q{c[3]}_nt{c[5]}")
def test_emulated_integer_pipeline_matches_oracle(c):
    ch, i, o, q, n, nt = c
    L = lib()
    info, h, sh, _ = fixed_taps(i, o, q)
    N, den = info.filt_len, info.den
    hist_frames = (N - 1 + 15) // 16 * 16
    ref = O.OracleResampler(ch, i, o, q)
    # first call only builds up history / a non-trivial stream position
    warm = 777
    x0 = synth_pcm(1, ch, warm, i, seed=5)[0]
    cap0 = int(np.ceil(warm * o / i)) + 2
    _, u0, _ = ref.process(x0, cap0)
    assert u0 == warm
    ls0, frac0, hist = ref.state(0)
    hists = [ref.state(c_)[2].astype(np.int64) for c_ in range(ch)]  # N-1 frames each
    x = synth_pcm(1, ch, n, i, seed=6, start_frame=warm)[0]
    x[: 40 * ch] = 32767  # saturation
    x[40 * ch: 80 * ch] = -32768
    cap = int(np.ceil(n * o / i)) + 2
    want, used, made = ref.process(x, cap)
    assert used == n

    tiles = np.zeros((4096, 4), np.int32)
    ks = C.c_uint32(0)
    nt_tiles = L.spxb_tensor_plan(i, o, q, ls0, frac0, made, nt, tiles.ctypes.data, 4096, C.byref(ks))
    assert nt_tiles == (made + nt - 1) // nt
    K = 32 * ks.value
    got = np.zeros(made * ch, np.int64)
    cache = {}
    for m0, kf0, phase0, delta in tiles[:nt_tiles]:
        assert 0 <= delta < 16 and (kf0 + hist_frames) % 16 == 0 and kf0 >= -hist_frames
        key = (int(phase0), int(delta))
        if key not in cache:
            tile = np.zeros(2 * ks.value * 3 * nt * 16, np.int8)
            assert L.spxb_tensor_tap_tile(i, o, q, nt, phase0, delta, tile.ctypes.data, tile.size) == tile.size
            # [chunk][row][16] -> [row][K]
            cache[key] = tile.reshape(2 * ks.value, 3 * nt, 16).transpose(1, 0, 2).reshape(3 * nt, K).astype(np.int64)
        B = cache[key]
        # digits recombine to the fixed-point taps, shifted to the right window position
        hrec = B[:nt] * 65536 + B[nt:2 * nt] * 256 + B[2 * nt:]
        first = int(delta) + (int(phase0) + 0 * info.num) // den
        assert np.array_equal(hrec[0, first:first + N], h[phase0])
        for c_ in range(ch):
            # X~ of this channel: history (N-1 live frames, zero lead) || input || zeros
            xt = np.zeros(K, np.int64)
            for k in range(K):
                f = kf0 + k
                if f < 0:
                    if f >= -(N - 1):
                        xt[k] = hists[c_][f + (N - 1)]
                elif f < n:
                    xt[k] = x[f * ch + c_]
            hi = xt >> 8           # floor division: s8
            lo = xt & 255          # u8
            assert np.all(hi * 256 + lo == xt) and hi.min() >= -128 and hi.max() <= 127
            Dh = B @ hi            # [3nt]
            Dl = B @ lo
            assert np.abs(Dh).max() < 2 ** 31 and np.abs(Dl).max() < 2 ** 31
            p0 = Dh[:nt]
            p1 = Dh[nt:2 * nt] + Dl[:nt]
            p2 = Dh[2 * nt:] + Dl[nt:2 * nt]
            p3 = Dl[2 * nt:]
            for p in (p0, p1, p2, p3):
                assert np.abs(p).max() < 2 ** 31
            v = (p0 << 24) + (p1 << 16) + (p2 << 8) + p3
            r = np.clip((v + (1 << (sh - 1))) >> sh, -32768, 32767)
            nv = min(nt, made - m0)
            got[(m0 + np.arange(nv)) * ch + c_] = r[:nv]
    d = np.abs(got - want.astype(np.int64))
    assert d.max() <= 1, (d.max(), int(d.argmax()))
    assert O.snr_db(want, got.astype(np.int16)) >= 90.0
    # the integer pipeline is closer to the reference than +-1 LSB suggests: mismatches are rare
    assert (d != 0).mean() < 0.02
