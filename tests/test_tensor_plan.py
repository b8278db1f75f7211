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


# ---------------------------------------------------------------------------------------------
# packed ("resident") tap tiles of the persistent tensor kernel (csrc/umma_plan.h, kernels_umma2.cu)
# ---------------------------------------------------------------------------------------------
def packed_plan(i, o, q, nt):
    L = lib()
    words = np.zeros(64 * 18, np.uint32)
    tb = C.c_uint32(0)
    ks = L.spxb_tensor_packed_plan(i, o, q, nt, words.ctypes.data, words.size, C.byref(tb))
    assert ks > 0, ks
    return words[: ks * 18].reshape(ks, 18).astype(np.int64), tb.value


@pytest.mark.parametrize("c", CASES, ids=lambda c: f"{c[0]}ch_{c[1]}to{c[2]}_q{c[3]}_nt{c[5]}")
def test_packed_tiles_hold_every_nonzero_of_the_dense_tiles(c):
    """For a spread of tile keys the packed tile, unpacked through its own block table, must equal
    the dense tile exactly: nothing non-zero falls outside the stored blocks, the stored cells are
    the dense cells, K step 0 is whole, offsets are the running sum of the rows."""
    L = lib()
    ch, i, o, q, n, nt = c
    info, h, sh, _ = fixed_taps(i, o, q)
    plan, tile_bytes = packed_plan(i, o, q, nt)
    ks = plan.shape[0]
    off = 0
    for k in range(ks):
        assert plan[k, 0] == off
        rows = sum(16 * (plan[k, 15 + d] - plan[k, 12 + d]) for d in range(3))
        assert plan[k, 1] == rows
        off += 2 * rows
    assert off * 16 == tile_bytes
    assert all(plan[0, 12 + d] == 0 and plan[0, 15 + d] == nt // 16 for d in range(3))
    dense_bytes = 2 * ks * 3 * nt * 16
    assert tile_bytes <= dense_bytes
    rng = np.random.default_rng(nt)
    keys = [(0, 0), (info.den - 1, 15), (info.den // 2, 7)] + [(int(rng.integers(info.den)), int(rng.integers(16))) for _ in range(5)]
    for phase0, delta in keys:
        dense = np.zeros(dense_bytes, np.int8)
        assert L.spxb_tensor_tap_tile(i, o, q, nt, phase0, delta, dense.ctypes.data, dense.size) == dense.size
        dense = dense.reshape(2 * ks, 3 * nt, 16)
        packed = np.zeros(tile_bytes, np.int8)
        assert L.spxb_tensor_tap_tile_packed(i, o, q, nt, phase0, delta, packed.ctypes.data, packed.size) == tile_bytes
        rebuilt = np.zeros_like(dense)
        for k in range(ks):
            rows = plan[k, 1]
            for half in range(2):
                cells = packed[(plan[k, 0] + half * rows) * 16: (plan[k, 0] + (half + 1) * rows) * 16].reshape(rows, 16)
                r = 0
                for d in range(3):
                    n0, n1 = 16 * plan[k, 12 + d], 16 * plan[k, 15 + d]
                    rebuilt[2 * k + half, d * nt + n0: d * nt + n1] = cells[r: r + (n1 - n0)]
                    r += n1 - n0
                assert r == rows
        assert np.array_equal(rebuilt, dense), (phase0, delta, np.argwhere(rebuilt != dense)[:4])


@pytest.mark.parametrize("c", CASES[:5], ids=lambda c: f"{c[0]}ch_{c[1]}to{c[2]}_q{c[3]}_nt{c[5]}")
def test_packed_mma_schedule_reproduces_the_dense_gemm(c):
    """The kernel's MMA schedule over the packed tile -- K step 0 whole, then per K step the table's
    entries (first B row, columns, first accumulator column; lo plane one digit block to the right) --
    must leave in the four accumulator blocks exactly what the dense banded GEMM leaves."""
    L = lib()
    ch, i, o, q, n, nt = c
    info, h, sh, _ = fixed_taps(i, o, q)
    plan, tile_bytes = packed_plan(i, o, q, nt)
    ks = plan.shape[0]
    rng = np.random.default_rng(7)
    phase0, delta = int(rng.integers(info.den)), int(rng.integers(16))
    dense = np.zeros(2 * ks * 3 * nt * 16, np.int8)
    L.spxb_tensor_tap_tile(i, o, q, nt, phase0, delta, dense.ctypes.data, dense.size)
    B = dense.reshape(2 * ks, 3 * nt, 16).transpose(1, 0, 2).reshape(3 * nt, 32 * ks).astype(np.int64)
    packed = np.zeros(tile_bytes, np.int8)
    L.spxb_tensor_tap_tile_packed(i, o, q, nt, phase0, delta, packed.ctypes.data, packed.size)
    x = rng.integers(-32768, 32768, size=32 * ks).astype(np.int64)
    hi, lo = x >> 8, x & 255
    want = np.zeros(4 * nt, np.int64)
    want[: 3 * nt] += B @ hi
    want[nt:] += B @ lo
    acc = np.full(4 * nt, 10 ** 12, np.int64)  # garbage until initialised
    for k in range(ks):
        rows = plan[k, 1]
        chunk = [packed[(plan[k, 0] + half * rows) * 16: (plan[k, 0] + (half + 1) * rows) * 16].reshape(rows, 16).astype(np.int64)
                 for half in range(2)]
        xs = [x[32 * k + 16 * half: 32 * k + 16 * half + 16] for half in range(2)]
        if k == 0:
            assert rows == 3 * nt
            full = sum(chunk[half] @ (xs[half] >> 8) for half in range(2))
            full_lo = sum(chunk[half] @ (xs[half] & 255) for half in range(2))
            acc[: 3 * nt] = full                    # hi plane, accumulate = 0
            acc[nt: 3 * nt] += full_lo[: 2 * nt]    # lo plane onto P1, P2
            acc[3 * nt:] = full_lo[2 * nt:]         # lo plane, P3 fresh
            continue
        assert 0 <= plan[k, 2] <= 3
        covered = 0
        for e in range(plan[k, 2]):
            row, ncol, dcol = plan[k, 3 + 3 * e], plan[k, 4 + 3 * e], plan[k, 5 + 3 * e]
            assert ncol % 16 == 0 and 16 <= ncol <= 256 and row + ncol <= rows and dcol + ncol <= 3 * nt
            for plane, shift in ((0, 0), (1, nt)):
                v = sum(chunk[half][row: row + ncol] @ ((xs[half] >> 8) if plane == 0 else (xs[half] & 255))
                        for half in range(2))
                acc[dcol + shift: dcol + shift + ncol] += v
            covered += ncol
        assert covered == rows
    assert np.array_equal(acc, want)
