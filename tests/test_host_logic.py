"""CPU-side checks of the shipped library's host logic (no kernel runs here): the C ABI
loads and exports what include/speexb200.h declares, the filter bank is bit-identical to
the oracle's, the call planner reproduces the reference's consumed/written/state walk, and
the error behaviour of the reference wrapper is mirrored."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cases import GOLDEN_CHUNKS, MATRIX, case_id
from node_speex_resampler_b200 import SpeexResampler, _lib, lib
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC = np.load(os.path.join(ROOT, "tests", "golden", "vectors.npz"))
EXTRA = [(1, 44100, 8000, 6, ""), (1, 8000, 44100, 9, ""), (2, 48000, 48000, 10, ""),
         (1, 192000, 8000, 5, ""), (1, 8000, 192000, 4, ""), (1, 32000, 48000, 2, ""),
         (1, 1000, 300000, 3, ""), (1, 300000, 1000, 3, "")]


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "speexb200.h")).read()
    declared = set(re.findall(r"^SPXB_API[^;(]*?\b(\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(_lib.DECLARED_SYMBOLS), declared ^ set(_lib.DECLARED_SYMBOLS)
    L = lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert b"sm_100a" in L.spxb_version()


def test_strerror_texts_match_reference():
    # deps/speex/resample.c:1222-1239 (code 5 has no case there -> default text)
    want = {0: "Success.", 1: "Memory allocation failed.", 2: "Bad resampler state.",
            3: "Invalid argument.", 4: "Input and output buffers overlap.",
            5: "Unknown error. Bad error code or strange version mismatch.",
            77: "Unknown error. Bad error code or strange version mismatch."}
    for k, v in want.items():
        assert _lib.strerror(k) == v


@pytest.mark.parametrize("args", [(0, 44100, 48000, 7), (2, 0, 48000, 7), (2, 44100, 0, 7),
                                  (2, 44100, 48000, 11), (2, 44100, 48000, -1)])
def test_init_rejects_bad_arguments_like_reference(args):
    # resample.c:804-809 -> NULL + RESAMPLER_ERR_INVALID_ARG, before any allocation
    err = C.c_int(-1)
    assert not lib().speex_resampler_init(*args, C.byref(err))
    assert err.value == 3
    r = SpeexResampler(*args)
    if args[0] == 0:  # JS: length % 0 is NaN -> the chunk-length check throws first
        with pytest.raises(ValueError, match="Chunk length"):
            r.processChunk(b"\0" * 4)
        return
    with pytest.raises(RuntimeError, match="Invalid argument."):
        r.processChunk(b"\0" * (4 * args[0]))
    assert not r._resamplerPtr  # index.ts:61-65: stays falsy, next call retries


def test_misaligned_chunk_message():
    r = SpeexResampler(2, 44100, 48000)
    with pytest.raises(ValueError, match="Chunk length should be a multiple of channels \\* 2 bytes"):
        r.processChunk(b"\0" * 6)


@pytest.mark.parametrize("c", MATRIX + EXTRA, ids=case_id)
def test_filter_bank_bit_identical_to_oracle(c):
    ch, i, o, q, _ = c
    ref = O.OracleResampler(ch, i, o, q)
    p = ref.params
    info = _lib.FilterInfo()
    assert lib().spxb_filter_describe(i, o, q, C.byref(info)) == 0
    for f in ("num", "den", "filt_len", "oversample", "int_advance", "frac_advance", "use_direct",
              "use_double", "table_len"):
        assert getattr(info, f) == getattr(p, f), f
    assert info.cutoff == p.cutoff
    t = np.empty(info.table_len, dtype=np.float32)
    assert lib().spxb_filter_table(i, o, q, t.ctypes.data, t.size) == t.size
    assert np.array_equal(t.view(np.uint32), ref.table().view(np.uint32))


def phase_taps(i, o, q):
    info = _lib.FilterInfo()
    lib().spxb_filter_describe(i, o, q, C.byref(info))
    h = np.empty(info.den * info.filt_len, dtype=np.float32)
    assert lib().spxb_filter_phase_taps(i, o, q, h.ctypes.data, h.size) == h.size
    return info, h.reshape(info.den, info.filt_len)


@pytest.mark.parametrize("c", MATRIX[:19] + MATRIX[20:22], ids=case_id)
def test_phase_taps_reproduce_golden_within_one_lsb(c):
    """The per-phase formulation the tiled kernel contracts with (cubic blend folded into the
    taps) evaluated here in numpy f64 on the first golden chunk: must land within 1 LSB of
    the reference build's output. (A test-side evaluation, not a product code path.)"""
    ch, i, o, q, _ = c
    info, h = phase_taps(i, o, q)
    key = case_id(c)
    n = GOLDEN_CHUNKS[0]
    x = VEC[key + "/in"][0][: n * ch].reshape(n, ch).astype(np.float64)
    n_out = int(VEC[key + "/lens"][0][0])
    want = VEC[key + "/out0"][: n_out * ch].reshape(n_out, ch)
    N = info.filt_len
    X = np.vstack([np.zeros((N - 1, ch)), x, np.zeros((N, ch))])
    m = np.arange(n_out, dtype=np.int64)
    t = m * info.num
    p, ph = t // info.den, t % info.den
    idx = p[:, None] + np.arange(N)[None, :]
    y = np.einsum("mj,mjc->mc", h[ph].astype(np.float64), X[idx])
    got = np.clip(np.floor(y + 0.5), -32768, 32767).astype(np.int64)
    assert np.max(np.abs(got - want.astype(np.int64))) <= 1
    assert O.snr_db(want, got) >= 90.0


def plan(i, o, ls, fr, n_in, cap):
    pl = _lib.CallPlan()
    assert lib().spxb_plan_call(i, o, ls, fr, n_in, cap, C.byref(pl)) == 0
    return pl


@pytest.mark.parametrize("c", MATRIX + EXTRA, ids=case_id)
def test_call_planner_walks_like_the_reference(c):
    """Chain spxb_plan_call against the oracle actually processing samples: consumed,
    written and the (last_sample, samp_frac_num) after every call must agree, including
    calls whose output capacity binds and ratios whose read position jumps whole blocks."""
    ch, i, o, q, _ = c
    ref = O.OracleResampler(1, i, o, q)
    rng = np.random.default_rng(hash((i, o, q)) & 0xFFFF)
    ls, fr = 0, 0
    ratio = o / i
    for step in range(60):
        n_in = int(rng.choice([0, 1, 7, 159, 160, 161, 320, 441, 882, 1000, 2047]))
        mode = step % 4
        natural = int(np.ceil(n_in * ratio))
        cap = [natural + 8, max(natural - int(rng.integers(0, 5)), 0), int(rng.integers(0, 40)),
               natural][mode]
        if ratio > 8:  # keep the oracle's work bounded for extreme up-sampling
            n_in = min(n_in, 320)
            cap = min(cap, 5000)
        x = rng.integers(-2000, 2000, size=n_in, dtype=np.int16)
        _, used, made = ref.process(x, cap)
        pl = plan(i, o, ls, fr, n_in, cap)
        assert (pl.consumed, pl.n_out) == (used, made), (step, n_in, cap, ls, fr)
        ls_ref, fr_ref, _ = ref.state(0)
        assert (pl.last_sample, pl.samp_frac_num) == (ls_ref, fr_ref), (step, n_in, cap)
        ls, fr = pl.last_sample, pl.samp_frac_num


def test_no_gpu_fails_loudly_not_silently():
    if lib().spxb_device_count() > 0:
        pytest.skip("a GPU is present")
    err = C.c_int(0)
    assert not lib().speex_resampler_init(2, 44100, 48000, 7, C.byref(err))
    assert err.value == 1
    assert b"no CUDA device" in lib().spxb_last_error()
    with pytest.raises(RuntimeError, match="Memory allocation failed"):
        SpeexResampler(2, 44100, 48000).processChunk(b"\0" * 8)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        SpeexResampler.initPromise.result()


def test_call_plan_float_entry_walk_matches_oracle():
    """lengths and next position of the float entry (unbounded output block, resample.c:944)
    against the oracle's float path, capacity-bound calls included"""
    from oracle import oracle as O
    L = lib()
    rng = np.random.default_rng(9)
    for i, o in ((8000, 96000), (44100, 48000), (48000, 16000), (96000, 44100), (8000, 48000)):
        r = O.OracleResampler(1, i, o, 3)
        for k in range(40):
            n = int(rng.choice([0, 1, 159, 160, 161, 480, 1000]))
            cap = int(rng.choice([0, 1, 100, 1023, 1024, 1025, 1500, 5000]))
            ls, fr, _ = r.state(0)
            plan = _lib.CallPlan()
            assert L.spxb_plan_call_f32(i, o, ls, fr, n, cap, C.byref(plan)) == 0
            _, used, made = r.process_float(np.zeros(n, np.float32), cap)
            ls1, fr1, _ = r.state(0)
            assert (plan.consumed, plan.n_out, plan.last_sample, plan.samp_frac_num) == (used, made, ls1, fr1), \
                (i, o, k, n, cap)


def test_init_frac_argument_checks_precede_any_allocation():
    """resample.c:804-809 for the _frac entry: zero channels / ratio terms, quality outside 0..10 ->
    NULL + RESAMPLER_ERR_INVALID_ARG, with or without a GPU"""
    L = lib()
    for args in ((0, 441, 480, 44100, 48000, 7), (2, 0, 480, 44100, 48000, 7), (2, 441, 0, 44100, 48000, 7),
                 (2, 441, 480, 44100, 48000, 11), (2, 441, 480, 44100, 48000, -1)):
        err = C.c_int(0)
        assert not L.speex_resampler_init_frac(*args, C.byref(err))
        assert err.value == 3
    # the nominal rates are informational: zero rates pass the argument check (resample.c:804)
    err = C.c_int(0)
    st = L.speex_resampler_init_frac(1, 3, 1, 0, 0, 4, C.byref(err))
    if L.spxb_device_count() <= 0:
        assert not st and err.value == 1      # no device: ALLOC_FAILED, never a CPU fallback
    else:
        assert st and err.value == 0
        L.speex_resampler_destroy(st)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built on this machine")
@pytest.mark.parametrize("float_entry", [False, True], ids=["int16_entry", "float_entry"])
def test_call_plan_with_magic_samples_matches_the_reference_walk(float_entry):
    """The reference build driven through filter changes that shorten the filter mid-stream
    (magic samples pending, memory larger than the filter needs -> input block > 160), then
    through calls of random sizes and capacities: spxb_plan_call_ex must predict consumed /
    written and the next position of every call from the state before it."""
    L, R = lib(), O._load_ref()
    R.speex_resampler_set_quality.restype = R.speex_resampler_set_rate.restype = C.c_int
    R.speex_resampler_set_quality.argtypes = [C.c_void_p, C.c_int]
    R.speex_resampler_set_rate.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    rng = np.random.default_rng(17 + float_entry)
    for (i, o, q0, changes) in ((44100, 48000, 10, [("q", 3)]), (48000, 16000, 8, [("r", 48000, 32000)]),
                               (96000, 44100, 10, [("q", 5), ("q", 5)]), (8000, 96000, 9, [("q", 0)])):
        r = O.RefResampler(1, i, o, q0)
        st = r._st()
        r.process(rng.integers(-9000, 9000, 700).astype(np.int16), 4000)       # started
        for chg in changes:
            if chg[0] == "q":
                assert R.speex_resampler_set_quality(r.h, chg[1]) == 0
            else:
                assert R.speex_resampler_set_rate(r.h, chg[1], chg[2]) == 0
                i, o = chg[1], chg[2]
        assert st.magic_samples[0] > 0
        for k in range(25):
            n = int(rng.choice([0, 1, 50, 160, 161, 400, 1000]))
            cap = int(rng.choice([0, 1, 7, 60, 300, 1023, 1024, 1025, 3000]))
            ls, fr, mg = st.last_sample[0], st.samp_frac_num[0], st.magic_samples[0]
            in_block = st.mem_alloc_size - (st.filt_len - 1)
            plan, used = _lib.CallPlan(), C.c_uint32(0)
            assert L.spxb_plan_call_ex(i, o, ls, fr, mg, n, cap, int(float_entry), in_block, C.byref(plan),
                                       C.byref(used)) == 0
            if float_entry:
                _, u, m = r.process_float(np.zeros(n, np.float32), cap)
            else:
                _, u, m = r.process(np.zeros(n, np.int16), cap)
            assert (plan.consumed, plan.n_out, plan.last_sample, plan.samp_frac_num, mg - used.value) == \
                (u, m, st.last_sample[0], st.samp_frac_num[0], st.magic_samples[0]), (i, o, k, n, cap, mg)


def test_wav_pcm_finds_the_payload(tmp_path):
    """the reference's fixtures are WAV files; the helper returns format + PCM bytes"""
    import io
    import wave

    from node_speex_resampler_b200 import wav_pcm
    frames = (np.arange(1000 * 2, dtype=np.int32) * 37 % 65536 - 32768).astype(np.int16)
    buf = io.BytesIO()
    with wave.open(buf, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(frames.tobytes())
    blob = buf.getvalue()
    info = wav_pcm(blob)
    assert (info.channels, info.sample_rate, info.bits_per_sample, info.format_tag) == (2, 44100, 16, 1)
    assert bytes(info.data) == frames.tobytes()
    # an extra odd-sized chunk in front of `data`, and a truncated data chunk
    riff = bytearray(blob[:36]) + b"LIST" + (3).to_bytes(4, "little") + b"abc\0" + blob[36:]
    riff[4:8] = (len(riff) - 8).to_bytes(4, "little")
    assert bytes(wav_pcm(bytes(riff)).data) == frames.tobytes()
    assert len(wav_pcm(blob[:-3]).data) == frames.nbytes - 4       # whole frames only
    with pytest.raises(ValueError):
        wav_pcm(b"RIFX" + blob[4:])
    res = "/root/reference/resources/44100hz_test.pcm"
    if os.path.exists(res):                                        # the reference's own fixture
        info = wav_pcm(open(res, "rb").read())
        assert (info.channels, info.sample_rate, info.bits_per_sample) == (2, 44100, 16)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built on this machine")
def test_call_plan_input_block_after_a_filter_got_shorter_before_the_first_sample():
    """The reference's memory never shrinks (resample.c:709-719): a state initialised with a long
    filter and switched to a short one BEFORE any sample keeps the long filter's room, so its walk
    takes input blocks longer than 160 frames. spxb_plan_call_ex with that input block must track
    the reference call for call, capacity-bound calls included."""
    L, R = lib(), O._load_ref()
    R.speex_resampler_set_quality.restype = C.c_int
    R.speex_resampler_set_quality.argtypes = [C.c_void_p, C.c_int]
    rng = np.random.default_rng(23)
    r = O.RefResampler(1, 8000, 96000, 10)
    assert R.speex_resampler_set_quality(r.h, 1) == 0
    st = r._st()
    in_block = st.mem_alloc_size - (st.filt_len - 1)
    assert in_block > 160 and st.magic_samples[0] == 0
    differs = 0
    for k in range(60):
        n = int(rng.choice([100, 161, 400, 1000]))
        cap = int(rng.choice([7, 300, 1023, 1500, 3000]))
        ls, fr = st.last_sample[0], st.samp_frac_num[0]
        plan, used, plain = _lib.CallPlan(), C.c_uint32(0), _lib.CallPlan()
        assert L.spxb_plan_call_ex(8000, 96000, ls, fr, 0, n, cap, 0, in_block, C.byref(plan), C.byref(used)) == 0
        assert L.spxb_plan_call(8000, 96000, ls, fr, n, cap, C.byref(plain)) == 0
        _, u, m = r.process(np.zeros(n, np.int16), cap)
        assert (plan.consumed, plan.n_out, plan.last_sample, plan.samp_frac_num) == \
            (u, m, st.last_sample[0], st.samp_frac_num[0]), (k, n, cap)
        differs += (plain.consumed, plain.n_out) != (u, m)
    # (whether the plain 160-frame walk would have differed depends on the ratio; the point is that
    # the walk with the state's real block size is exact)
