"""Parity of the CUDA path against the oracle, through the C ABI / the reference-shaped
wrapper. Needs a B200:  python -m pytest tests -m gpu

Bars (BASELINE.json north_star): strict kernel bit-exact; tiled (fp32 FMA) and tensor
(tcgen05 int8) kernels within +-1 LSB per int16 sample and >= 90 dB SNR against the reference
output. The golden vectors were
produced by the reference's own C (oracle/gen_golden.py); the oracle restatement is the
live checker for everything else.
"""
import ctypes as C
import os

import numpy as np
import pytest

from cases import GOLDEN_CHUNKS, GOLDEN_STREAMS, MATRIX, case_id
from node_speex_resampler_b200 import (KERNEL_AUTO, KERNEL_STRICT, KERNEL_TENSOR, KERNEL_TILED, SpeexResampler,
                                       SpeexResamplerBatchTransform, SpeexResamplerTransform, StreamBatch, _lib, lib,
                                       synth_pcm)
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC = np.load(os.path.join(ROOT, "tests", "golden", "vectors.npz"))
LSB_TOL = 1        # north_star: +-1 LSB per int16 sample
SNR_MIN_DB = 90.0  # north_star: >= 90 dB against the reference output


def check_close(want: np.ndarray, got: np.ndarray, exact: bool, what=""):
    assert want.shape == got.shape, (what, want.shape, got.shape)
    if exact:
        bad = np.flatnonzero(want != got)
        assert bad.size == 0, (what, "first mismatch at", bad[:5], want[bad[:5]], got[bad[:5]])
        return
    d = np.abs(want.astype(np.int32) - got.astype(np.int32))
    assert d.max(initial=0) <= LSB_TOL, (what, "max diff", d.max(), "at", int(d.argmax()))
    if want.size and np.any(want):
        assert O.snr_db(want, got) >= SNR_MIN_DB, (what, O.snr_db(want, got))


def new_resampler(c, kernel):
    ch, i, o, q, _ = c
    r = SpeexResampler(ch, i, o, q)
    r.kernel = kernel
    return r


# ---------------------------------------------------------------------------
# golden vectors (made by the real reference build) through processChunk
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("c", MATRIX, ids=case_id)
def test_golden_strict_is_bit_exact(c):
    ch = c[0]
    key = case_id(c)
    for s in range(GOLDEN_STREAMS):
        r = new_resampler(c, KERNEL_STRICT)
        pos, outs = 0, []
        for k, n in enumerate(GOLDEN_CHUNKS):
            y = np.frombuffer(r.processChunk(VEC[key + "/in"][s][pos * ch:(pos + n) * ch]), dtype=np.int16)
            assert y.size // ch == VEC[key + "/lens"][s][k]
            outs.append(y)
            pos += n
        check_close(VEC[key + f"/out{s}"], np.concatenate(outs), exact=True, what=key)
        r.destroy()


@pytest.mark.parametrize("kernel", [KERNEL_TILED, KERNEL_TENSOR], ids=["tiled", "tensor"])
@pytest.mark.parametrize("c", [c for c in MATRIX if c[0] <= 2], ids=case_id)
def test_golden_fast_kernels_within_one_lsb(c, kernel):
    ch = c[0]
    key = case_id(c)
    r = new_resampler(c, kernel)
    pos, outs = 0, []
    for k, n in enumerate(GOLDEN_CHUNKS):
        try:
            raw = r.processChunk(VEC[key + "/in"][0][pos * ch:(pos + n) * ch])
        except RuntimeError as e:
            assert "Bad resampler state" in str(e) and "does not qualify" in _lib.last_error()
            pytest.skip("this kernel does not cover this filter (strict serves it)")
        y = np.frombuffer(raw, dtype=np.int16)
        assert y.size // ch == VEC[key + "/lens"][0][k]
        outs.append(y)
        pos += n
    check_close(VEC[key + "/out0"], np.concatenate(outs), exact=False, what=key)
    r.destroy()


# ---------------------------------------------------------------------------
# batched streams with state carry-over (BASELINE configs 3-5 at reduced stream counts)
# ---------------------------------------------------------------------------
SHAPES = [
    # name, n_streams, ch, in, out, q, frames per 20 ms call, calls
    ("C3", 70, 2, 44100, 48000, 7, 882, 50),
    ("C4", 66, 1, 48000, 16000, 10, 960, 30),
    ("C5", 65, 2, 96000, 44100, 10, 1920, 10),
    ("up2_direct", 37, 1, 24000, 48000, 5, 480, 20),
    ("sweep_q9", 16, 1, 24000, 44100, 9, 480, 20),
]


@pytest.mark.parametrize("kernel", [KERNEL_STRICT, KERNEL_TILED, KERNEL_TENSOR, KERNEL_AUTO],
                         ids=["strict", "tiled", "tensor", "auto"])
@pytest.mark.parametrize("shape", SHAPES, ids=[s[0] for s in SHAPES])
def test_batch_state_carry(shape, kernel):
    name, S, ch, i, o, q, n, calls = shape
    b = StreamBatch(S, ch, i, o, q)
    b.set_kernel(kernel)
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(S)]
    cap = int(np.ceil(n * o / i)) + 2
    kernels_used = set()
    saturated = 0
    for k in range(calls):
        pcm = synth_pcm(S, ch, n, i, seed=0xC0DE, start_frame=k * n)
        out, used, made = b.process(pcm, n, cap)
        kernels_used.add(b.last_kernel())
        for s in range(S):
            y, u, m = refs[s].process(pcm[s], cap)
            assert (u, m) == (int(used[s]), int(made[s])), (name, k, s)
            check_close(y, out[s, : m * ch], exact=kernel == KERNEL_STRICT, what=(name, k, s))
            if s % 64 == 63:  # the stream with full-scale square bursts: WORD2INT's clamp is live
                hit = (y == 32767) | (y == -32768)
                saturated += int(np.count_nonzero(hit))
                got_hit = (out[s, : m * ch] == 32767) | (out[s, : m * ch] == -32768)
                if kernel == KERNEL_STRICT:
                    assert np.array_equal(hit, got_hit), (name, k, s)
                else:  # a value that rounds to exactly +-32767/8 unclamped may sit 1 LSB inside
                    assert abs(int(np.count_nonzero(got_hit)) - int(np.count_nonzero(hit))) <= 2 + hit.size // 200
    if S > 63:
        assert saturated > 0, "stream 63 should drive the output into saturation"
    if kernel == KERNEL_AUTO:
        assert kernels_used == {KERNEL_TENSOR}, "auto should pick the tensor kernel for uniform mono/stereo batches"
    else:
        assert kernels_used == {kernel}
    # device-resident state equals the oracle's
    for s in (0, S - 1):
        ls, fr, mg, hist = b.get_state(s)
        for c_ in range(ch):
            rls, rfr, rhist = refs[s].state(c_)
            assert (ls, fr, mg) == (rls, rfr, 0)
            assert np.array_equal(hist.reshape(-1, ch)[:, c_].astype(np.float32), rhist)
    b.close()


def test_full_size_c3_properties():
    """BASELINE configs[2] at full size (1024 stereo streams, 882 -> 960 frames per call):
    size-independent properties + a sample of streams against the oracle."""
    S, ch, i, o, q, n = 1024, 2, 44100, 48000, 7, 882
    b = StreamBatch(S, ch, i, o, q)
    pick = [0, 1, 63, 511, 1023]
    refs = {s: O.OracleResampler(ch, i, o, q) for s in pick}
    twin = StreamBatch(S, ch, i, o, q)  # same inputs, streams permuted: outputs must permute
    perm = np.random.default_rng(3).permutation(S)
    for k in range(12):
        pcm = synth_pcm(S, ch, n, i, seed=0xFEED, start_frame=k * n)
        pcm[7] = pcm[5]  # duplicate input -> duplicate output (streams are independent)
        out, used, made = b.process(pcm, n, 960)
        assert np.all(used == n) and np.all(made == 960)  # 6 whole phase periods per call
        assert np.array_equal(out[7], out[5])
        out2, _, _ = twin.process(pcm[perm], n, 960)
        assert np.array_equal(out2, out[perm])
        for s in pick:
            y, _, m = refs[s].process(pcm[s], 960)
            check_close(y, out[s, : m * ch], exact=False, what=("C3full", k, s))
    assert b.last_kernel() == KERNEL_TENSOR
    b.close()
    twin.close()


@pytest.mark.parametrize("name,S,ch,i,o,q,n,calls", [
    ("C4", 4096, 1, 48000, 16000, 10, 960, 6),    # BASELINE configs[3]: long direct filter (N = 768)
    ("C5", 8192, 2, 96000, 44100, 10, 1920, 3),   # BASELINE configs[4], one GPU's share of 65536 streams
], ids=["C4", "C5"])
def test_full_size_long_filter_properties(name, S, ch, i, o, q, n, calls):
    """BASELINE configs[3] / [4] at full per-GPU size on the tensor kernel: every call consumes
    and produces whole frames, streams are independent (duplicate input -> duplicate output,
    a permuted batch gives permuted outputs), a sample of streams matches the oracle within
    1 LSB / 90 dB, and the carried state equals the oracle's."""
    cap = -(-n * o // i)
    b = StreamBatch(S, ch, i, o, q)
    twin = StreamBatch(S, ch, i, o, q)
    pick = [0, 1, S // 2 + 3, S - 1]
    refs = {s: O.OracleResampler(ch, i, o, q) for s in pick}
    perm = np.random.default_rng(11).permutation(S)
    base = synth_pcm(256, ch, n * calls, i, seed=0xC0FFEE)
    sel = np.resize(np.arange(256), S)
    for k in range(calls):
        pcm = np.ascontiguousarray(base[np.roll(sel, 3 * k), k * n * ch:(k + 1) * n * ch])
        pcm[7] = pcm[5]
        out, used, made = b.process(pcm, n, cap)
        assert b.last_kernel() == KERNEL_TENSOR
        assert np.all(used == n) and np.all(made == made[0]) and made[0] in (cap, cap - 1)
        assert np.array_equal(out[7], out[5])
        out2, _, _ = twin.process(pcm[perm], n, cap)
        assert np.array_equal(out2, out[perm])
        for s in pick:
            y, u_, m = refs[s].process(pcm[s], cap)
            assert (u_, m) == (int(used[s]), int(made[s]))
            check_close(y, out[s, : m * ch], exact=False, what=(name, k, s))
    for s in pick:
        ls, fr, mg, hist = b.get_state(s)
        rls, rfr, rhist = refs[s].state(0)
        assert (ls, fr, mg) == (rls, rfr, 0)
        assert np.array_equal(hist.reshape(-1, ch)[:, 0].astype(np.float32), rhist)
    b.close()
    twin.close()


# ---------------------------------------------------------------------------
# the reference wrapper's observable behaviour (src/index.ts:50-116, :121-162)
# ---------------------------------------------------------------------------
def test_process_chunks_equals_per_stream_process_chunk():
    """ragged chunk lengths, capacity rule and silent input drop, per stream"""
    ch, i, o, q = 2, 44100, 48000, 7
    S = 9
    rs = [SpeexResampler(ch, i, o, q) for _ in range(S)]
    for r in rs:
        r.kernel = KERNEL_STRICT
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(S)]
    rng = np.random.default_rng(5)
    for k in range(25):
        chunks = []
        for s in range(S):
            n = int(rng.choice([0, 1, 100, 441, 882, 1000, 1234]))
            chunks.append(synth_pcm(1, ch, max(n, 1), i, seed=s * 100 + k)[0][: n * ch].tobytes())
        got = SpeexResampler.processChunks(rs, chunks)
        for s in range(S):
            assert got[s] == refs[s].processChunk(chunks[s]), (k, s)


def test_process_chunk_then_process_chunks_migrates_state():
    ch, i, o, q = 1, 48000, 16000, 10
    rs = [SpeexResampler(ch, i, o, q) for _ in range(3)]
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(3)]
    for r in rs:
        r.kernel = KERNEL_STRICT
    x = synth_pcm(3, ch, 4000, i, seed=9)
    for s in range(3):
        assert rs[s].processChunk(x[s, :1000]) == refs[s].processChunk(x[s, :1000])
    got = SpeexResampler.processChunks(rs, [x[s, 1000:2500] for s in range(3)])
    for s in range(3):
        assert got[s] == refs[s].processChunk(x[s, 1000:2500])
    got = SpeexResampler.processChunks(rs, [x[s, 2500:] for s in range(3)])
    for s in range(3):
        assert got[s] == refs[s].processChunk(x[s, 2500:])


@pytest.mark.parametrize("c", MATRIX[:7], ids=case_id)
def test_transform_stream_like_reference_test(c):
    """src/test.ts:46-77: pipe the data through SpeexResamplerTransform in 64 KiB reads (odd
    sizes added so the alignment carry of index.ts:139-154 is exercised); only the duration
    is asserted there, here the bytes are checked too."""
    ch, i, o, q, _ = c
    data = synth_pcm(1, ch, 60000, i, seed=77)[0].tobytes()
    data = b"RIFF" + data[4:]  # the fixtures are WAV files fed header and all
    t = SpeexResamplerTransform(ch, i, o, q)
    t.resampler.kernel = KERNEL_STRICT
    ref = O.OracleResampler(ch, i, o, q)
    sizes = [65536, 65536, 4097, 3, 65536, 1, 30001]
    pos, got, want, carry = 0, b"", b"", b""
    k = 0
    while pos < len(data):
        n = sizes[k % len(sizes)]
        k += 1
        chunk = data[pos:pos + n]
        pos += n
        got += t.transform(chunk)
        buf = carry + chunk
        extra = len(buf) % (ch * 2)
        carry = buf[len(buf) - extra:] if extra else b""
        want += ref.processChunk(buf[: len(buf) - extra])
    assert got == want
    din = len(data) / i / 2 / ch
    dout = len(got) / o / 2 / ch
    assert abs(din - dout) < 0.01  # the reference's own assertion (src/test.ts:74)


def test_batch_transform_equals_per_stream_transforms():
    """the multi-stream Transform. Every stream gets its own odd-sized writes
    (alignment carry per stream, src/index.ts:139-154; empty writes included); stream i's bytes
    must equal what the oracle gives for the same aligned chunks."""
    ch, i, o, q, S = 2, 44100, 48000, 7, 6
    t = SpeexResamplerBatchTransform(S, ch, i, o, q)
    for r in t.resamplers:
        r.kernel = KERNEL_STRICT
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(S)]
    data = [synth_pcm(1, ch, 9000, i, seed=300 + s)[0].tobytes() for s in range(S)]
    rng = np.random.default_rng(5)
    pos = [0] * S
    carry = [b""] * S
    got = [b""] * S
    want = [b""] * S
    for k in range(14):
        sizes = rng.choice([0, 1, 3, 4, 882 * 4, 4097, 2999], size=S)
        chunks = []
        for s in range(S):
            c = data[s][pos[s]:pos[s] + int(sizes[s])]
            pos[s] += len(c)
            chunks.append(c)
            buf = carry[s] + c
            extra = len(buf) % (ch * 2)
            carry[s] = buf[len(buf) - extra:] if extra else b""
            want[s] += refs[s].processChunk(buf[: len(buf) - extra])
        res = t.transform(chunks)
        for s in range(S):
            got[s] += res[s]
    assert got == want
    with pytest.raises(ValueError):
        t.transform([b""] * (S - 1))


def test_edge_lengths_and_capacities():
    """empty chunks, single frames, zero capacity, capacity that binds mid-block"""
    ch, i, o, q = 2, 44100, 24000, 5
    b = StreamBatch(5, ch, i, o, q)
    b.set_kernel(KERNEL_STRICT)
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(5)]
    rng = np.random.default_rng(11)
    for k in range(40):
        n_in = rng.choice([0, 1, 2, 159, 160, 161, 500], size=5).astype(np.uint32)
        cap = rng.choice([0, 1, 50, 87, 88, 300], size=5).astype(np.uint32)
        pcm = synth_pcm(5, ch, 500, i, seed=k)
        out, used, made = b.process(pcm, n_in, cap)
        for s in range(5):
            y, u, m = refs[s].process(pcm[s, : n_in[s] * ch], int(cap[s]))
            assert (u, m) == (int(used[s]), int(made[s])), (k, s, n_in[s], cap[s])
            assert np.array_equal(y, out[s, : m * ch]), (k, s)
    b.close()


def test_checkpoint_resume_roundtrip():
    ch, i, o, q = 2, 44100, 48000, 7
    a = StreamBatch(4, ch, i, o, q)
    x = synth_pcm(4, ch, 3000, i, seed=21)
    a.process(x[:, : 1000 * ch], 1000, 1200)
    snap = [a.get_state(s) for s in range(4)]
    out_a, _, made_a = a.process(x[:, 1000 * ch:], 2000, 2400)
    b = StreamBatch(4, ch, i, o, q)
    for s in range(4):
        b.set_state(s, snap[s][0], snap[s][1], snap[s][3])
    out_b, _, made_b = b.process(x[:, 1000 * ch:], 2000, 2400)
    assert np.array_equal(made_a, made_b) and np.array_equal(out_a, out_b)
    a.close()
    b.close()


def test_pipelined_submit_wait_equals_sync():
    """spxb_batch_submit / wait with pinned buffers: same bytes as the synchronous call"""
    L = lib()
    S, ch, i, o, q, n, cap = 64, 2, 44100, 48000, 7, 882, 960
    a, b = StreamBatch(S, ch, i, o, q), StreamBatch(S, ch, i, o, q)
    steps = 6
    x = [synth_pcm(S, ch, n, i, seed=31, start_frame=k * n) for k in range(steps)]
    want = [a.process(x[k], n, cap)[0] for k in range(steps)]
    in_bytes, out_bytes = S * n * ch * 2, S * cap * ch * 2
    hin = [L.spxb_host_alloc(in_bytes) for _ in range(steps)]
    hout = [L.spxb_host_alloc(out_bytes) for _ in range(steps)]
    tickets = []
    for k in range(steps):
        C.memmove(hin[k], x[k].ctypes.data, in_bytes)
        nin = np.full(S, n, np.uint32)
        nout = np.full(S, cap, np.uint32)
        t = C.c_uint64(0)
        assert L.spxb_batch_submit(b._h, hin[k], n, nin.ctypes.data, hout[k], cap, nout.ctypes.data, C.byref(t)) == 0
        assert np.all(nout == cap) and np.all(nin == n)
        tickets.append(t.value)
        if k >= 2:
            assert L.spxb_batch_wait(b._h, tickets[k - 2]) == 0
    for k in range(steps):
        assert L.spxb_batch_wait(b._h, tickets[k]) == 0
        got = np.ctypeslib.as_array(C.cast(hout[k], C.POINTER(C.c_int16)), shape=(S, cap * ch))
        assert np.array_equal(got, want[k]), k
    for p in hin + hout:
        L.spxb_host_free(p)
    a.close()
    b.close()


def test_device_pointer_entry_matches_host_entry():
    torch = pytest.importorskip("torch")
    L = lib()
    S, ch, i, o, q, n, cap = 130, 2, 44100, 48000, 7, 882, 960
    a, b = StreamBatch(S, ch, i, o, q), StreamBatch(S, ch, i, o, q)
    ts = torch.cuda.Stream()
    assert L.spxb_batch_set_stream(b._h, C.c_void_p(ts.cuda_stream)) == 0
    torch.cuda.set_stream(ts)
    for k in range(4):
        x = synth_pcm(S, ch, n, i, seed=41, start_frame=k * n)
        want, _, _ = a.process(x, n, cap)
        d_in = torch.from_numpy(x).cuda()
        d_out = torch.zeros((S, cap * ch), dtype=torch.int16, device="cuda")
        used, made = C.c_uint32(), C.c_uint32()
        assert L.spxb_batch_process_device_uniform(b._h, d_in.data_ptr(), n, n, d_out.data_ptr(), cap, cap,
                                                   C.byref(used), C.byref(made)) == 0
        ts.synchronize()  # the kernel ran on `ts`, not on the batch's own stream
        assert (used.value, made.value) == (n, cap)
        assert b.last_kernel() == a.last_kernel() == KERNEL_TENSOR  # odd row stride still on the tensor cores
        assert np.array_equal(d_out.cpu().numpy(), want)
    a.close()
    b.close()


def test_speex_c_api_getters_and_skip_zeros():
    L = lib()
    err = C.c_int(-1)
    st = L.speex_resampler_init(2, 44100, 48000, 7, C.byref(err))
    assert st and err.value == 0
    a, b = C.c_uint32(), C.c_uint32()
    L.speex_resampler_get_rate(st, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (44100, 48000)
    L.speex_resampler_get_ratio(st, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (147, 160)
    assert L.speex_resampler_get_input_latency(st) == 64
    assert L.speex_resampler_get_output_latency(st) == (64 * 160 + 73) // 147
    # skip_zeros (resample.c:1200): the first output then already sees the filter centre
    assert L.speex_resampler_skip_zeros(st) == 0
    x = synth_pcm(1, 2, 882, 44100, seed=51)[0]
    n_in, n_out = C.c_uint32(882), C.c_uint32(2000)
    out = np.zeros(4000, np.int16)
    assert L.speex_resampler_process_interleaved_int(st, x.ctypes.data, C.byref(n_in), out.ctypes.data, C.byref(n_out)) == 0
    pl = _lib.CallPlan()
    L.spxb_plan_call(44100, 48000, 64, 0, 882, 2000, C.byref(pl))
    assert (n_in.value, n_out.value) == (pl.consumed, pl.n_out)
    L.speex_resampler_destroy(st)


@pytest.mark.parametrize("kernel,shape", [
    (KERNEL_TENSOR, (70, 2, 44100, 48000, 7, 882, 960)),
    (KERNEL_TENSOR, (200, 1, 48000, 16000, 10, 960, 320)),    # long filter: the persistent kernel, its tensor maps captured
    (KERNEL_TENSOR, (150, 2, 96000, 48000, 10, 1920, 960)),   # ... stereo, several tiles per CTA
    (KERNEL_STRICT, (70, 2, 44100, 48000, 7, 882, 960)),
], ids=["tensor", "tensor_long_mono", "tensor_long_stereo", "strict"])
def test_device_ring_graph_replay_equals_single_hops(kernel, shape):
    """spxb_batch_process_device_ring captures a repeating hop sequence into a CUDA graph (second
    unchanged sighting) and replays it afterwards. Every round -- launch by launch, captured,
    replayed -- must leave the same outputs and the same stream state as single uniform hops."""
    torch = pytest.importorskip("torch")
    L = lib()
    S, ch, i, o, q, n, cap = shape
    ring, steps, rounds = 4, 8, 5
    a, b = StreamBatch(S, ch, i, o, q), StreamBatch(S, ch, i, o, q)
    a.set_kernel(kernel)
    b.set_kernel(kernel)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    for h in (a, b):
        assert L.spxb_batch_set_stream(h._h, C.c_void_p(ts.cuda_stream)) == 0
    x = np.stack([synth_pcm(S, ch, n, i, seed=51, start_frame=k * n) for k in range(ring)])
    d_in = torch.from_numpy(x).cuda()
    d_ring = torch.zeros((ring, S, cap * ch), dtype=torch.int16, device="cuda")
    d_one = torch.zeros((ring, S, cap * ch), dtype=torch.int16, device="cuda")
    in_slot, out_slot = S * n * ch, S * cap * ch
    for r in range(rounds):
        first = r * steps
        assert L.spxb_batch_process_device_ring(b._h, d_in.data_ptr(), n, in_slot, d_ring.data_ptr(), cap, out_slot,
                                                ring, n, cap, first, steps) == 0, _lib.last_error()
        for k in range(first, first + steps):
            sl = k % ring
            assert L.spxb_batch_process_device_uniform(a._h, d_in[sl].data_ptr(), n, n, d_one[sl].data_ptr(), cap,
                                                       cap, None, None) == 0
        torch.cuda.synchronize()
        assert torch.equal(d_ring, d_one), r
        assert b.counters().kernel_launches == a.counters().kernel_launches
        sa, sb = a.get_state(3), b.get_state(3)
        assert sa[:3] == sb[:3] and np.array_equal(sa[3], sb[3]), r
        d_ring.zero_()
        d_one.zero_()
    a.close()
    b.close()


# ---- float entry ---------------------------------------------------------
from cases import F32_CALLS, F32_ROWS, f32_input  # noqa: E402

VEC_F32 = np.load(os.path.join(os.path.dirname(__file__), "golden", "vectors_f32.npz"))


@pytest.mark.parametrize("row", F32_ROWS, ids=lambda r: case_id(MATRIX[r]))
def test_c_api_float_entry_matches_golden_bit_exact(row):
    """speex_resampler_process_interleaved_float and ..._int mixed on ONE state through the C
    ABI, against vectors from the reference's own float build: every f32 bit and every length,
    capacity-bound calls included (float entry block walk, resample.c:927-963)."""
    L = lib()
    ch, i, o, q, _ = MATRIX[row]
    key = case_id(MATRIX[row])
    err = C.c_int(0)
    st = L.speex_resampler_init(ch, i, o, q, C.byref(err))
    assert st and err.value == 0
    try:
        for k, (kind, n, cap) in enumerate(F32_CALLS):
            x = f32_input(row, k, kind, n, ch, i)
            xin = x if x.size else np.zeros(1, x.dtype)
            out = np.zeros(max(cap * ch, 1), np.float32 if kind == "f" else np.int16)
            n_in, n_out = C.c_uint32(n), C.c_uint32(cap)
            fn = L.speex_resampler_process_interleaved_float if kind == "f" else L.speex_resampler_process_interleaved_int
            assert fn(st, xin.ctypes.data, C.byref(n_in), out.ctypes.data, C.byref(n_out)) == 0, _lib.last_error()
            assert [n_in.value, n_out.value] == VEC_F32[f"{key}/call{k}/lens"].tolist(), (key, k)
            want = VEC_F32[f"{key}/call{k}/out"]
            got = out[: n_out.value * ch]
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (key, k)
    finally:
        L.speex_resampler_destroy(st)


def test_float_batch_matches_oracle_per_stream():
    """spxb_batch_process_f32 over a ragged batch: every stream equals its own oracle float
    state bit for bit, int16 calls interleaved on the same (float-history) batch"""
    ch, i, o, q, S = 2, 44100, 48000, 7, 9
    b = StreamBatch(S, ch, i, o, q, sample_format="f32")
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(S)]
    rng = np.random.default_rng(3)
    for k in range(8):
        n_in = rng.choice([0, 1, 160, 441, 500], size=S).astype(np.uint32)
        cap = rng.choice([0, 37, 480, 600], size=S).astype(np.uint32)
        pcm = synth_pcm(S, ch, 500, i, seed=70 + k)
        if k % 3 == 2:  # an int16 call on the float-history batch
            out, used, made = b.process(pcm, n_in, cap)
            for s in range(S):
                y, u, m = refs[s].process(pcm[s, : n_in[s] * ch], int(cap[s]))
                assert (u, m) == (int(used[s]), int(made[s])), (k, s)
                assert np.array_equal(y, out[s, : m * ch]), (k, s)
        else:
            x = (pcm.astype(np.float32) * np.float32(1.0 / 32768.0)).astype(np.float32)
            out, used, made = b.process_f32(x, n_in, cap)
            for s in range(S):
                y, u, m = refs[s].process_float(x[s, : n_in[s] * ch], int(cap[s]))
                assert (u, m) == (int(used[s]), int(made[s])), (k, s)
                assert np.array_equal(y.view(np.uint32), out[s, : m * ch].view(np.uint32)), (k, s)
    # the float view of the state round-trips
    L = lib()
    info = b.filter_info()
    hist = np.zeros((info.filt_len - 1) * ch, np.float32)
    ls, fr, mg = C.c_int32(), C.c_uint32(), C.c_uint32()
    assert L.spxb_batch_get_state_f32(b._h, 4, C.byref(ls), C.byref(fr), C.byref(mg), hist.ctypes.data) == 0
    rl, rf, rh = refs[4].state(0)
    assert (ls.value, fr.value) == (rl, rf)
    assert np.array_equal(hist[0::ch].view(np.uint32), rh.view(np.uint32))
    with pytest.raises(RuntimeError):
        StreamBatch(2, ch, i, o, q).process_f32(np.zeros((2, 8), np.float32), 4, 8)
    b.close()


@pytest.mark.skipif(not O.have_wasm(), reason="oracle/_ref/libspeex_wasm.so not present")
@pytest.mark.parametrize("c", [MATRIX[k] for k in (0, 2, 3, 4, 17, 18)], ids=case_id)
def test_gpu_matches_the_shipped_wasm_module(c):
    """north star: "output must match the reference WASM resampler". The module embedded in the
    reference's src/speex_wasm.js, executed through oracle/wasm2c_lite.py's translation: the strict
    kernel is bit-exact against it, the tensor kernel within 1 LSB and above 90 dB SNR."""
    ch, i, o, q, _ = c
    S, n = 3, 882
    cap = int(np.ceil(n * o / i)) + 1
    strict, fast = StreamBatch(S, ch, i, o, q), StreamBatch(S, ch, i, o, q)
    strict.set_kernel(KERNEL_STRICT)
    refs = [O.WasmResampler(ch, i, o, q) for _ in range(S)]
    for k in range(4):
        pcm = synth_pcm(S, ch, n, i, seed=500 + k, start_frame=k * n)
        ys, us, ms = strict.process(pcm, n, cap)
        yf, uf, mf = fast.process(pcm, n, cap)
        for s in range(S):
            y, u, m = refs[s].process(pcm[s], cap)
            assert (u, m) == (int(us[s]), int(ms[s])) == (int(uf[s]), int(mf[s]))
            assert np.array_equal(y, ys[s, : m * ch]), (k, s)
            d = y.astype(np.int32) - yf[s, : m * ch].astype(np.int32)
            assert np.abs(d).max() <= 1, (k, s)
            if k:
                assert O.snr_db(y, yf[s, : m * ch]) >= 90.0
    strict.close()
    fast.close()


def test_process_chunks_groups_mixed_configurations():
    """processChunks over resamplers of different (channels, rates, quality): each configuration
    becomes one device batch; results equal per-stream processChunk, in the caller's order"""
    cfgs = [(2, 44100, 48000, 7), (1, 48000, 16000, 10), (2, 44100, 48000, 7), (1, 24000, 48000, 5),
            (1, 48000, 16000, 10), (2, 44100, 48000, 7)]
    rs = [SpeexResampler(*c) for c in cfgs]
    for r in rs:
        r.kernel = KERNEL_STRICT
    refs = [O.OracleResampler(*c) for c in cfgs]
    for k in range(3):
        chunks = [synth_pcm(1, c[0], 480 + 7 * s, c[1], seed=800 + 10 * k + s)[0].tobytes()
                  for s, c in enumerate(cfgs)]
        got = SpeexResampler.processChunks(rs, chunks)
        for s in range(len(cfgs)):
            assert got[s] == refs[s].processChunk(chunks[s]), (k, s)


def test_node_addon_executes_like_the_oracle():
    """bindings/node/src/addon.c EXECUTED (through the in-process N-API stand-in, no Node in this
    image) with the calls index.ts makes -- init, process, batchCreate/Process/Adopt, the two error
    paths, an empty chunk -- must return byte for byte what the CPU ORACLE returns for the same
    inputs through the wrapper's capacity rule (the single stream runs the bit-exact kernel by
    default, the batches are switched to it with batchSetKernel)."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bindings", "node", "test",
                       "addon_selftest")
    if not os.path.exists(exe):
        pytest.skip("addon_selftest not built (python -c 'import __graft_entry__ as g; g.build()')")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = dict((ln.split()[0], ln.split(None, 1)[1]) for ln in r.stdout.strip().split("\n") if " " in ln)
    assert lines["init_error"] == "1 Invalid argument."
    assert lines["align_error"] == "1 Chunk length should be a multiple of channels * 2 bytes"
    assert lines["empty"] == "0"

    ch, i, o, q = 2, 44100, 48000, 7

    def pcm(s, first_frame, frames):
        k = (np.arange(frames * ch, dtype=np.uint64) + first_frame * ch + 1) * 2654435761 + s * 40503
        v = (k & 0xFFFFFFFF).astype(np.uint32)
        v ^= v >> np.uint32(15)
        return ((v & 0x3FFF).astype(np.int32) - 8192).astype(np.int16).tobytes()

    def line(y):
        return f"{len(y) // (ch * 2)} {O.fnv1a64(y)}"

    single = O.OracleResampler(ch, i, o, q)
    pos = 0
    for k, n in enumerate((882, 441, 7, 882)):
        assert lines[f"single{k}"] == line(single.processChunk(pcm(0, pos, n))), k
        pos += n
    assert single.processChunk(b"") == b""
    rs = [O.OracleResampler(ch, i, o, q) for _ in range(8)]
    posb = [0] * 8
    for tag, frames in (("batchA", lambda s: 882), ("batchB", lambda s: 300 + 41 * s)):
        for s in range(8):
            assert lines[f"{tag}{s}"] == line(rs[s].processChunk(pcm(10 + s, posb[s], frames(s)))), (tag, s)
            posb[s] += frames(s)
    # a fresh stream and the single stream (4 hops + an empty chunk in) continued inside a batch
    assert lines["adopt0"] == line(O.OracleResampler(ch, i, o, q).processChunk(pcm(30, 0, 882)))
    assert lines["adopt1"] == line(single.processChunk(pcm(0, pos, 882)))


def test_ragged_batch_cohorts_take_the_tensor_kernel():
    """A batch whose streams move in a few cohorts (same chunk sizes, different start times): each
    cohort of >= 32 streams gets a tensor-kernel launch over its stream-id list, the stragglers
    (and streams sitting a call out) one strict launch. Every stream must match its own oracle
    (<= 1 LSB; stragglers bit-exact) and carry its state across calls."""
    ch, i, o, q = 2, 44100, 48000, 5
    S = 150
    b = StreamBatch(S, ch, i, o, q)
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(S)]
    cohort = np.zeros(S, np.int64)
    cohort[40:110] = 1          # a second cohort, three calls behind the first
    cohort[110:145] = 2         # a third one with another chunk size
    cohort[145:] = 3            # five stragglers: ragged, strict kernel
    rng = np.random.default_rng(8)
    launches0 = b.counters().kernel_launches
    for k in range(7):
        n_in = np.zeros(S, np.uint32)
        n_in[cohort == 0] = 1000
        n_in[cohort == 1] = 1000 if k >= 3 else 0
        n_in[cohort == 2] = 777
        n_in[cohort == 3] = rng.integers(1, 900, size=5)
        cap = np.ceil(n_in * (o / i)).astype(np.uint32) + 1
        pcm = synth_pcm(S, ch, 1000, i, seed=600 + k)
        out, used, made = b.process(pcm, n_in, cap)
        assert b.last_kernel() == KERNEL_TENSOR
        for s in range(S):
            y, u, m = refs[s].process(pcm[s, : n_in[s] * ch], int(cap[s]))
            assert (u, m) == (int(used[s]), int(made[s])), (k, s)
            d = np.abs(y.astype(np.int32) - out[s, : m * ch].astype(np.int32))
            assert d.max(initial=0) <= (0 if cohort[s] == 3 else 1), (k, s, int(d.max(initial=0)))
    # cohort launches + one strict launch per call, never one launch per stream
    assert b.counters().kernel_launches - launches0 <= 7 * 4
    # a uniform call afterwards goes back to a single launch
    c0 = b.counters().kernel_launches
    b2 = StreamBatch(64, ch, i, o, q)
    b2.process(synth_pcm(64, ch, 500, i, seed=1), 500, 600)
    assert b2.counters().kernel_launches == 1 and c0 > 0
    b.close()
    b2.close()


def test_one_shot_long_chunk_on_the_tensor_kernel():
    """the reference's own test feeds a whole file to ONE processChunk (src/test.ts:31-32): a long
    one-shot call has far more output tiles than fit the kernel-parameter tile table, so the tensor
    kernel reads its tile table from HBM -- same results as chunked calls and as the oracle"""
    ch, i, o, q = 2, 44100, 48000, 7
    frames = 60000
    x = synth_pcm(2, ch, frames, i, seed=91)
    b = StreamBatch(2, ch, i, o, q)
    b.set_kernel(KERNEL_TENSOR)
    cap = int(np.ceil(frames * o / i))
    out, used, made = b.process(x, frames, cap)
    assert b.tensor_geometry()["tiles"] > 64
    for s in range(2):
        y, u, m = O.OracleResampler(ch, i, o, q).process(x[s], cap)
        assert (u, m) == (int(used[s]), int(made[s]))
        d = np.abs(y.astype(np.int32) - out[s, : m * ch].astype(np.int32))
        assert d.max() <= 1 and O.snr_db(y, out[s, : m * ch]) >= 90.0
    # the same stream in 20 ms hops ends in the same state and produces the same samples
    c = StreamBatch(2, ch, i, o, q)
    c.set_kernel(KERNEL_TENSOR)
    parts = []
    for k in range(0, frames, 882):
        n = min(882, frames - k)
        yo, _, mo = c.process(x[:, k * ch:(k + n) * ch], n, 962)
        parts.append(yo[:, : int(mo[0]) * ch])
    hops = np.concatenate(parts, axis=1)
    assert hops.shape[1] == int(made[0]) * ch
    assert np.abs(hops.astype(np.int32) - out[:, : hops.shape[1]].astype(np.int32)).max() <= 1
    assert b.get_state(1)[:2] == c.get_state(1)[:2]
    b.close()
    c.close()


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not present")
def test_c_api_filter_changes_before_the_first_sample_match_the_reference():
    """init_frac / set_quality / set_rate / set_rate_frac / skip_zeros before any sample has been
    resampled (resample.c:1107-1163 with update_filter's !started branch): same lengths and bytes
    as the reference's own build given the same call sequence (mid-stream changes have their own
    tests below)."""
    L, R = lib(), O._load_ref()
    R.speex_resampler_init_frac.restype = C.c_void_p
    R.speex_resampler_init_frac.argtypes = [C.c_uint32] * 5 + [C.c_int, C.POINTER(C.c_int)]
    for fn in (R.speex_resampler_set_quality, R.speex_resampler_set_rate, R.speex_resampler_set_rate_frac,
               R.speex_resampler_skip_zeros):
        fn.restype = C.c_int
    R.speex_resampler_set_quality.argtypes = [C.c_void_p, C.c_int]
    R.speex_resampler_set_rate.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    R.speex_resampler_set_rate_frac.argtypes = [C.c_void_p] + [C.c_uint32] * 4
    R.speex_resampler_skip_zeros.argtypes = [C.c_void_p]
    ch = 2
    err = C.c_int(0)
    ours = L.speex_resampler_init_frac(ch, 441, 480, 44100, 48000, 7, C.byref(err))
    ref = R.speex_resampler_init_frac(ch, 441, 480, 44100, 48000, 7, C.byref(err))
    assert ours and ref
    assert L.spxb_batch_set_kernel(L.spxb_resampler_batch(ours), KERNEL_STRICT) == 0
    for lib_, st in ((L, ours), (R, ref)):
        assert lib_.speex_resampler_set_quality(st, 3) == 0
        assert lib_.speex_resampler_skip_zeros(st) == 0          # last_sample = filt_len / 2 of the q3 filter
        assert lib_.speex_resampler_set_rate(st, 48000, 16000) == 0
        assert lib_.speex_resampler_set_quality(st, 10) == 0
        assert lib_.speex_resampler_set_rate_frac(st, 320, 147, 96000, 44100) == 0
        assert lib_.speex_resampler_set_quality(st, 11) == 3      # RESAMPLER_ERR_INVALID_ARG
    num, den = C.c_uint32(), C.c_uint32()
    L.speex_resampler_get_ratio(ours, C.byref(num), C.byref(den))
    assert (num.value, den.value) == (320, 147)
    x = synth_pcm(1, ch, 3000, 96000, seed=77)[0]
    pos = 0
    for n, cap in ((1000, 600), (7, 600), (1500, 300), (493, 600)):
        chunk = np.ascontiguousarray(x[pos * ch:(pos + n) * ch])
        pos += n
        res = []
        for lib_, st in ((L, ours), (R, ref)):
            out = np.zeros(cap * ch, np.int16)
            n_in, n_out = C.c_uint32(n), C.c_uint32(cap)
            assert lib_.speex_resampler_process_interleaved_int(st, chunk.ctypes.data, C.byref(n_in), out.ctypes.data,
                                                                C.byref(n_out)) == 0
            res.append((n_in.value, n_out.value, out[: n_out.value * ch].copy()))
        assert res[0][:2] == res[1][:2]
        assert np.array_equal(res[0][2], res[1][2])
    # the same reduced ratio under other numbers: no filter change on either side
    assert L.speex_resampler_set_rate_frac(ours, 640, 294, 96000, 44100) == 0
    assert R.speex_resampler_set_rate_frac(ref, 640, 294, 96000, 44100) == 0
    L.speex_resampler_destroy(ours)
    R.speex_resampler_destroy(ref)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("ch", [1, 2])
def test_c_api_mid_stream_rate_and_quality_changes_match_the_reference(ch):
    """Mid-stream filter changes that do not shorten the filter, call for call against the
    reference build: clock-drift correction while up-sampling (same filter length, samp_frac_num
    rescaled, resample.c:1131-1140), a quality increase and a deeper down-sampling (longer filter:
    history re-anchored behind zeros, last_sample advanced, resample.c:727-758)."""
    L, R = lib(), O._load_ref()
    R.speex_resampler_set_quality.restype = R.speex_resampler_set_rate.restype = C.c_int
    R.speex_resampler_set_quality.argtypes = [C.c_void_p, C.c_int]
    R.speex_resampler_set_rate.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    err = C.c_int(0)
    ours = L.speex_resampler_init(ch, 44100, 48000, 5, C.byref(err))
    ref = R.speex_resampler_init(ch, 44100, 48000, 5, C.byref(err))
    assert L.spxb_batch_set_kernel(L.spxb_resampler_batch(ours), KERNEL_STRICT) == 0
    x = synth_pcm(1, ch, 12000, 44100, seed=123)[0]
    pos = [0]

    def both(n, cap):
        chunk = np.ascontiguousarray(x[pos[0] * ch:(pos[0] + n) * ch])
        pos[0] += n
        res = []
        for lib_, st in ((L, ours), (R, ref)):
            out = np.zeros(cap * ch, np.int16)
            n_in, n_out = C.c_uint32(n), C.c_uint32(cap)
            assert lib_.speex_resampler_process_interleaved_int(st, chunk.ctypes.data, C.byref(n_in), out.ctypes.data,
                                                                C.byref(n_out)) == 0
            res.append((n_in.value, n_out.value, out[: n_out.value * ch].copy()))
        assert res[0][:2] == res[1][:2], (pos[0], res[0][:2], res[1][:2])
        assert np.array_equal(res[0][2], res[1][2]), pos[0]

    def change(fn, *args):
        a, b = getattr(L, fn)(ours, *args), getattr(R, fn)(ref, *args)
        assert a == b == 0, (fn, args, a, b, _lib.last_error())

    both(441, 600)
    both(333, 600)
    change("speex_resampler_set_rate", 44100, 48010)     # drift correction: same filter length
    both(441, 600)
    change("speex_resampler_set_rate", 44100, 47990)
    both(500, 300)                                        # capacity binds
    both(441, 600)
    change("speex_resampler_set_quality", 8)              # longer filter
    both(441, 600)
    both(7, 600)
    change("speex_resampler_set_rate", 44100, 22050)      # now down-sampling 2:1: longer again
    both(882, 600)
    both(441, 600)
    change("speex_resampler_set_quality", 10)
    both(441, 600)
    # shorter filters: the surplus history becomes "magic samples", resampled before the next input
    # (resample.c:759-776, :904-922); the reference's memory does not shrink, so its input block grows
    change("speex_resampler_set_rate", 44100, 44100)
    both(100, 7)                                          # capacity binds inside the magic block
    both(0, 50)                                           # nothing moves without input (resample.c:988)
    both(441, 30)
    both(441, 600)
    both(1000, 600)                                       # input block > 160 shows in a capacity-bound call
    change("speex_resampler_set_quality", 1)
    both(300, 600)
    both(441, 600)
    # further length changes while magic samples are still pending (resample.c:727-776 with
    # magic_samples != 0): shorter again, then longer -- first below, then above the "augmented" length
    change("speex_resampler_set_quality", 0)
    change("speex_resampler_set_quality", 4)
    both(50, 3)
    both(441, 600)
    change("speex_resampler_set_quality", 10)
    both(300, 40)
    change("speex_resampler_set_quality", 3)              # shorter with magic pending: the magic regions add up
    both(10, 2)
    change("speex_resampler_set_quality", 1)
    both(200, 600)
    both(441, 600)
    L.speex_resampler_destroy(ours)
    R.speex_resampler_destroy(ref)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not present")
def test_c_api_magic_samples_on_the_float_entry_match_the_reference():
    """float entry after a filter shortening (resample.c:940-941: the pending samples take all the
    capacity that is left, then the input follows): every f32 bit and length against the
    reference build"""
    L, R = lib(), O._load_ref()
    R.speex_resampler_set_quality.restype = C.c_int
    R.speex_resampler_set_quality.argtypes = [C.c_void_p, C.c_int]
    ch = 2
    err = C.c_int(0)
    ours = L.speex_resampler_init(ch, 48000, 32000, 9, C.byref(err))
    ref = R.speex_resampler_init(ch, 48000, 32000, 9, C.byref(err))
    x = (synth_pcm(1, ch, 6000, 48000, seed=321)[0].astype(np.float32) * np.float32(0.61)).astype(np.float32)
    pos = [0]

    def both(n, cap):
        chunk = np.ascontiguousarray(x[pos[0] * ch:(pos[0] + n) * ch])
        pos[0] += n
        res = []
        for lib_, st in ((L, ours), (R, ref)):
            out = np.zeros(max(cap * ch, 1), np.float32)
            n_in, n_out = C.c_uint32(n), C.c_uint32(cap)
            buf = chunk if chunk.size else np.zeros(1, np.float32)
            assert lib_.speex_resampler_process_interleaved_float(st, buf.ctypes.data, C.byref(n_in), out.ctypes.data,
                                                                  C.byref(n_out)) == 0, _lib.last_error()
            res.append((n_in.value, n_out.value, out[: n_out.value * ch].copy()))
        assert res[0][:2] == res[1][:2], (pos[0], res[0][:2], res[1][:2])
        assert np.array_equal(res[0][2].view(np.uint32), res[1][2].view(np.uint32)), pos[0]

    both(480, 400)
    both(333, 400)
    assert L.speex_resampler_set_quality(ours, 2) == 0 and R.speex_resampler_set_quality(ref, 2) == 0
    both(200, 5)      # capacity binds inside the magic block: no input is taken
    both(50, 0)       # no room at all: the pending samples the read position has passed are still consumed
    both(200, 400)
    both(7, 400)
    both(1000, 100)
    both(480, 400)
    L.speex_resampler_destroy(ours)
    R.speex_resampler_destroy(ref)


def test_batches_release_their_device_memory():
    """create / use / destroy in a loop (uniform calls, a ragged call with cohorts, a captured hop
    sequence, a float batch, a C-API state that changes filter): free device memory returns to
    where it started"""
    torch = pytest.importorskip("torch")
    L = lib()
    ch, i, o, q, S, n, cap = 2, 44100, 48000, 7, 96, 882, 962

    def one_round():
        b = StreamBatch(S, ch, i, o, q)
        pcm = synth_pcm(S, ch, n, i, seed=3)
        b.process(pcm, n, cap)
        b.process(pcm, np.where(np.arange(S) < 48, n, 500).astype(np.uint32), cap)     # two cohorts
        ts = torch.cuda.Stream()
        assert L.spxb_batch_set_stream(b._h, C.c_void_p(ts.cuda_stream)) == 0
        d_in = torch.zeros((2, S, n * ch), dtype=torch.int16, device="cuda")
        d_out = torch.zeros((2, S, cap * ch), dtype=torch.int16, device="cuda")
        c = StreamBatch(S, ch, i, o, q)
        assert L.spxb_batch_set_stream(c._h, C.c_void_p(ts.cuda_stream)) == 0
        for r in range(4):
            assert L.spxb_batch_process_device_ring(c._h, d_in.data_ptr(), n, S * n * ch, d_out.data_ptr(), cap,
                                                    S * cap * ch, 2, n, cap, r * 4, 4) == 0
        torch.cuda.synchronize()
        f = StreamBatch(4, ch, i, o, q, sample_format="f32")
        f.process_f32(np.zeros((4, n * ch), np.float32), n, cap)
        err = C.c_int(0)
        st = L.speex_resampler_init(ch, i, o, 10, C.byref(err))
        x = np.zeros(n * ch, np.int16)
        out = np.zeros(cap * ch, np.int16)
        a, bb = C.c_uint32(n), C.c_uint32(cap)
        L.speex_resampler_process_interleaved_int(st, x.ctypes.data, C.byref(a), out.ctypes.data, C.byref(bb))
        assert L.speex_resampler_set_quality(st, 3) == 0
        a, bb = C.c_uint32(n), C.c_uint32(cap)
        L.speex_resampler_process_interleaved_int(st, x.ctypes.data, C.byref(a), out.ctypes.data, C.byref(bb))
        L.speex_resampler_destroy(st)
        for h in (b, c, f):
            h.close()
        del d_in, d_out
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    one_round()                       # first round pays one-time costs (module load, pools)
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(12):
        one_round()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (8 << 20), (free0, free1)


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("ch", [1, 2, 3])
@pytest.mark.parametrize("kernel", [KERNEL_STRICT, KERNEL_TENSOR], ids=["strict", "tensor"])
def test_c_api_reset_mem_matches_the_reference(kernel, ch):
    """speex_resampler_reset_mem (resample.c:1208-1220) mid-stream, call for call against the
    reference's own build. The reference zeroes last_sample / samp_frac_num / magic_samples and the
    FIRST nb_channels * (filt_len - 1) floats of `mem`, whose channels lie mem_alloc_size apart: only
    channel 0's history is certainly cleared (a stereo stream keeps its right-channel history). That
    is reproduced, not corrected; a mono stream is indistinguishable from a new one afterwards."""
    if ch == 3 and kernel == KERNEL_TENSOR:
        pytest.skip("the tensor kernel serves mono and stereo")
    L, R = lib(), O._load_ref()
    R.speex_resampler_reset_mem.restype = C.c_int
    R.speex_resampler_reset_mem.argtypes = [C.c_void_p]
    i, o, q = 44100, 48000, 7
    err = C.c_int(0)
    ours = L.speex_resampler_init(ch, i, o, q, C.byref(err))
    ref = R.speex_resampler_init(ch, i, o, q, C.byref(err))
    fresh = O.OracleResampler(ch, i, o, q)
    assert L.spxb_batch_set_kernel(L.spxb_resampler_batch(ours), kernel) == 0
    x = synth_pcm(1, ch, 6000, i, seed=515)[0]
    pos = 0
    for k, (n, cap) in enumerate(((882, 962), (441, 482), (300, 100), ("reset", 0), (882, 962), (7, 9), (882, 962))):
        if n == "reset":
            assert L.speex_resampler_reset_mem(ours) == 0 and R.speex_resampler_reset_mem(ref) == 0
            ls, fr, mg = C.c_int32(1), C.c_uint32(1), C.c_uint32(1)
            hist = np.ones((127) * ch, np.int16)
            assert L.spxb_batch_get_state(L.spxb_resampler_batch(ours), 0, C.byref(ls), C.byref(fr), C.byref(mg),
                                          hist.ctypes.data) == 0
            assert (ls.value, fr.value, mg.value) == (0, 0, 0) and not hist.reshape(-1, ch)[:, 0].any()
            if ch == 2:  # mem_alloc_size 287 > 2 * 127: the right channel's history survives
                assert hist.reshape(-1, ch)[:, 1].any()
            after_reset = pos
            continue
        chunk = np.ascontiguousarray(x[pos * ch:(pos + n) * ch])
        pos += n
        res = []
        for lib_, st in ((L, ours), (R, ref)):
            out = np.zeros(cap * ch, np.int16)
            n_in, n_out = C.c_uint32(n), C.c_uint32(cap)
            assert lib_.speex_resampler_process_interleaved_int(st, chunk.ctypes.data, C.byref(n_in), out.ctypes.data,
                                                                C.byref(n_out)) == 0
            res.append((n_in.value, n_out.value, out[: n_out.value * ch].copy()))
        assert res[0][:2] == res[1][:2], k
        check_close(res[1][2], res[0][2], exact=kernel == KERNEL_STRICT, what=("reset_mem", k))
        if k > 3 and ch == 1:  # after the reset a mono stream is indistinguishable from a new one
            y, u, m = fresh.process(chunk, cap)
            assert (u, m) == res[1][:2] and np.array_equal(y, res[1][2])
    assert after_reset == 882 + 441 + 300
    L.speex_resampler_destroy(ours)
    R.speex_resampler_destroy(ref)


def test_batch_reset_restarts_every_stream():
    """spxb_batch_reset: all streams of a batch back to the state of a new resampler"""
    S, ch, i, o, q, n, cap = 40, 1, 48000, 16000, 10, 960, 322
    b = StreamBatch(S, ch, i, o, q)
    first = None
    for rnd in range(2):
        for k in range(3):
            out, used, made = b.process(synth_pcm(S, ch, n, i, seed=31, start_frame=k * n), n, cap)
            if rnd == 0 and k == 0:
                first = (out.copy(), used.copy(), made.copy())
        b.reset()
        assert b.get_state(S - 1)[:3] == (0, 0, 0) and not b.get_state(S - 1)[3].any()
    out, used, made = b.process(synth_pcm(S, ch, n, i, seed=31, start_frame=0), n, cap)
    assert np.array_equal(out, first[0]) and np.array_equal(used, first[1]) and np.array_equal(made, first[2])
    b.close()


def test_empty_chunks_return_empty_buffers():
    """src/index.ts:50-116 with an empty Buffer: the reference's loop does not run and an empty Buffer
    comes back; state untouched (used to be 'Invalid argument.' through N-API's NULL data)"""
    r = SpeexResampler(2, 44100, 48000, 7)
    ref = O.OracleResampler(2, 44100, 48000, 7)
    x = synth_pcm(1, 2, 882 * 2, 44100, seed=3)[0]
    assert r.processChunk(b"") == b"" == ref.processChunk(b"")
    assert r.processChunk(x[: 882 * 2]) == ref.processChunk(x[: 882 * 2])
    assert r.processChunk(b"") == b""
    assert r.processChunk(x[882 * 2:]) == ref.processChunk(x[882 * 2:])
    L = lib()
    n_in, n_out = C.c_uint32(0), C.c_uint32(10)
    assert L.speex_resampler_process_interleaved_int(r._resamplerPtr, None, C.byref(n_in), None, C.byref(n_out)) == 0
    assert (n_in.value, n_out.value) == (0, 0)
    t = SpeexResamplerTransform(2, 44100, 48000, 7)
    assert t.transform(b"\x01") == b""  # shorter than one frame: carried, nothing to resample yet
    r.destroy()


@pytest.mark.parametrize("shape", [
    ("C3x", 1300, 2, 44100, 48000, 7, 882, 3),    # 21 groups x 9 tiles on <= 148 CTAs: CTAs walk two tiles
    ("C5x", 700, 2, 96000, 44100, 10, 1920, 2),   # packed tile of 172 KB, 3 PCM stages, tap tile changes mid-CTA
    ("C4x", 600, 1, 48000, 16000, 10, 960, 2),    # mono, direct filter N = 768
    ("q0", 200, 2, 44100, 48000, 0, 441, 2),      # shortest filter: a single stage per tile
], ids=lambda s: s[0])
def test_persistent_kernel_parity(shape):
    """The persistent, TMA-fed tensor kernel (csrc/kernels_umma2.cu) is what long filters run by default
    (8 or more 64-frame stages per tile). Forced on (SPXB_UMMA_RESIDENT=1) it must meet the
    bar on every shape, short filters and single-stage tiles included -- <= 1 LSB, >= 90 dB, lengths and
    device state equal to the oracle's."""
    import subprocess
    import sys
    env = dict(os.environ, SPXB_UMMA_RESIDENT="1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "resident_check.py")] + [str(v) for v in shape],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "resident ok" in r.stdout and "'stages'" in r.stdout


@pytest.mark.parametrize("ch,i,o", [(2, 96000, 44100), (1, 48000, 16000)], ids=["stereo", "mono"])
def test_persistent_kernel_takes_calls_shorter_than_the_row_pitch(ch, i, o):
    """The persistent kernel reads PCM through tensor maps whose extent is the call's length, not the row
    pitch: calls of varying length inside rows of a fixed (16-byte aligned) pitch -- odd lengths, a box
    that ends inside the input, boxes entirely past its end, the straddle of history and input moving
    from call to call -- must match the oracle, lengths, samples and state."""
    q, pitch, S = 10, 1920, 300
    b = StreamBatch(S, ch, i, o, q)
    pick = [0, 1, 63, 64, 150, S - 1]
    refs = {s: O.OracleResampler(ch, i, o, q) for s in pick}
    want, got = {s: [] for s in pick}, {s: [] for s in pick}
    pos = 0
    for k, n in enumerate((1001, 1920, 333, 64, 1500, 17, 1919)):
        pcm = np.zeros((S, pitch * ch), np.int16)
        pcm[:, : n * ch] = synth_pcm(S, ch, n, i, seed=0xABBA, start_frame=pos)
        pcm[:, n * ch:] = 12345      # bytes past the call's length must not be read
        pos += n
        cap = -(-n * o // i) + 1
        out, used, made = b.process(pcm, n, cap)
        assert b.last_kernel() == KERNEL_TENSOR
        for s in pick:
            y, u_, m = refs[s].process(pcm[s, : n * ch], cap)
            assert (u_, m) == (int(used[s]), int(made[s])), (k, s)
            d = np.abs(y.astype(np.int32) - out[s, : m * ch].astype(np.int32))
            assert d.max(initial=0) <= 1, (k, s, int(d.max()))        # per hop: 1 LSB
            want[s].append(y)
            got[s].append(out[s, : m * ch].copy())
    for s in pick:  # the SNR bar over the stream (a 16-sample hop with one LSB off is below 90 dB by itself)
        assert O.snr_db(np.concatenate(want[s]), np.concatenate(got[s])) >= 90.0, s
    ls, fr, mg, hist = b.get_state(pick[-1])
    rls, rfr, rhist = refs[pick[-1]].state(0)
    assert (ls, fr) == (rls, rfr) and np.array_equal(hist.reshape(-1, ch)[:, 0].astype(np.float32), rhist)
    b.close()


@pytest.mark.parametrize("S,ch,i,o,q,n", [
    (4096, 1, 48000, 16000, 10, 960),     # C4: mono, 16 stages per tile, one tile per CTA
    (4096, 2, 96000, 48000, 10, 1920),    # stereo long filter with 16-byte aligned output rows, several tiles per CTA
], ids=["mono", "stereo"])
def test_persistent_kernel_is_deterministic_under_repetition(S, ch, i, o, q, n):
    """A race that corrupted one row of one tile in ~2 % of the launches (a ring slot released before
    the loads from it had completed) passed every oracle comparison on a handful of calls: two batches
    fed the same input (the second one permuted) must agree bit for bit on EVERY one of many calls."""
    cap = -(-n * o // i)
    a_, b_ = StreamBatch(S, ch, i, o, q), StreamBatch(S, ch, i, o, q)
    perm = np.random.default_rng(5).permutation(S)
    base = synth_pcm(128, ch, n * 4, i, seed=0xD1CE)
    sel = np.resize(np.arange(128), S)
    for k in range(48):
        pcm = np.ascontiguousarray(base[np.roll(sel, 5 * k), (k % 4) * n * ch:((k % 4) + 1) * n * ch])
        out, used, made = a_.process(pcm, n, cap)
        out2, _, _ = b_.process(pcm[perm], n, cap)
        assert a_.last_kernel() == KERNEL_TENSOR and b_.last_kernel() == KERNEL_TENSOR
        assert np.array_equal(out2, out[perm]), (k, np.argwhere(out2 != out[perm])[:4])
    a_.close()
    b_.close()


# The variant with the byte planes in shared memory (SPXB_UMMA_ATMEM=0) is an experiment kept for the record:
# its converters still follow every phase of barriers that both groups share, which tests/test_ring_protocol.py
# shows can deadlock when a group falls a slot cycle behind. It has passed these cases every time they ran,
# but a hang costs ten minutes per case, so they run only on request (SPXB_TEST_INPLACE=1).
INPLACE = [
    {"SPXB_UMMA_RESIDENT": "1", "SPXB_UMMA_ATMEM": "0"},                                             # planes converted in place in shared memory
    {"SPXB_UMMA_RESIDENT": "1", "SPXB_UMMA_ATMEM": "0", "SPXB_UMMA2_XSTAGES": "3"},                  # ... with an odd ring: slots change converter group
    {"SPXB_UMMA_RESIDENT": "1", "SPXB_UMMA_ATMEM": "0", "SPXB_UMMA_NT": "64", "SPXB_UMMA_DENSE": "1"},  # ... two accumulator sets, dedicated epilogue warps
] if os.environ.get("SPXB_TEST_INPLACE") else []


@pytest.mark.parametrize("env", [
    {"SPXB_UMMA_RESIDENT": "0"},                                                   # the one-tile-per-CTA kernel on a long filter
    {"SPXB_UMMA_RESIDENT": "1", "SPXB_UMMA_NT": "64", "SPXB_UMMA_DENSE": "1"},     # planes in TMEM, narrow dense tiles
    {"SPXB_UMMA_RESIDENT": "1", "SPXB_UMMA_NT": "96"},                             # planes in TMEM, 4 A slots, even raw ring
    {"SPXB_UMMA_RESIDENT": "1", "SPXB_UMMA_NT": "112"},                            # planes in TMEM, 2 A slots, odd raw ring
] + INPLACE, ids=lambda e: "_".join(f"{k[5:].lower()}{v}" for k, v in e.items()))
def test_long_filter_on_either_tensor_kernel(env):
    """The long-filter shapes through the kernel they do NOT get by default, and through the
    persistent kernel's other configurations (byte planes in tensor memory or in shared memory, ring
    sizes, one or two accumulator sets): same bar."""
    import subprocess
    import sys
    for shape in (("C5x", 700, 2, 96000, 44100, 10, 1920, 3), ("C4x", 600, 1, 48000, 16000, 10, 960, 2)):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "resident_check.py")] + [str(v) for v in shape],
                           capture_output=True, text=True, timeout=600, env=dict(os.environ, PYTHONPATH=ROOT, **env), cwd=ROOT)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        assert "resident ok" in r.stdout


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not present")
def test_c_api_per_channel_entries_and_strides_match_the_reference():
    """speex_resampler_process_int / _process_float on single channels with input / output strides
    (resample.c:925-1036, :1170-1188), mixed with the interleaved entries on the same state, call for
    call against the reference's own build: lengths, every written sample bit for bit, and the samples
    BETWEEN the strided ones untouched. Channels are advanced by different amounts, so their positions
    diverge -- the reference keeps last_sample / samp_frac_num / mem per channel."""
    L, R = lib(), O._load_ref()
    for fn in ("speex_resampler_process_int", "speex_resampler_process_float"):
        getattr(R, fn).restype = C.c_int
        getattr(R, fn).argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p,
                                   C.POINTER(C.c_uint32)]
    for fn in ("speex_resampler_set_input_stride", "speex_resampler_set_output_stride"):
        getattr(R, fn).argtypes = [C.c_void_p, C.c_uint32]
    R.speex_resampler_get_input_stride.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    ch, i, o, q = 3, 44100, 48000, 5
    err = C.c_int(0)
    ours = L.speex_resampler_init(ch, i, o, q, C.byref(err))
    ref = R.speex_resampler_init(ch, i, o, q, C.byref(err))
    x = synth_pcm(1, ch, 9000, i, seed=4242)[0]
    pos = [0]

    def interleaved(n, cap, kind="i"):
        dt = np.int16 if kind == "i" else np.float32
        chunk = np.ascontiguousarray(x[pos[0] * ch:(pos[0] + n) * ch].astype(dt))
        pos[0] += n
        res = []
        for lib_, st in ((L, ours), (R, ref)):
            out = np.full(cap * ch + 5, 77, dt)
            n_in, n_out = C.c_uint32(n), C.c_uint32(cap)
            fn = lib_.speex_resampler_process_interleaved_int if kind == "i" else lib_.speex_resampler_process_interleaved_float
            assert fn(st, chunk.ctypes.data, C.byref(n_in), out.ctypes.data, C.byref(n_out)) == 0, _lib.last_error()
            res.append((n_in.value, n_out.value, out.copy()))
        assert res[0][:2] == res[1][:2], ("interleaved", pos[0], res[0][:2], res[1][:2])
        assert np.array_equal(res[0][2].view(np.uint8), res[1][2].view(np.uint8)), ("interleaved", pos[0])

    def channel(c, n, cap, istride, ostride, kind="i", null_in=False):
        dt = np.int16 if kind == "i" else np.float32
        src = np.full(n * istride + 3, -5, dt)
        src[: n * istride: istride] = x[(pos[0] * ch + c): (pos[0] + n) * ch: ch].astype(dt)
        res = []
        for lib_, st in ((L, ours), (R, ref)):
            lib_.speex_resampler_set_input_stride(st, istride)
            lib_.speex_resampler_set_output_stride(st, ostride)
            out = np.full(cap * ostride + 7, 99, dt)
            n_in, n_out = C.c_uint32(n), C.c_uint32(cap)
            fn = lib_.speex_resampler_process_int if kind == "i" else lib_.speex_resampler_process_float
            assert fn(st, c, None if null_in else src.ctypes.data, C.byref(n_in), out.ctypes.data, C.byref(n_out)) == 0, \
                _lib.last_error()
            res.append((n_in.value, n_out.value, out.copy()))
        assert res[0][:2] == res[1][:2], ("channel", c, pos[0], res[0][:2], res[1][:2])
        # written samples equal, everything between and after them still holds the fill value
        assert np.array_equal(res[0][2].view(np.uint8), res[1][2].view(np.uint8)), ("channel", c, pos[0])
        return res[0][0]

    interleaved(441, 600)
    interleaved(300, 100)                      # capacity binds
    s = C.c_uint32(0)
    L.speex_resampler_get_input_stride(ours, C.byref(s))
    assert s.value == 1                        # resample.c:836: strides start at 1
    used = [channel(c, 200 + 30 * c, 600, 1 + c, 3 - c) for c in range(ch)]   # channels advance differently
    assert used == [200, 230, 260]
    channel(1, 100, 7, 2, 2)                   # capacity binds on one channel only
    channel(0, 50, 600, 1, 1, null_in=True)    # in == NULL: zeros (resample.c:1007-1010)
    interleaved(441, 600)                      # interleaved entry on the planar state: channels keep their own positions
    channel(2, 160, 600, 4, 1, kind="f")       # float per-channel entry (the state turns float)
    interleaved(333, 600, kind="f")
    interleaved(200, 30)
    L.speex_resampler_get_input_stride(ours, C.byref(s))
    assert s.value == 4                        # the interleaved entries restore the caller's strides (resample.c:1080)
    L.speex_resampler_destroy(ours)
    R.speex_resampler_destroy(ref)


# ---------------------------------------------------------------------------
# formats either side of the path, on the device
# ---------------------------------------------------------------------------
def test_scaled_float_pcm_is_converted_inside_the_kernel():
    """float PCM (+-1.0 full scale) in and out of an int16 batch: converted on load / store inside the
    strict kernel. Must equal, bit for bit, the oracle fed the converted int16 samples, scaled back;
    saturating inputs (|x| >= 1) and values between two int16 steps included."""
    S, ch, i, o, q, n = 37, 2, 44100, 48000, 7, 882
    cap = 962
    b = StreamBatch(S, ch, i, o, q)
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(S)]
    rng = np.random.default_rng(12)
    for k in range(3):
        pcm = synth_pcm(S, ch, n, i, seed=0xF10A, start_frame=k * n)
        x = pcm.astype(np.float32) / np.float32(32768.0)
        x += rng.uniform(-0.49, 0.49, size=x.shape).astype(np.float32) / np.float32(32768.0)  # off the int16 grid
        x[3, :40] = 1.5       # saturates high
        x[4, :40] = -1.25     # saturates low
        want_in = np.clip(np.rint(x.astype(np.float32) * np.float32(32768.0)), -32768, 32767).astype(np.int16)
        out, used, made = b.process_pcm_f32(x, n, cap)
        assert b.last_kernel() == KERNEL_STRICT
        for s in range(S):
            y, u, m = refs[s].process(want_in[s], cap)
            assert (u, m) == (int(used[s]), int(made[s]))
            assert np.array_equal(out[s, : m * ch], y.astype(np.float32) / np.float32(32768.0)), (k, s)
    # the same batch keeps serving int16 calls from the same state
    pcm = synth_pcm(S, ch, n, i, seed=0xF10A, start_frame=3 * n)
    out, used, made = b.process(pcm, n, cap)
    y, u, m = refs[5].process(pcm[5], cap)
    assert np.abs(out[5, : m * ch].astype(np.int32) - y.astype(np.int32)).max() <= 1
    b.close()


def test_planar_layout_is_a_mono_batch_on_the_tensor_kernel():
    """planar PCM ([stream][channel][frame]) needs no kernel of its own: channel c of stream s is
    series s*channels + c of a MONO batch, and the tensor kernel's rows are series anyway. Against
    the oracle run on the interleaved form of the same audio, channel by channel."""
    S, ch, i, o, q, n = 48, 2, 44100, 48000, 7, 882
    cap = 962
    planar = StreamBatch(S * ch, 1, i, o, q)
    planar.set_kernel(KERNEL_TENSOR)
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(S)]
    for k in range(3):
        inter = synth_pcm(S, ch, n, i, seed=0x91A, start_frame=k * n)                 # [S, n*ch] interleaved
        rows = np.ascontiguousarray(inter.reshape(S, n, ch).transpose(0, 2, 1)).reshape(S * ch, n)  # [S*ch, n] planar
        out, used, made = planar.process(rows, n, cap)
        assert planar.last_kernel() == KERNEL_TENSOR
        for s in (0, 7, S - 1):
            y, u, m = refs[s].process(inter[s], cap)
            for c in range(ch):
                got = out[s * ch + c, :m]
                assert (u, m) == (int(used[s * ch + c]), int(made[s * ch + c]))
                d = np.abs(got.astype(np.int32) - y.reshape(m, ch)[:, c].astype(np.int32))
                assert d.max() <= 1 and O.snr_db(y.reshape(m, ch)[:, c], got) >= 90.0
    planar.close()


@pytest.mark.skipif(O.fixture_path("44100hz_test.pcm") is None, reason="oracle/_ref/resources absent")
@pytest.mark.parametrize("kernel", [KERNEL_TENSOR, KERNEL_STRICT], ids=["tensor", "strict"])
def test_wav_payload_is_read_in_place_past_the_riff_header(kernel):
    """The reference's fixtures are RIFF/WAVE files; formats.wav_pcm finds the payload 44 bytes in.
    The kernels read it IN PLACE -- device-visible rows that start 44 bytes into a buffer are only
    4-byte aligned, which takes the tensor kernel's unaligned load path -- 20 ms hop after hop."""
    from node_speex_resampler_b200 import wav_pcm
    L = lib()
    blob = open(O.fixture_path("44100hz_test.pcm"), "rb").read()
    w = wav_pcm(blob)
    off = len(blob) - len(w.data)
    assert off == 44 and (w.channels, w.sample_rate) == (2, 44100)
    ch, i, o, q, n, cap = 2, 44100, 48000, 7, 882, 960
    hops = 40
    host = L.spxb_host_alloc(len(blob))           # pinned, device-visible under UVA: the kernel reads it over PCIe
    C.memmove(host, blob, len(blob))
    hout = L.spxb_host_alloc(hops * cap * ch * 2)
    b = StreamBatch(1, ch, i, o, q)
    b.set_kernel(kernel)
    ref = O.OracleResampler(ch, i, o, q)
    payload = np.frombuffer(w.data, np.int16)
    wants, gots = [], []
    for k in range(hops):
        nin, nout = np.array([n], np.uint32), np.array([cap], np.uint32)
        e = L.spxb_batch_process_device(b._h, C.c_void_p(host + off + k * n * ch * 2), n, nin.ctypes.data,
                                        C.c_void_p(hout + k * cap * ch * 2), cap, nout.ctypes.data)
        assert e == 0, _lib.last_error()
        b.synchronize()
        y, u, m = ref.process(payload[k * n * ch:(k + 1) * n * ch], cap)
        assert (u, m) == (int(nin[0]), int(nout[0]))
        got = np.frombuffer((C.c_char * (m * ch * 2)).from_address(hout + k * cap * ch * 2), np.int16)
        if kernel == KERNEL_STRICT:
            check_close(y, got, exact=True, what=("wav payload", k))
        else:  # +-1 LSB per hop; the 90 dB bar over the whole excerpt (single quiet hops sit a little below it)
            assert np.abs(y.astype(np.int32) - got.astype(np.int32)).max() <= LSB_TOL, k
        wants.append(y)
        gots.append(got.copy())
    assert O.snr_db(np.concatenate(wants), np.concatenate(gots)) >= SNR_MIN_DB
    assert b.last_kernel() == kernel
    b.close()
    L.spxb_host_free(host)
    L.spxb_host_free(hout)
