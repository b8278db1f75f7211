"""Parity matrix shared by the CPU and GPU tests and by oracle/gen_golden.py.

Rows 0-6 are the seven cases of the reference's own test (src/test.ts:14-22), rows 7-16 the
24000->44100 quality sweep of BASELINE.json configs[1], then the C4 / C5 shapes and edge
ratios (q0, pure up/down-sampling, 3 channels, a huge denominator, a x12 up-sampler that
makes the 1024-output block limit of resample.c:982-991 bind)."""

# (channels, in_rate, out_rate, quality, reference kernel)
MATRIX = [
    (1, 24000, 48000, 5, "direct_single"),
    (2, 24000, 24000, 5, "direct_single"),
    (2, 24000, 48000, 10, "direct_double"),
    (2, 44100, 48000, 7, "interpolate_single"),
    (2, 44100, 48000, 10, "interpolate_double"),
    (2, 44100, 48000, 1, "interpolate_single"),
    (2, 44100, 24000, 5, "interpolate_single"),
] + [(1, 24000, 44100, q, "interpolate_double" if q > 8 else "interpolate_single")
     for q in range(1, 11)] + [
    (1, 48000, 16000, 10, "direct_double"),
    (2, 96000, 44100, 10, "interpolate_double"),
    (2, 44100, 48000, 0, "interpolate_single"),
    (1, 8000, 48000, 3, "direct_single"),
    (3, 48000, 44100, 4, "interpolate_single"),
    (1, 16000, 8000, 8, "direct_single"),
    (2, 44100, 48001, 3, "interpolate_single"),
    (1, 8000, 96000, 2, "interpolate_single"),
]

# chunk sizes (frames) of the golden vectors: 20 ms-ish, ragged, tiny, and one long chunk
GOLDEN_CHUNKS = [480, 441, 7, 300, 1, 882]
GOLDEN_STREAMS = 2


def case_id(c):
    ch, i, o, q, _ = c
    return f"{ch}ch_{i}to{o}_q{q}"


# Float-entry golden vectors (oracle/gen_golden_f32.py -> tests/golden/vectors_f32.npz): MATRIX
# rows covering the four reference kernels, each driven through this sequence of calls on ONE
# state -- (kind, input frames, output capacity in frames), kind "f" = float entry, "i" = int16
# entry. Capacities 40 and 1500 bind on some rows, which is where the float entry's block walk
# (no 1024-frame output block, resample.c:944) differs from the int16 entry's.
F32_ROWS = [0, 2, 3, 4, 17, 20, 24]
F32_CALLS = [("f", 480, 4000), ("i", 441, 4000), ("f", 7, 4000), ("f", 300, 40), ("i", 1, 4000),
             ("f", 882, 1500), ("f", 0, 10), ("i", 200, 0), ("f", 333, 4000)]


def f32_input(row, call, kind, n, ch, in_rate):
    """seeded input of one call: int16 PCM for "i", non-integer floats of PCM scale for "f" """
    import numpy as np
    from node_speex_resampler_b200.signals import synth_pcm
    pcm = synth_pcm(1, ch, max(n, 1), in_rate, seed=0xF32 + 100 * row + call)[0][: n * ch]
    if kind == "i":
        return pcm
    return (pcm.astype(np.float32) * np.float32(0.731) + np.float32(0.123)).astype(np.float32)
