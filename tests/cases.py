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
