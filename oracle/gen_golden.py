"""Generate tests/golden/ from the REAL reference build (oracle/_ref/libspeex_ref.so).

Run in the build container (where /root/reference exists):  python oracle/gen_golden.py
Writes
  tests/golden/vectors.npz        seeded inputs + reference outputs for every MATRIX row,
                                  fed through the processChunk capacity rule in ragged chunks
  tests/golden/oracle_hashes.json FNV-1a-64 of the reference output of the reference's own
                                  resources/*.pcm fixtures (the 7 src/test.ts cases + the
                                  24000->44100 q1/q7/q10 sweep), one-shot
TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from cases import GOLDEN_CHUNKS, GOLDEN_STREAMS, MATRIX, case_id  # noqa: E402
from node_speex_resampler_b200.signals import synth_pcm  # noqa: E402

RES = "/root/reference/resources"
FILE_CASES = [
    ("24000hz_mono_test.pcm", 1, 24000, 48000, 5), ("24000hz_test.pcm", 2, 24000, 24000, 5),
    ("24000hz_test.pcm", 2, 24000, 48000, 10), ("44100hz_test.pcm", 2, 44100, 48000, 7),
    ("44100hz_test.pcm", 2, 44100, 48000, 10), ("44100hz_test.pcm", 2, 44100, 48000, 1),
    ("44100hz_test.pcm", 2, 44100, 24000, 5), ("24000hz_mono_test.pcm", 1, 24000, 44100, 1),
    ("24000hz_mono_test.pcm", 1, 24000, 44100, 7), ("24000hz_mono_test.pcm", 1, 24000, 44100, 10),
]


def main():
    O.build()
    assert O.have_ref(), "needs oracle/_ref/libspeex_ref.so (build container only)"
    arrays = {}
    total = sum(GOLDEN_CHUNKS)
    for idx, c in enumerate(MATRIX):
        ch, i, o, q, _ = c
        pcm = synth_pcm(GOLDEN_STREAMS, ch, total, i, seed=0xB200 + idx)
        outs, lens = [], []
        for s in range(GOLDEN_STREAMS):
            r = O.RefResampler(ch, i, o, q)
            pos, so, sl = 0, [], []
            for n in GOLDEN_CHUNKS:
                chunk = pcm[s, pos * ch:(pos + n) * ch]
                y = np.frombuffer(r.processChunk(chunk), dtype=np.int16)
                so.append(y)
                sl.append(y.size // ch)
                pos += n
            outs.append(np.concatenate(so))
            lens.append(sl)
        key = case_id(c)
        arrays[key + "/in"] = pcm
        arrays[key + "/out0"] = outs[0]
        arrays[key + "/out1"] = outs[1]
        arrays[key + "/lens"] = np.asarray(lens, dtype=np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vectors.npz"), **arrays)

    hashes = {}
    for f, ch, i, o, q in FILE_CASES:
        data = open(os.path.join(RES, f), "rb").read()
        y = O.RefResampler(ch, i, o, q).processChunk(data)
        hashes[f"{f}|{ch}|{i}|{o}|{q}"] = {"frames": len(y) // 2 // ch, "fnv1a64": O.fnv1a64(y)}
    with open(os.path.join(ROOT, "tests", "golden", "oracle_hashes.json"), "w") as fh:
        json.dump({"how": "oracle/gen_golden.py; native gcc build of /root/reference/deps/speex/resample.c "
                          "(-O2 -ffp-contract=off -DFLOATING_POINT -DOUTSIDE_SPEEX), one-shot processChunk",
                   "cases": hashes}, fh, indent=1)
    print("wrote", len(MATRIX), "vector sets and", len(hashes), "file hashes")


if __name__ == "__main__":
    main()
