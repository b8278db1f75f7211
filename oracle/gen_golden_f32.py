"""Generate tests/golden/vectors_f32.npz from the REAL reference build (oracle/_ref): the float
entry (speex_resampler_process_interleaved_float, float build) mixed with int16 calls on one
state, including calls whose output capacity binds (where the float entry's block walk differs
from the int16 entry's). Run in the build container:  python oracle/gen_golden_f32.py
TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from cases import F32_CALLS, F32_ROWS, MATRIX, case_id, f32_input  # noqa: E402


def main():
    O.build()
    assert O.have_ref(), "needs oracle/_ref/libspeex_ref.so (build container only)"
    arrays = {}
    for row in F32_ROWS:
        ch, i, o, q, _ = MATRIX[row]
        r = O.RefResampler(ch, i, o, q)
        key = case_id(MATRIX[row])
        for k, (kind, n, cap) in enumerate(F32_CALLS):
            x = f32_input(row, k, kind, n, ch, i)
            y, used, made = (r.process_float if kind == "f" else r.process)(x, cap)
            arrays[f"{key}/call{k}/out"] = y
            arrays[f"{key}/call{k}/lens"] = np.asarray([used, made], dtype=np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vectors_f32.npz"), **arrays)
    print("wrote", len(F32_ROWS), "float/int16 call sequences")


if __name__ == "__main__":
    main()
