"""Run by tests/test_parity_gpu.py (test_persistent_kernel_parity, test_long_filter_on_either_tensor_kernel)
in a subprocess whose environment picks the tensor kernel and its geometry (SPXB_UMMA_RESIDENT, _NT,
_DENSE, SPXB_UMMA2_XSTAGES are read once per process): the tensor kernel against the oracle -- a batch
large enough that persistent CTAs walk several tiles and change tap tile on the way."""
import sys

import numpy as np

import node_speex_resampler_b200 as pkg
from oracle import oracle as O

name, S, ch, i, o, q, n, calls = sys.argv[1], *map(int, sys.argv[2:9])
cap = -(-n * o // i) + 1
b = pkg.StreamBatch(S, ch, i, o, q)
b.set_kernel(pkg.KERNEL_TENSOR)
check = sorted(set([0, 1, 63, S // 2, S - 2, S - 1]) & set(range(S)))
refs = {s: O.OracleResampler(ch, i, o, q) for s in check}
worst = 0
for k in range(calls):
    pcm = pkg.synth_pcm(S, ch, n, i, seed=0xE51D, start_frame=k * n)
    out, used, made = b.process(pcm, n, cap)
    for s in check:
        y, u, m = refs[s].process(pcm[s], cap)
        assert (u, m) == (int(used[s]), int(made[s])), (k, s)
        d = np.abs(y.astype(np.int32) - out[s, : m * ch].astype(np.int32))
        worst = max(worst, int(d.max(initial=0)))
        assert d.max(initial=0) <= 1 and O.snr_db(y, out[s, : m * ch]) >= 90.0, (k, s, int(d.max()))
geom = b.tensor_geometry()
ls, fr, mg, hist = b.get_state(check[-1])
rls, rfr, rhist = refs[check[-1]].state(0)
assert (ls, fr) == (rls, rfr) and np.array_equal(hist.reshape(-1, ch)[:, 0].astype(np.float32), rhist)
print(f"resident ok {name} geom={geom} worst_lsb={worst}")
