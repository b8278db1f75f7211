"""N > 1 path on CPU: two gloo ranks each own a block of streams (no data-path collective),
process it, and the gathered result equals the single-process result. The CPU oracle stands in
for the device here (the partition / gather plumbing is what is under test); on a GPU box the
same partition feeds one StreamBatch per rank (bench.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from node_speex_resampler_b200.sharding import owner_of, shard_range, shard_sizes, split_chunks  # noqa: E402
from node_speex_resampler_b200.signals import synth_pcm  # noqa: E402

S, CH, IN, OUT, Q, N = 11, 2, 44100, 48000, 7, 441


def test_partition_is_exact_and_balanced():
    for n in (0, 1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 8):
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
            hi_prev = 0
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                assert lo == hi_prev
                hi_prev = hi
            if n:
                assert owner_of(n - 1, n, world) == world - 1 or sizes[-1] == 0
    assert shard_sizes(65536, 8) == [8192] * 8  # BASELINE configs[4]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    pcm = synth_pcm(S, CH, N * 3, IN, seed=123)
    lo, hi = shard_range(S, world, rank)
    mine = split_chunks([pcm[s] for s in range(S)], world, rank)
    assert len(mine) == hi - lo
    outs = []
    refs = [O.OracleResampler(CH, IN, OUT, Q) for _ in mine]
    for k in range(3):
        dist.barrier()  # the only collective: step alignment, never sample data
        outs.append([np.frombuffer(r.processChunk(x[k * N * CH:(k + 1) * N * CH]), dtype=np.int16)
                     for r, x in zip(refs, mine)])
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks bench.py takes of its timing
    assert t.item() == world
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi,
             **{f"s{lo + i}_k{k}": outs[k][i] for k in range(3) for i in range(hi - lo)})
    dist.destroy_process_group()


def test_two_ranks_equal_one_process(tmp_path):
    from oracle import oracle as O
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    pcm = synth_pcm(S, CH, N * 3, IN, seed=123)
    got = {}
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        got.update({k: z[k] for k in z.files if k.startswith("s")})
    for s in range(S):
        ref = O.OracleResampler(CH, IN, OUT, Q)
        for k in range(3):
            want = np.frombuffer(ref.processChunk(pcm[s][k * N * CH:(k + 1) * N * CH]), dtype=np.int16)
            assert np.array_equal(got[f"s{s}_k{k}"], want), (s, k)
