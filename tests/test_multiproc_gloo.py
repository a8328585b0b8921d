"""N>1 path on CPU: world_size-2 gloo.  The data path has no collective (clips shard in
contiguous blocks); torch.distributed only carries the barrier and the max-over-ranks timing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from audio_edge_ml_pipeline_b200 import dist as D


def test_shard_bounds_partition_exactly():
    for n in (0, 1, 7, 100000, 2025):
        for w in (1, 2, 3, 4, 8):
            b = [D.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    w, r = D.init("gloo")
    assert (w, r) == (world, rank)
    from audio_edge_ml_pipeline_b200 import synth
    from oracle import librosa_restated as L
    pcm = synth.make_suite(6, 16000, 4000, seed=9)          # every rank holds the same list, takes its block
    lo, hi = D.shard_bounds(len(pcm), w, r)
    mine = np.stack([L.audio_mel_spec(L.pcm16_to_float(c)) for c in pcm[lo:hi]])
    D.barrier()
    secs = 0.5 if rank == 0 else 2.0                         # pretend rank 1 is the slow one
    thr = D.aggregate_throughput(hi - lo, secs)
    mx = D.max_over_ranks(secs)
    # device-to-host "gather" stand-in: ranks write disjoint slices, rank 0 checks the concatenation
    parts = [None] * w
    torch.distributed.all_gather_object(parts, (lo, hi, mine))
    if rank == 0:
        full = np.concatenate([p[2] for p in sorted(parts, key=lambda x: x[0])])
        ref = np.stack([L.audio_mel_spec(L.pcm16_to_float(c)) for c in pcm])
        q.put((thr, mx, bool(np.array_equal(full, ref))))
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_shards_and_times_like_bench():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    thr, mx, same = q.get()
    assert mx == 2.0 and thr == pytest.approx(6 / 2.0) and same
