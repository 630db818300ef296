"""CPU test of the multi-GPU plan (no data-path collective: plain batch splitting).  Two gloo ranks take
their share of the work through the very function bench.py's timed legs call (`rank_batch`: the full batch per
rank under weak scaling, a `shard_batch` span of one batch under --scaling strong), and the per-rank
counts/timings combine into the whole-job throughput exactly as bench.py reports it."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import combine_ranks, rank_batch, shard_batch  # noqa: E402


def test_shard_batch_covers_everything_once():
    for batch in (0, 1, 7, 1 << 20, 262144, 16385):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_batch(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_rank_batch_weak_and_strong():
    for world in (1, 2, 4, 8):
        assert [rank_batch(4096, r, world, "weak") for r in range(world)] == [4096] * world
        assert sum(rank_batch(4097, r, world, "strong") for r in range(world)) == 4097
        assert rank_batch(4097, 0, world, "strong") == shard_batch(4097, 0, world)[1]


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms = 10.0 if rank == 0 else 20.0  # rank 1 is slower: the job is as slow as the slowest rank
    total, worst_ms = combine_ranks(rank_batch(1001, rank, world, "strong"), ms, torch.device("cpu"))
    assert total == 1001 and worst_ms == 20.0
    total, worst_ms = combine_ranks(rank_batch(1001, rank, world, "weak"), ms, torch.device("cpu"))
    assert total == 2002 and worst_ms == 20.0
    dist.destroy_process_group()


def test_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
