"""Multi-rank host logic on CPU (gloo, world_size 2): structure sharding and the max-over-ranks timing
reduction used by bench.py.  No GPU, no data-path collective."""
import os
import random

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rmpnn_b200 import shard


def _worker(rank, world, port, lengths, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = shard.partition(lengths, world, top_k=30, token_budget=4096)
    mine = sorted(i for b in parts[rank] for i in b)
    # every rank derives the same partition; exchange the owned ids to prove it covers everything once
    counts = torch.zeros(len(lengths), dtype=torch.int64)
    counts[mine] = 1
    dist.all_reduce(counts)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)      # per-rank step time -> max over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nt = torch.tensor([float(sum(lengths[i] for i in mine))], dtype=torch.float64)
    dist.all_reduce(nt)
    if rank == 0:
        out.put((counts.tolist(), t.item(), nt.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_structure_sharding_two_ranks():
    random.seed(5)
    lengths = [random.randint(50, 1000) for _ in range(200)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randint(0, 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, tmax, nt = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert counts == [1] * len(lengths)                # each structure owned by exactly one rank
    assert tmax == 11.0 and nt == float(sum(lengths))


def test_partition_is_balanced_and_batches_respect_budget():
    random.seed(1)
    lengths = [random.randint(50, 1000) for _ in range(1000)]
    parts = shard.partition(lengths, 8, top_k=30, token_budget=8192)
    assert sorted(i for p in parts for b in p for i in b) == list(range(1000))
    for p in parts:
        for b in p:
            assert max(lengths[i] for i in b) * len(b) <= 8192 or len(b) == 1
    worst, mean = shard.imbalance(lengths, parts)
    assert worst / mean < 1.10          # ~9 batches per rank: longest-processing-time-first granularity


def test_transposed_csr_host_logic():
    """Edges grouped by neighbour node (the training backward's atomics-free d Pn): stable permutation + offsets, against numpy;
    nodes that are nobody's neighbour get an empty range.  Runs on the CPU (torch ops only)."""
    import numpy as np
    from rmpnn_b200 import ops
    rng = np.random.default_rng(3)
    n_seg, rows = 57, 1000
    idx = rng.integers(0, n_seg - 5, size=rows).astype(np.int32)
    perm, row_ptr = ops.transposed_csr(torch.from_numpy(idx), n_seg)
    assert perm.dtype == row_ptr.dtype == torch.int32 and row_ptr.shape[0] == n_seg + 1
    assert np.array_equal(perm.numpy(), np.argsort(idx, kind="stable").astype(np.int32))
    counts = np.bincount(idx, minlength=n_seg)
    assert np.array_equal(row_ptr.numpy(), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
    for v in (0, 17, n_seg - 1):
        seg = perm.numpy()[row_ptr[v]:row_ptr[v + 1]]
        assert (idx[seg] == v).all() and len(seg) == counts[v]
    assert abs(ops.keep_scale(0.1) - 256.0 / 230.0) < 1e-12 and ops.keep_scale(0.0) == 1.0
