#!/usr/bin/env python
"""Batched design sweep (BASELINE.json configs[4], SURVEY.md section 8d "C5"): many synthetic structures of mixed
length, sharded by whole structures across the ranks, temperature sampling of a sequence for each.

    python scripts/design_sweep.py [--structures 100000] [--temperature 0.5] [--token-budget 32768] [--precision bf16]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/design_sweep.py ...

Per rank: lengths L_i ~ U[50, 1000] (seed 1005, identical on every rank) -> shard.partition (length-sorted
token-budget batches, longest-processing-time-first) -> for each owned batch: device-side collate from the packed
coordinates (rmp_collate) -> kNN + featurizer + 9 layers + readout with Philox temperature sampling -> int8
sequences gathered into one device buffer.  No data-path collective; NCCL only for the barrier / max-reduce.
Shapes change per batch, so the forward is launched eagerly (no CUDA graph).

Prints one JSON line: nt/s and structures/s with inputs resident in HBM, and the same with the packed coordinates
streamed from pinned host memory batch by batch (e2e).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rna-mpnn_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--structures", type=int, default=100000)
    ap.add_argument("--min-len", type=int, default=50)
    ap.add_argument("--max-len", type=int, default=1000)
    ap.add_argument("--temperature", type=float, default=0.5)
    ap.add_argument("--token-budget", type=int, default=32768)
    ap.add_argument("--k", type=int, default=30)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "fp32x3", "bf16"])
    ap.add_argument("--check", action="store_true", help="T=0 argmax of ~200 structures against the CPU oracle")
    ap.add_argument("--check-structures", type=int, default=200)
    ap.add_argument("--total-nt", type=int, default=0,
                    help="draw structures until this many nucleotides are reached (overrides --structures)")
    return ap.parse_args(argv)


def run(args, init_dist=True):
    """Runs the sweep on this rank; returns the result dict on rank 0 (None elsewhere)."""

    import torch.distributed as dist
    from rmpnn_b200.synthetic import RDesignConfig, make_state_dict
    from rmpnn_b200 import _lib, ops, shard
    from rmpnn_b200.engine import RDesignEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("design_sweep.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and init_dist and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    g = torch.Generator().manual_seed(1005)
    lengths = torch.randint(args.min_len, args.max_len + 1, (args.structures,), generator=g).tolist()
    if args.total_nt:
        acc, n = 0, 0
        while n < len(lengths) and acc < args.total_nt:
            acc += lengths[n]
            n += 1
        lengths = lengths[:n]
    parts = shard.partition(lengths, world, top_k=args.k, token_budget=args.token_budget)
    mine = parts[rank]
    own = [i for b in mine for i in b]
    # this rank's structures, packed back to back in batch order: random-walk backbones generated on the device
    own_len = torch.tensor([lengths[i] for i in own], dtype=torch.int64)
    starts = torch.cumsum(own_len, 0) - own_len
    total = int(own_len.sum())
    gd = torch.Generator(device=dev).manual_seed(1005 * 1000 + rank)
    step = 3.5 * torch.randn(total, 1, 3, generator=gd, device=dev)
    P = torch.cumsum(step, 0)                      # one long walk; structure boundaries are just cuts of it
    packed = (P + 1.5 * torch.randn(total, 6, 3, generator=gd, device=dev)).contiguous()
    del step, P
    packed_host = packed.cpu().pin_memory()

    cfg = RDesignConfig(k_neighbors=args.k)
    prec = {"bf16": _lib.PREC_BF16, "fp32x3": _lib.PREC_FP32X3}.get(args.precision, _lib.PREC_FP32)
    eng = RDesignEngine(make_state_dict(cfg, 42), args.k, cfg.scale, device=dev, precision=prec)

    # per batch: host-side sizes (no device read-back), device-side row_start / lengths
    plan, at = [], 0
    for b in mine:
        L = [lengths[i] for i in b]
        n = len(b)
        N = max(L)
        Nv = sum(L)
        Ne = sum(l * min(args.k, N, l) for l in L)
        plan.append((at, n, N, Nv, Ne))
        at += n
    starts_d = starts.to(dev)
    len_d = own_len.to(torch.int32).to(dev)
    seqs = torch.empty(total, dtype=torch.int8, device=dev)
    stream = torch.cuda.current_stream()

    def sweep(temperature, from_host, limit=None):
        out_at = 0
        for bi, (s0, n, N, Nv, Ne) in enumerate(plan[:limit]):
            if from_host:          # stream this batch's residues from pinned host memory
                r0 = int(starts[s0])
                src = packed_host[r0:r0 + Nv].to(dev, non_blocking=True)
                rs = starts_d[s0:s0 + n] - r0
            else:
                src, rs = packed, starts_d[s0:s0 + n]
            X, mask = ops.collate(src, rs, len_d[s0:s0 + n], N, 6)
            out = eng.forward(X, mask, sizes=(Nv, Ne), temperature=temperature, seed=1005 + bi)
            seqs[out_at:out_at + Nv] = out["samples"].to(torch.int8)
            out_at += Nv
        return out_at

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sweep(args.temperature, False, limit=min(8, len(plan)))          # warm-up
    barrier()
    L_ = _lib.lib()
    n0 = L_.rmp_launch_count()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(stream); done = sweep(args.temperature, False); e[1].record(stream)
    barrier()
    launches = L_.rmp_launch_count() - n0
    e[2].record(stream); sweep(args.temperature, True); host_seqs = seqs.cpu(); e[3].record(stream)
    barrier()
    t = torch.tensor([e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])], dtype=torch.float64, device=dev)
    nt = torch.tensor([float(done), float(len(own))], dtype=torch.float64, device=dev)
    # per-rank busy time and load (nucleotides, launches): shows whether the limiter is the partition or the launches
    mine_stats = torch.tensor([e[0].elapsed_time(e[1]), float(done), float(launches), float(len(plan))],
                              dtype=torch.float64, device=dev)
    all_stats = [torch.zeros_like(mine_stats) for _ in range(world)]
    if world > 1:
        dist.all_gather(all_stats, mine_stats)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nt)
    else:
        all_stats = [mine_stats]
    ok = None
    if args.check and rank == 0:
        # T = 0 argmax against the CPU oracle on ~200 structures spread over this rank's batches (short and long).
        # The oracle sees the same padded rows (a structure's result depends on its batch's N).
        from oracle import rdesign_oracle as O
        params = make_state_dict(cfg, 42)
        want, per_batch = args.check_structures, 8
        picks = sorted({int(round(j * (len(plan) - 1) / max(1, want // per_batch - 1))) for j in range(max(1, want // per_batch))})
        same = seen = n_struct = 0
        worst = 0.0
        for bi in picks:
            s0, n, N, Nv, Ne = plan[bi]
            X, mask = ops.collate(packed, starts_d[s0:s0 + n], len_d[s0:s0 + n], N, 6)
            out = eng.forward(X, mask, sizes=(Nv, Ne), temperature=0.0)
            m = min(n, per_batch)
            ref = O.forward(params, X[:m].cpu(), torch.zeros(m, N, dtype=torch.long), mask[:m].cpu(), args.k, cfg.scale)
            k = ref["samples"].shape[0]
            same += int((out["samples"][:k].cpu() == ref["samples"]).sum())
            worst = max(worst, float((out["logits"][:k].cpu() - ref["logits"]).abs().max()))
            seen += k
            n_struct += m
        ok = {"structures": n_struct, "nucleotides": seen, "agreement": same / max(1, seen), "max_abs_dlogit": worst}
    result = None
    if rank == 0:
        busy = [float(a[0]) for a in all_stats]
        per_rank = {"busy_ms": busy, "nucleotides": [int(a[1]) for a in all_stats],
                    "launches": [int(a[2]) for a in all_stats], "batches": [int(a[3]) for a in all_stats],
                    "worst_over_mean_busy": max(busy) / (sum(busy) / len(busy)),
                    "worst_over_mean_nt": max(float(a[1]) for a in all_stats) / (sum(float(a[1]) for a in all_stats) / world)}
        resident_ms, e2e_ms = t.tolist()
        total_nt, total_structs = nt.tolist()
        pad = sum(n * N for _, n, N, _, _ in plan) / max(1, sum(Nv for *_, Nv, _ in plan))
        result = {
            "metric": "RNA nucleotides/sec (design sweep, temperature sampling)", "value": total_nt / (resident_ms * 1e-3),
            "unit": "nt/s", "structures_per_s": total_structs / (resident_ms * 1e-3), "n_gpus": world,
            "sweep_ms": resident_ms, "higher_is_better": True, "scaling": "strong",
            "dtype": "bf16" if prec == _lib.PREC_BF16 else "f32", "data": "synthetic",
            "config": {"workload": f"c5: {len(lengths)} synthetic structures L~U[{args.min_len},{args.max_len}], K={args.k}, "
                                   f"T={args.temperature}, RNAModel default", "token_budget": args.token_budget,
                       "batches_rank0": len(plan), "padding_overhead_rank0": pad,
                       "sharding": "whole structures, length-sorted token-budget batches, LPT over ranks; no collective",
                       "launch": "eager (shapes change per batch)"},
            "e2e": {"value": total_nt / (e2e_ms * 1e-3), "unit": "nt/s", "h2d_bytes": int(total_nt) * 72,
                    "d2h_bytes": int(total_nt)},
            "gpu_launches": int(launches), "per_rank": per_rank, "argmax_agreement_with_oracle_T0": ok}
    return result


def main():
    import torch.distributed as dist
    res = run(parse_args())
    if res is not None:
        print(json.dumps(res), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
