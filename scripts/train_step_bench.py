#!/usr/bin/env python
"""Training step of BASELINE.json configs[3] (SURVEY.md "C4"): 32 synthetic structures of L=512 per GPU, K=25,
bf16 autocast, forward + backward + one flat-buffer NCCL all-reduce of the gradients + Adam.

    python scripts/train_step_bench.py [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/train_step_bench.py

INTERIM PATH: in ``train()`` mode the drop-in ``RNAModel`` takes the kNN graph and the raw
geometric features from the sm_100a kernels (they carry no gradient in the reference either) and composes the
differentiable part -- embeddings, message / dense MLPs, LayerNorms, readout, loss -- from torch-CUDA ops, so
autograd provides the backward.  Native backward kernels are SURVEY section 8f rank 1 and not built yet; this script
exists so that the named config runs end to end and the number it prints is the one the native backward has to beat.
The gradient exchange is what the reference's Lightning DDP does, restated as ONE all-reduce of a flat fp32 buffer
(2,551,812 parameters = 10.2 MB) over NCCL / NVLink, averaged, followed by a fused Adam step.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--length", type=int, default=512)
    ap.add_argument("--no-autocast", action="store_true")
    args = ap.parse_args()

    import torch.distributed as dist
    import torch.nn.functional as F
    from rmpnn_b200.synthetic import RDesignConfig, make_state_dict, synthetic_backbones
    from rdesign.model.rdesign import RNAModel
    from rmpnn_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("train_step_bench.py: no CUDA device")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = RDesignConfig()                                   # K = 25, 9 layers (the reference's training default)
    model = RNAModel(**cfg.as_kwargs())
    model.load_state_dict(make_state_dict(cfg, 42), strict=True)
    model = model.to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    n_params = sum(p.numel() for p in params)
    flat = torch.zeros(n_params, dtype=torch.float32, device=dev)
    views, at = [], 0
    for p in params:                                        # gradients live as views of one flat buffer
        v = flat[at:at + p.numel()].view_as(p)
        p.grad = v
        views.append(v)
        at += p.numel()
    opt = torch.optim.Adam(params, lr=model.hparams.lr, fused=True)
    X, S, mask = synthetic_backbones(args.batch, args.length, 1004 + rank)
    X, S, mask = X.to(dev), S.to(dev), mask.to(dev)
    Nv = int(mask.sum())
    L_ = _lib.lib()

    def step():
        flat.zero_()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not args.no_autocast):
            h_V, S_v = model(X, S, mask)
            loss = F.cross_entropy(model.readout(h_V).float(), S_v)
        loss.backward()
        if world > 1:
            dist.all_reduce(flat)                           # the one data-path collective of training
            flat.div_(world)
        opt.step()
        return loss

    for _ in range(max(3, args.warmup)):
        loss = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = L_.rmp_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        loss = step()
    ev[1].record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = t.item() / args.steps
        print(json.dumps({
            "metric": "RNA nucleotides/sec (training step: fwd + bwd + grad all-reduce + Adam)",
            "value": Nv * world / (ms * 1e-3), "unit": "nt/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "dtype": "f32" if args.no_autocast else "bf16 autocast", "data": "synthetic", "loss": float(loss),
            "config": {"workload": f"c4: {args.batch} x L={args.length} per GPU, K={cfg.k_neighbors}, RNAModel default, Adam",
                       "path": "INTERIM: graph + raw features on the sm_100a kernels, differentiable part composed from "
                               "torch-CUDA ops (autograd backward); native backward kernels not built yet",
                       "collective": f"one all-reduce of a flat fp32 gradient buffer, {n_params * 4 / 1e6:.1f} MB"},
            "gpu_launches": int((L_.rmp_launch_count() - n0) // args.steps)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
