#!/usr/bin/env python
"""Stock-PyTorch-on-B200 column (BASELINE.md section 3 / SURVEY.md section 8d): the UNMODIFIED reference RNAModel
(staged copy oracle/_ref, see oracle/make_ref.py) moved to the GPU with `.cuda()` and timed on the bench workload
under torch.set_float32_matmul_precision 'highest' and 'medium' (the reference's own seeding() asks for 'medium').

    python scripts/bench_reference_gpu.py [--workload c2|c3] [--steps 10]

This is a comparison baseline (stock ATen / cuBLAS kernels), not part of the product; it prints one JSON line."""
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"])
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from bench import WORKLOADS
    from oracle.ref_loader import load_reference_classes
    from rmpnn_b200.synthetic import RDesignConfig, make_state_dict, synthetic_backbones
    ref = load_reference_classes()
    if ref is None:
        print(json.dumps({"unavailable": "no staged reference (run python oracle/make_ref.py where /root/reference exists)"}))
        return
    B, L, K, seed = WORKLOADS[args.workload]
    cfg = RDesignConfig(k_neighbors=K)
    model = ref["RNAModel"](**cfg.as_kwargs()).eval()
    model.load_state_dict(make_state_dict(cfg, 42), strict=True)
    model = model.cuda()
    X, S, mask = synthetic_backbones(B, L, seed)
    if args.workload == "c3":
        X, S, mask = X[:1], S[:1], mask[:1]        # one 4096-nt structure: the reference's [B,N,N,3] tensor is 1.6 GB per 8
    Xd, Sd, md = X.cuda(), S.cuda(), mask.cuda()
    nt = float(mask.sum())
    out = {"workload": args.workload, "structures": int(X.shape[0]), "L": L, "K": K, "gpu": torch.cuda.get_device_name(0),
           "torch": torch.__version__, "what": "unmodified reference RNAModel.forward + readout + argmax, stock PyTorch"}
    for prec in ("highest", "medium"):
        torch.set_float32_matmul_precision(prec)
        with torch.no_grad():
            for _ in range(3):
                h, _ = model(Xd, Sd, md)
                model.readout(h).argmax(-1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                h, _ = model(Xd, Sd, md)
                seq = model.readout(h).argmax(-1)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[prec] = {"ms_per_step": ms, "nt_per_s": nt / (ms * 1e-3), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
