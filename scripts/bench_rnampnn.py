#!/usr/bin/env python
"""Throughput of the secondary (rnampnn) path on one B200 -- same JSON shape as bench.py, which stays the
headline (rdesign) bench the driver runs.

    python scripts/bench_rnampnn.py [--steps K] [--warmup W] [--workload r2|r3] [--no-graph] [--no-cpu-baseline]

One step = RNAMPNN.forward (centroid kNN + raw features + embeddings + 10 ResMPNN layers + 2 attention layers +
FFNs + readout/argmax) over one synthetic padded batch of 7-atom random-walk backbones.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rna-mpnn_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from bench import ClockSampler, peaks  # noqa: E402

WORKLOADS = {"r2": (64, 256, 1012), "r3": (8, 4096, 1013), "r4": (256, 256, 1014)}     # (B, L, seed); K = 3 (reference default)
UNIT = "nt/s"


def flops_per_nt(cfg):
    """Dense-contraction flops per residue of the reference formulation (MAC = 2), attention counted on the
    L_b x L_b valid block only (the reference pads to 4500 x 4500)."""
    H, K = 128, cfg.num_res_neighbours
    f = 2 * (cfg.raw_dim * H + H * 512 + 2 * 512 * 512 + 512 * H)                 # raw_project + embedding FFN
    f += K * 2 * (cfg.raw_edge_dim * H + H * H)                                    # edge embedding
    f += cfg.num_res_mpnn_layers * K * 2 * 2 * (3 * H * H + H * H)                 # message + edge MLPs
    f += cfg.num_post_fusion_attn_layers * 2 * (3 * H * H + H * H)                 # in_proj + out_proj
    f += 2 * (H * 512 + 2 * 512 * 512 + 512 * H)                                   # post-fusion FFN
    f += 2 * (cfg.raw_dim * 512 + 2 * 512 * 512 + 512 * H)                         # RawFFN
    f += 2 * (2 * H * 512 + 512 * 4)                                               # readout
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="r2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", dest="graph", action="store_false", default=True)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    args = ap.parse_args()

    from rmpnn_b200.synthetic import RNAMPNNConfig, make_rnampnn_state_dict, synthetic_backbones
    from rmpnn_b200 import _lib
    from rmpnn_b200.engine_rnampnn import RNAMPNNEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench_rnampnn.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", 0)
    B, L, seed = WORKLOADS[args.workload]
    cfg = RNAMPNNConfig()
    params = make_rnampnn_state_dict(cfg, 42)
    coords, _, mask = synthetic_backbones(B, L, seed, atoms=7)
    prec = _lib.PREC_BF16 if args.precision == "bf16" else _lib.PREC_FP32
    eng = RNAMPNNEngine.from_config(params, cfg, device=dev, precision=prec)
    Nv = int(mask.sum())
    cd, md = coords.to(dev), mask.to(dev)
    cp, mp = coords.pin_memory(), mask.pin_memory()
    samples_host = torch.empty(B * L, dtype=torch.int64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    L_ = _lib.lib()

    replay = static = None
    if args.graph:
        replay, static = eng.capture(cd, md)

    def step_resident():
        return replay() if replay else eng.forward(cd, md)

    def step_e2e():
        if replay:
            static["coords"].copy_(cp, non_blocking=True)
            static["mask"].copy_(mp, non_blocking=True)
            out = replay()
        else:
            out = eng.forward(cp.to(dev, non_blocking=True), mp.to(dev, non_blocking=True))
        samples_host.copy_(out["samples"], non_blocking=True)
        return out

    for _ in range(max(3, args.warmup)):
        step_resident(); step_e2e()
    torch.cuda.synchronize()

    def timed(fn):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        torch.cuda.synchronize()
        for s, e in ev:
            flush.zero_()
            s.record(stream); fn(); e.record(stream)
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in ev)

    with ClockSampler(0) as clocks:
        total_ms = timed(step_resident)
    e2e_ms = timed(step_e2e)
    n0 = L_.rmp_launch_count()
    eng.kernel_timer = []
    eng.forward(cd, md)
    torch.cuda.synchronize()
    launches = L_.rmp_launch_count() - n0
    layer_ms = sum(s.elapsed_time(e) for s, e in eng.kernel_timer)
    eng.kernel_timer = None

    cb = None
    if not args.no_cpu_baseline:
        from oracle import rnampnn_oracle as O
        sB = max(1, min(B, 8 if L <= 512 else 1))
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        O.forward(params, coords[:1], mask[:1], cfg)
        reps, t0 = 0, time.perf_counter()
        while reps < 1 or time.perf_counter() - t0 < 10.0:
            O.forward(params, coords[:sB], mask[:sB], cfg)
            reps += 1
        el = time.perf_counter() - t0
        cb = {"value": float(mask[:sB].sum()) * reps / el, "unit": UNIT, "cores": cores, "kind": "port",
              "sample": f"{sB} of {B} structures (L={L}), {reps} forward passes in {el:.1f} s; the port attends over the "
                        "L valid rows, the reference itself over a 4500-row padding"}

    pk = peaks()
    step_ms = total_ms / args.steps
    fl = flops_per_nt(cfg) * Nv
    line = {"metric": "RNA nucleotides/sec (rnampnn fwd, secondary path)", "value": Nv * args.steps / (total_ms * 1e-3),
            "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: synthetic 7-atom RNA backbones B={B} L={L} K=3, RNAMPNN default "
                                   "(10 ResMPNN layers, 2 attention layers, H=128)",
                       "l2": "256 MiB flush buffer written between timed iterations",
                       "precision": args.precision + (" (raw-geometry stacks with fp16 operands)" if args.precision == "bf16" else ""),
                       "launch": "CUDA graph replay" if replay else "eager"},
            "roofline": {"bound": "tensor" if args.precision == "bf16" else "fp32 FMA pipe (CUDA-core arm)",
                         "achieved": fl / (step_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "algorithmic_flop_per_step": fl,
                         "res_mpnn_layers_ms_eager": layer_ms,
                         "peak": pk["bf16_tflops_sustained"] if args.precision == "bf16" else None,
                         "frac": fl / (step_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"] if args.precision == "bf16" else None,
                         "traffic": None, "note": "whole-step algorithmic flops / step time (all kernels), not one kernel"},
            "cpu_baseline": cb,
            "e2e": {"value": Nv * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": cp.numel() * 4 + mp.numel() * 4, "d2h_bytes_per_step": samples_host.numel() * 8},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "peaks": pk["source"]}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
