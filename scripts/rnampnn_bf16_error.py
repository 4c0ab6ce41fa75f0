"""Logit error of the rnampnn bf16 arm on the golden cases, with and without the raw-geometry stacks in fp32,
and the step time of both settings at the r2 workload (tuning aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rna-mpnn_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from conftest import load_rnampnn_golden, RNAMPNN_GOLDEN_CASES
from oracle import rnampnn_oracle as O
from rmpnn_b200.synthetic import RNAMPNNConfig, make_rnampnn_state_dict, synthetic_backbones
from rmpnn_b200 import _lib
from rmpnn_b200.engine_rnampnn import RNAMPNNEngine

for name in RNAMPNN_GOLDEN_CASES:
    g = load_rnampnn_golden(name)
    ref = O.forward(g["params"], g["coords"], g["mask"], g["config"])
    for raw32 in ("bf16", "fp16", "fp32"):
        eng = RNAMPNNEngine.from_config(g["params"], g["config"], device="cuda", precision=_lib.PREC_BF16)
        eng.raw_stacks = raw32
        out = eng.forward(g["coords"].cuda(), g["mask"].cuda())
        err = (out["logits"].cpu() - ref["logits"].reshape(-1, 4)).abs().max().item()
        valid = g["mask"].reshape(-1).bool()
        agree = (out["samples"].cpu()[valid] == ref["samples"].reshape(-1)[valid]).float().mean().item()
        print(f"{name:18s} raw_stacks={raw32:5s} max|dlogit|={err:.2e} argmax agreement={agree:.4f}")

cfg = RNAMPNNConfig(); params = make_rnampnn_state_dict(cfg, 42)
coords, _, mask = synthetic_backbones(64, 256, 1012, atoms=7)
ref = O.forward(params, coords[:4], mask[:4], cfg)
for prec, raw32 in ((_lib.PREC_FP32, "fp32"), (_lib.PREC_BF16, "bf16"), (_lib.PREC_BF16, "fp16"), (_lib.PREC_BF16, "fp32")):
    eng = RNAMPNNEngine.from_config(params, cfg, device="cuda", precision=prec)
    eng.raw_stacks = raw32
    out = eng.forward(coords[:4].cuda(), mask[:4].cuda())
    err = (out["logits"].cpu() - ref["logits"].reshape(-1, 4)).abs().max().item()
    replay, _ = eng.capture(coords.cuda(), mask.cuda())
    for _ in range(3): replay()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(20): replay()
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 20
    print(f"r2 precision={'bf16' if prec else 'fp32'} raw_stacks={raw32}: {ms:.3f} ms/step, {64*256/ms*1e3/1e6:.2f} M nt/s, max|dlogit| (4 structures) {err:.2e}")
