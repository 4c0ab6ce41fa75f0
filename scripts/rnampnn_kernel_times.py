"""Warm per-kernel device times of one rnampnn forward (torch.profiler / CUPTI), tuning aid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rna-mpnn_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
from rmpnn_b200.synthetic import RNAMPNNConfig, make_rnampnn_state_dict, synthetic_backbones
from rmpnn_b200 import _lib
from rmpnn_b200.engine_rnampnn import RNAMPNNEngine

prec = _lib.PREC_FP32 if "--fp32" in sys.argv else _lib.PREC_BF16
cfg = RNAMPNNConfig(); params = make_rnampnn_state_dict(cfg, 42)
B, L = (8, 4096) if "--r3" in sys.argv else (64, 256)
coords, _, mask = synthetic_backbones(B, L, 1012, atoms=7)
eng = RNAMPNNEngine.from_config(params, cfg, device="cuda", precision=prec)
cd, md = coords.cuda(), mask.cuda()
for _ in range(3):
    eng.forward(cd, md)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        eng.forward(cd, md)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time per forward {tot:.1f} us")
for k, t, n in rows[:18]:
    print(f"{t:9.1f} us {100 * t / tot:5.1f}% {n:6.1f}x {t / n:8.1f} us  {k[:90]}")
