"""Warm per-kernel device times of one rdesign forward (torch.profiler / CUPTI), tuning aid.  --c3 for 8 x 4096, K = 48."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rna-mpnn_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
from rmpnn_b200.synthetic import RDesignConfig, make_state_dict, synthetic_backbones
from rmpnn_b200 import _lib
from rmpnn_b200.engine import RDesignEngine

B, L, K = (8, 4096, 48) if "--c3" in sys.argv else (64, 256, 30)
prec = _lib.PREC_FP32 if "--fp32" in sys.argv else _lib.PREC_BF16
cfg = RDesignConfig(k_neighbors=K)
eng = RDesignEngine(make_state_dict(cfg, 42), K, cfg.scale, device="cuda", precision=prec)
X, S, mask = synthetic_backbones(B, L, 1002)
Xd, md = X.cuda(), mask.cuda()
sizes = (B * L, B * L * min(K, L))
for _ in range(3):
    eng.forward(Xd, md, sizes=sizes)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        eng.forward(Xd, md, sizes=sizes)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time per forward {tot:.1f} us")
for k, t, n in rows[:14]:
    print(f"{t:9.1f} us {100 * t / tot:5.1f}% {n:6.1f}x {t / n:8.1f} us  {k[:80]}")
