#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -u -m pytest tests/test_forest_head.py -x -q > gpurun_out/r2c23_forest.log 2>&1; tail -15 gpurun_out/r2c23_forest.log
python - <<'PY' 2>&1 | tee gpurun_out/r2c23_forest_time.log
import numpy as np, torch, sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "rna-mpnn_b200")]
from rmpnn_b200 import ops
from rmpnn_b200.host import xgb_model as XM
g = np.load("tests/golden/xgb_v0_forest.npz")
f = XM.Forest(**{k: g[k] for k in ("threshold","feature","left","right","default_left","tree_root","tree_class")}, n_class=4, n_feat=128, base_score=0.5)
dev = ops.pack_forest(f, torch.device("cuda"))
for rows in (16384, 131072):
    X = torch.randn(rows, 128, device="cuda")
    for _ in range(3): ops.forest_predict(dev, X)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.forest_predict(dev, X)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"forest 400 trees depth 6, rows {rows}: {us:.1f} us per call, {rows / us:.2f} M rows/s, {rows * 400 / us / 1e3:.1f} G tree walks/s")
PY
