import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200"), os.path.join(ROOT, "tests")]
import torch
from conftest import load_golden
from oracle import rdesign_oracle as O
from rmpnn_b200 import ops
g = load_golden("ragged_k8")
X, mask = g["X"], g["mask"]
D_o, I_o = O.knn(X[:, :, 0, :], mask, 8)
D_c, I_c = ops.knn(X.cuda(), mask.cuda(), 8)
D_c = D_c.cpu()
bad = (D_c != D_o).nonzero()
print("mismatches", bad.shape[0])
for b, i, k in bad[:10].tolist():
    j = int(I_o[b, i, k])
    pi, pj = X[b, i, 0], X[b, j, 0]
    print(b, i, k, j, D_c[b, i, k].item().hex(), D_o[b, i, k].item().hex(), pi.tolist(), pj.tolist(), mask[b, i].item(), mask[b, j].item())
