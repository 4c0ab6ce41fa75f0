"""Small end-to-end invocation of every kernel family for compute-sanitizer (memcheck): rdesign fp32 / fp32x3 / bf16 forward,
rnampnn fp32 / bf16 forward (cluster GraphNorm, tcgen05 attention), device collate, tree-ensemble head, the native training step.  Kept tiny: the
sanitizer slows kernels down by 10-100x."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rna-mpnn_b200")):
    sys.path.insert(0, p)
import torch
from rmpnn_b200.synthetic import (RDesignConfig, RNAMPNNConfig, make_rnampnn_state_dict, make_state_dict, synthetic_backbones)
from rmpnn_b200 import _lib, ops
from rmpnn_b200.engine import RDesignEngine
from rmpnn_b200.engine_rnampnn import RNAMPNNEngine

cfg = RDesignConfig(k_neighbors=8, num_mpnn_layers=2)
X, S, mask = synthetic_backbones(3, 70, 7, lengths=[70, 33, 5])
for prec in (_lib.PREC_FP32, _lib.PREC_FP32X3, _lib.PREC_BF16):
    out = RDesignEngine(make_state_dict(cfg, 1), 8, 30.0, device="cuda", precision=prec).forward(X.cuda(), mask.cuda())
    torch.cuda.synchronize()
    print("rdesign", prec, float(out["logits"].abs().max()))
rcfg = RNAMPNNConfig(num_res_mpnn_layers=2, num_post_fusion_attn_layers=1)
coords, _, rmask = synthetic_backbones(3, 150, 8, lengths=[150, 77, 2], atoms=7)
for prec in (_lib.PREC_FP32, _lib.PREC_BF16):
    out = RNAMPNNEngine.from_config(make_rnampnn_state_dict(rcfg, 1), rcfg, device="cuda", precision=prec).forward(coords.cuda(), rmask.cuda())
    torch.cuda.synchronize()
    print("rnampnn", prec, float(out["logits"].abs().max()))
packed = torch.randn(100, 7, 3, dtype=torch.float64).cuda()
Xc, mc = ops.collate(packed, torch.tensor([0, 60], dtype=torch.int64).cuda(), torch.tensor([60, 40], dtype=torch.int32).cuda(), 64, 6)
torch.cuda.synchronize()
print("collate", tuple(Xc.shape), float(mc.sum()))
import numpy as np
sys.path.insert(0, ROOT)
from oracle import forest_oracle as FO
from rmpnn_b200.host import xgb_model as XM
fr = FO.random_forest(np.random.default_rng(0), 20, 5, 128, 4)
f = XM.Forest(**{k: fr[k] for k in ("threshold", "feature", "left", "right", "default_left", "tree_root", "tree_class")},
              n_class=4, n_feat=128, base_score=0.5)
m, pr = ops.forest_predict(ops.pack_forest(f, torch.device("cuda")), torch.randn(300, 128, device="cuda"))
torch.cuda.synchronize()
print("forest", tuple(m.shape), int(pr.sum()))
qkv16 = torch.randn(2 * 150, 384, device="cuda").to(torch.bfloat16)
att = ops.res_attention_tc(qkv16, torch.ones(300, device="cuda"), 2, 150, 8)
torch.cuda.synchronize()
print("attention_tc", float(att.abs().max()))
# native training step (fused edge forward with dropout and saved tensors, backward GEMMs incl. the K-concatenated one,
# permuted segment sums, fused add + LayerNorm, column sums, Adam) on a tiny model; lengths 5 puts > 16 nodes into a tile
from rdesign.model.rdesign import RNAModel
from rmpnn_b200.train_engine import RDesignTrainer
tcfg = RDesignConfig(k_neighbors=8, num_mpnn_layers=2)
model = RNAModel(**tcfg.as_kwargs())
model.load_state_dict(make_state_dict(tcfg, 3), strict=True)
tr = RDesignTrainer(model.cuda().train())
for it in range(2):
    loss = tr.step(X.cuda(), S.cuda(), mask.cuda(), seed=it)
torch.cuda.synchronize()
print("training step", float(loss))
