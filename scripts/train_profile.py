"""Tuning aid: torch.profiler kernel table of one native training step (C4 shapes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200")]
import torch
from rdesign.model.rdesign import RNAModel
from rmpnn_b200.synthetic import RDesignConfig, make_state_dict, synthetic_backbones
from rmpnn_b200.train_engine import RDesignTrainer
cfg = RDesignConfig()
model = RNAModel(**cfg.as_kwargs()); model.load_state_dict(make_state_dict(cfg, 42), strict=True)
model = model.cuda().train()
tr = RDesignTrainer(model)
X, S, mask = (t.cuda() for t in synthetic_backbones(32, 512, 1004))
for _ in range(3):
    tr.step(X, S, mask, sizes=(32 * 512, 32 * 512 * 25))
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        tr.step(X, S, mask, sizes=(32 * 512, 32 * 512 * 25))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
