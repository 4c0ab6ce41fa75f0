"""Tuning aid: clock64 trace of one epilogue thread of the tcgen05 node kernel (CTA 0)."""
import ctypes as C, os, sys
os.environ["RMP_NODE_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200")]
import torch
from rmpnn_b200 import _lib, ops
torch.manual_seed(0)
n = 16384
h = torch.randn(n, 128).cuda(); agg = torch.randn(n, 128).cuda()
g = lambda: (torch.ones(128).cuda(), torch.zeros(128).cuda())
dims = [128, 256, 256, 256, 128]
stages = [((torch.randn(dims[i + 1], dims[i]) / dims[i] ** 0.5).cuda().to(torch.bfloat16), torch.zeros(dims[i + 1]).cuda(), 0 if i < 3 else 1) for i in range(4)]
stages.append(((torch.randn(256, 128) / 11.3).cuda().to(torch.bfloat16), torch.zeros(256).cuda(), 2))
tl = torch.zeros(800, dtype=torch.int64).cuda()
L = _lib.lib()
L.rmp_set_node_timeline.argtypes = [C.c_void_p]; L.rmp_set_node_timeline.restype = None
L.rmp_set_node_timeline(tl.data_ptr())
out = torch.empty(n, 128).cuda(); x1 = torch.empty(n, 128).cuda(); pp = torch.empty(n, 128).cuda(); pn = torch.empty(n, 128, dtype=torch.bfloat16).cuda()
for _ in range(3):
    ops.node_mlp_bf16(h, stages, n, agg=agg, ln1=g(), ln2=g(), x1=x1, h_out=out, p_out=pp, pn_out=pn)
torch.cuda.synchronize()
s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(20):
    ops.node_mlp_bf16(h, stages, n, agg=agg, ln1=g(), ln2=g(), x1=x1, h_out=out, p_out=pp, pn_out=pn)
s1.record(); torch.cuda.synchronize()
print("avg launch us", s0.elapsed_time(s1) * 1000 / 20)
t = tl.cpu().tolist(); k = t[799]
ev = [(t[2 * i], t[2 * i + 1]) for i in range(k)]
for (a, ta), (b, tb) in zip(ev[:-1], ev[1:]):
    print(f"{a:3d} -> {b:3d}: {tb - ta}")
print("total", ev[-1][1] - ev[0][1])
