import sys, os
ROOT="/root/repo"; sys.path[:0]=[ROOT, ROOT+"/rna-mpnn_b200"]
import torch
from rmpnn_b200 import ops
torch.manual_seed(0)
n = 16384
h = torch.randn(n, 128).cuda(); agg = torch.randn(n, 128).cuda()
g = lambda: (torch.ones(128).cuda(), torch.zeros(128).cuda())
dims = [128, 256, 256, 256, 128]
stages = [((torch.randn(dims[i + 1], dims[i]) / dims[i] ** 0.5).cuda().to(torch.bfloat16), torch.zeros(dims[i + 1]).cuda(), 0 if i < 3 else 1) for i in range(4)]
stages.append(((torch.randn(256, 128) / 11.3).cuda().to(torch.bfloat16), torch.zeros(256).cuda(), 2))
out = torch.empty(n, 128).cuda(); x1 = torch.empty(n, 128).cuda(); pp = torch.empty(n, 128).cuda(); pn = torch.empty(n, 128, dtype=torch.bfloat16).cuda()
def run():
    ops.node_mlp_bf16(h, stages, n, agg=agg, ln1=g1, ln2=g2, x1=x1, h_out=out, p_out=pp, pn_out=pn)
g1, g2 = g(), g()
for _ in range(3): run()
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for _ in range(10): run()
graph.replay(); torch.cuda.synchronize()
s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
s0.record(); graph.replay(); s1.record(); torch.cuda.synchronize()
print("node kernel us (graph of 10, warm L2):", s0.elapsed_time(s1) * 100)
