"""Tuning aid: the fused training forward of the edge kernel (rmp_edge_mlp_train_fwd_bf16) against the inference launch and
the stage-by-stage path it replaces, C4 shapes (16,384 nodes x K 25)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200")]
import torch
from rmpnn_b200 import ops
Nv, deg = 16384, 25
g = torch.Generator().manual_seed(0)
ctr = torch.repeat_interleave(torch.arange(Nv), deg).to(torch.int32).cuda()
Ne = ctr.numel()
rp = (torch.arange(Nv + 1) * deg).to(torch.int32).cuda()
nbr = torch.randint(0, Nv, (Ne,), generator=g, dtype=torch.int32).cuda()
hE = torch.randn(Ne, 128, generator=g).cuda().to(torch.bfloat16)
W = (torch.randn(384, 128, generator=g) / 11.3).cuda().to(torch.bfloat16)
b = [None, (0.1 * torch.randn(128, generator=g)).cuda(), (0.1 * torch.randn(128, generator=g)).cuda()]
Pc = torch.randn(Nv, 128, generator=g).cuda(); Pn = torch.randn(Nv, 128, generator=g).cuda().to(torch.bfloat16)
Pc16 = Pc.to(torch.bfloat16)
agg = torch.zeros(Nv, 128, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    ts.sort()
    return ts[len(ts) // 2]

def staged(p):
    z0, a0 = ops.gemm_rows(hE, W[:128], Pc=Pc16, Pn=Pn, ctr=ctr, nbr=nbr, act=(p, 1, 0))
    z1, a1 = ops.gemm_rows(a0, W[128:256], bias=b[1], act=(p, 1, 1))
    z2, a2 = ops.gemm_rows(a1, W[256:], bias=b[2], act=(p, 1, 2))
    return ops.segment_sum_rows(a2, rp)

print(f"inference edge_mlp_bf16                      {timed(lambda: ops.edge_mlp_bf16(hE, W, b, Pc, Pn, ctr, nbr, rp, agg, 1 / 30)):8.1f} us")
for p in (0.0, 0.1):
    print(f"train fwd fused, p_drop {p}                   {timed(lambda: ops.edge_mlp_train_fwd(hE, W, b, Pc, Pn, ctr, nbr, rp, agg, 1 / 30, p, 1, 0)):8.1f} us"
          "   (includes the torch.empty of the five saved tensors)")
    print(f"train fwd stage by stage (3 GEMMs + segsum), p {p}  {timed(lambda: staged(p)):8.1f} us")
