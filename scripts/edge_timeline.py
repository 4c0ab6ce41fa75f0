"""Tuning aid: clock64 trace of one epilogue thread of the tcgen05 edge kernel (CTA 0)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200")]
import torch
from rmpnn_b200 import _lib, ops
torch.manual_seed(0)
Nv, deg = 16384, 30
ctr = torch.repeat_interleave(torch.arange(Nv), deg).to(torch.int32).cuda()
Ne = ctr.numel()
rp = (torch.arange(Nv + 1) * deg).to(torch.int32).cuda()
nbr = torch.randint(0, Nv, (Ne,), dtype=torch.int32).cuda()
hE = torch.randn(Ne, 128).cuda().to(torch.bfloat16)
W = (torch.randn(384, 128) / 11.3).cuda().to(torch.bfloat16)
b = [None, torch.zeros(128).cuda(), torch.zeros(128).cuda()]
Pc = torch.randn(Nv, 128).cuda(); Pn = torch.randn(Nv, 128).cuda().to(torch.bfloat16)
agg = torch.zeros(Nv, 128).cuda()
tl = torch.zeros(8000, dtype=torch.int64).cuda()
barr = (C.c_void_p * 4)(None, b[1].data_ptr(), b[2].data_ptr(), None)
L = _lib.lib()
for _ in range(3):
    _lib.check(L.rmp_edge_mlp_bf16_traced(hE.data_ptr(), W.data_ptr(), 3, C.cast(barr, C.c_void_p), Pc.data_ptr(),
               Pn.data_ptr(), ctr.data_ptr(), nbr.data_ptr(), rp.data_ptr(), Ne, 1 / 30, agg.data_ptr(), None,
               tl.data_ptr(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
t = tl.cpu().tolist()
n = t[7999]
ev = [(t[2 * i], t[2 * i + 1]) for i in range(n)]
t0 = ev[0][1]
print("events", n, "total cycles", ev[-1][1] - t0)
# per-kind average deltas to the previous event
import collections
acc = collections.defaultdict(list)
for (a, ta), (b_, tb) in zip(ev[:-1], ev[1:]):
    acc[(a, b_)].append(tb - ta)
for k in sorted(acc, key=lambda k: -sum(acc[k])):
    v = acc[k]
    print(f"{k[0]:4d} -> {k[1]:4d}: n={len(v):4d} mean={sum(v)/len(v):8.0f} total={sum(v):9d}")
print("first 60 events:")
for a, ta in ev[:60]:
    print(a, ta - t0)
