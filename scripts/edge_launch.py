"""Tuning aid: launches the bf16 edge-message kernel alone on C2-sized synthetic operands (for ncu captures).

    python scripts/edge_launch.py [n_launches] [deg] [Nv]
"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200")]
import torch
from rmpnn_b200 import _lib
n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 5
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 30
Nv = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
torch.manual_seed(0)
ctr = torch.repeat_interleave(torch.arange(Nv), deg).to(torch.int32).cuda()
Ne = ctr.numel()
rp = (torch.arange(Nv + 1) * deg).to(torch.int32).cuda()
nbr = torch.randint(0, Nv, (Ne,), dtype=torch.int32).cuda()
hE = torch.randn(Ne, 128).cuda().to(torch.bfloat16)
W = (torch.randn(384, 128) / 11.3).cuda().to(torch.bfloat16)
b = [None, torch.zeros(128).cuda(), torch.zeros(128).cuda()]
Pc = torch.randn(Nv, 128).cuda(); Pn = torch.randn(Nv, 128).cuda().to(torch.bfloat16)
agg = torch.zeros(Nv, 128).cuda()
barr = (C.c_void_p * 4)(None, b[1].data_ptr(), b[2].data_ptr(), None)
L = _lib.lib()
def launch():
    _lib.check(L.rmp_edge_mlp_bf16(hE.data_ptr(), W.data_ptr(), 3, C.cast(barr, C.c_void_p), Pc.data_ptr(),
               Pn.data_ptr(), ctr.data_ptr(), nbr.data_ptr(), rp.data_ptr(), Ne, 1 / 30, agg.data_ptr(), None,
               torch.cuda.current_stream().cuda_stream))
for _ in range(2):
    launch()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n_launch):
    launch()
e1.record(); torch.cuda.synchronize()
print(f"edge_mlp_bf16 Ne={Ne}: {e0.elapsed_time(e1) * 1000 / n_launch:.1f} us per launch")
