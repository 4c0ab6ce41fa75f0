"""Tuning aid: times the bf16 edge-message kernel's epilogue variants / producer-issuer flags (rmp_set_edge_tuning) on
C2-sized synthetic operands and checks every variant against variant 0 on the same inputs.

    python scripts/edge_variants.py [n_launches] [deg] [Nv]
"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200")]
import torch
from rmpnn_b200 import _lib
n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 20
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 30
Nv = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
torch.manual_seed(0)
ctr = torch.repeat_interleave(torch.arange(Nv), deg).to(torch.int32).cuda()
Ne = ctr.numel()
rp = (torch.arange(Nv + 1) * deg).to(torch.int32).cuda()
nbr = torch.randint(0, Nv, (Ne,), dtype=torch.int32).cuda()
hE = torch.randn(Ne, 128).cuda().to(torch.bfloat16)
W = (torch.randn(384, 128) / 11.3).cuda().to(torch.bfloat16)
b = [None, (0.1 * torch.randn(128)).cuda(), (0.1 * torch.randn(128)).cuda()]
Pc = torch.randn(Nv, 128).cuda(); Pn = torch.randn(Nv, 128).cuda().to(torch.bfloat16)
agg = torch.zeros(Nv, 128).cuda()
barr = (C.c_void_p * 4)(None, b[1].data_ptr(), b[2].data_ptr(), None)
L = _lib.lib()
L.rmp_set_edge_tuning.argtypes = [C.c_int, C.c_int]
L.rmp_set_edge_tuning.restype = None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def launch():
    _lib.check(L.rmp_edge_mlp_bf16(hE.data_ptr(), W.data_ptr(), 3, C.cast(barr, C.c_void_p), Pc.data_ptr(),
               Pn.data_ptr(), ctr.data_ptr(), nbr.data_ptr(), rp.data_ptr(), Ne, 1 / 30, agg.data_ptr(), None,
               torch.cuda.current_stream().cuda_stream))
ref = None
configs = [(0, 0), (0, 2), (0, 3), (1, 0), (1, 3)]
if len(sys.argv) > 4:
    configs = [tuple(int(x) for x in c.split(",")) for c in sys.argv[4:]]
for v, f in configs:
    L.rmp_set_edge_tuning(v, f)
    agg.zero_(); launch(); torch.cuda.synchronize()
    out = agg.clone()
    if True:
        if ref is None:
            ref = out
        err = (out - ref).abs().max().item()
    else:
        err = float("nan")                      # timing-only variant (the 1/2 of GELU is not folded on the host here)
    ts = []
    for _ in range(n_launch):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    ts.sort()
    print(f"variant {v} flags {f}: median {ts[len(ts) // 2]:.1f} us  min {ts[0]:.1f} us  max|agg - agg(v0,f0)| {err:.3e}", flush=True)
L.rmp_set_edge_tuning(1, 3)
