"""Tuning aid: rmp_gemm_rows_bf16 variants at C4 size (409,600 rows), CUDA-event timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200")]
import torch
from rmpnn_b200 import ops
rows, Nv = int(sys.argv[1]) if len(sys.argv) > 1 else 409600, 16384
g = torch.Generator().manual_seed(0)
A = torch.randn(rows, 128, generator=g).cuda().to(torch.bfloat16)
W = (torch.randn(128, 128, generator=g) / 11.3).cuda().to(torch.bfloat16)
b = torch.randn(128, generator=g).cuda()
Pc, Pn = torch.randn(Nv, 128, generator=g).cuda().to(torch.bfloat16), torch.randn(Nv, 128, generator=g).cuda().to(torch.bfloat16)
ctr = torch.repeat_interleave(torch.arange(Nv), rows // Nv).to(torch.int32).cuda()
nbr = torch.randint(0, Nv, (rows,), generator=g, dtype=torch.int32).cuda()
acc = torch.zeros(rows, 128, device="cuda")
variants = {
    "z only (backward-data)": lambda: ops.gemm_rows(A, W, w_transposed=True),
    "z + bias": lambda: ops.gemm_rows(A, W, bias=b),
    "z + a (GELU, p=0)": lambda: ops.gemm_rows(A, W, bias=b, act=(0.0, 1, 2)),
    "z + a (GELU + dropout 0.1)": lambda: ops.gemm_rows(A, W, bias=b, act=(0.1, 1, 2)),
    "z + a + Pc/Pn gathers": lambda: ops.gemm_rows(A, W, Pc=Pc, Pn=Pn, ctr=ctr, nbr=nbr, act=(0.1, 1, 2)),
    "acc32 += only": lambda: ops.gemm_rows(A, W, want_z=False, acc32=acc, w_transposed=True),
}
which = sys.argv[2] if len(sys.argv) > 2 else None
for name, fn in variants.items():
    if which and which not in name:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:34s} {e0.elapsed_time(e1) * 100:.1f} us")
