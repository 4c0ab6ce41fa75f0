"""Tuning aid: SIMT vs tcgen05 attention of the rnampnn path on synthetic q | k | v.

    python scripts/attn_bench.py [B] [N] [heads]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200")]
import torch
from rmpnn_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 8
torch.manual_seed(0)
qkv = torch.randn(B * N, 384, device="cuda")
qkv16 = qkv.to(torch.bfloat16)
mask = torch.ones(B * N, device="cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n
a = timeit(lambda: ops.res_attention(qkv, mask, B, N, heads))
b = timeit(lambda: ops.res_attention_tc(qkv16, mask, B, N, heads))
d = (ops.res_attention(qkv16.float(), mask, B, N, heads) - ops.res_attention_tc(qkv16, mask, B, N, heads)).abs().max().item()
print(f"attention B={B} N={N} heads={heads}: SIMT fp32 {a:.1f} us, tcgen05 bf16 {b:.1f} us ({a / b:.2f}x), max |diff| {d:.2e}")
