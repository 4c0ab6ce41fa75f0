"""Tuning aid: launches the fp32-parity (split-bf16) message MLP alone on C2-sized synthetic operands (for ncu captures).

    python scripts/x3_launch.py [n_launches] [deg] [Nv] [variant]      variant: edge (default) | ffn | noadds | nogelu
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rna-mpnn_b200")]
import torch
from rmpnn_b200 import ops
n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 5
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 30
Nv = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
variant = sys.argv[4] if len(sys.argv) > 4 else "edge"
torch.manual_seed(0)
ctr = torch.repeat_interleave(torch.arange(Nv), deg).to(torch.int32).cuda()
Ne = ctr.numel()
nbr = torch.randint(0, Nv, (Ne,), dtype=torch.int32).cuda()
hE = torch.randn(Ne, 128).cuda()
hV = torch.randn(Nv, 128).cuda()
P = torch.randn(Nv, 256).cuda()
agg = torch.zeros(Nv, 128).cuda()
lin = lambda k, n: ops.pack_linear((torch.randn(n, k) / k ** 0.5).cuda(), torch.zeros(n).cuda())
msg = [lin(128, 128) for _ in range(3)]
ffn = [lin(128, 256), lin(256, 256), lin(256, 256), lin(256, 128)]
out = torch.empty(Nv, 128).cuda()


def launch():
    if variant == "ffn":
        ops.row_mlp([(hV, None)], [(m, True) for m in ffn[:-1]] + [(ffn[-1], False)], ops.EPI_STORE, out, Nv, split=True)
    else:
        adds = [] if variant == "noadds" else [(P, ctr), (P[:, 128:], nbr)]
        ops.row_mlp([(hE, None)], [(m, variant != "nogelu") for m in msg], ops.EPI_SEGSUM, agg, Ne, seg_id=ctr, alpha=1 / 30,
                    split=True, adds=adds, add_stride=256)


for _ in range(2):
    launch()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n_launch):
    launch()
e1.record(); torch.cuda.synchronize()
print(f"row_mlp_x3 {variant} rows={Nv if variant == 'ffn' else Ne}: {e0.elapsed_time(e1) * 1000 / n_launch:.1f} us per launch")

if os.environ.get("X3_TIMELINE"):
    import ctypes
    from rmpnn_b200 import _lib
    L = _lib.lib()
    L.rmp_set_row_x3_timeline.argtypes = [ctypes.c_void_p]; L.rmp_set_row_x3_timeline.restype = None
    buf = torch.zeros(500, dtype=torch.int64, device="cuda")
    L.rmp_set_row_x3_timeline(buf.data_ptr()); launch(); torch.cuda.synchronize(); L.rmp_set_row_x3_timeline(None)
    t = buf.cpu().tolist(); n = t[499]; prev = t[1]
    names = {1: "start", 2: "prologue done", 3: "end", 4: "adds in TMEM"}
    for i in range(min(n, 40)):
        tag, clk = t[2 * i], t[2 * i + 1]
        nm = names.get(tag, f"acc ready stage {tag - 10}" if tag < 20 else f"epilogue done stage {tag - 20}")
        print(f"{clk - t[1]:9d} (+{clk - prev:7d})  {nm}"); prev = clk
