"""clock64 trace of CTA 0's first epilogue thread of the tcgen05 row-MLP kernel on the ResMPNN message shape
(49,152 gathered rows, 384 -> 128 -> 128, SEGSUM) and on the FFN shape (16,384 rows, 128 -> 512 -> 512 -> 512 -> 128)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rna-mpnn_b200")):
    sys.path.insert(0, p)
import torch
from rmpnn_b200 import ops, _lib

L = _lib.lib()
L.rmp_set_row_tc_timeline.argtypes = [ctypes.c_void_p]
L.rmp_set_row_tc_timeline.restype = None
g = torch.Generator().manual_seed(0)
dev = "cuda"
lin = lambda n, k: ops.pack_linear((torch.randn(n, k, generator=g) / k ** 0.5).to(dev), (torch.randn(n, generator=g) * 0.1).to(dev))
TAGS = {1: "start", 2: "prologue done", 3: "end"}

def show(buf, title):
    t = buf.cpu().tolist()
    n = t[499]
    print(f"--- {title}: {n} stamps")
    t0 = prev = t[1]
    for i in range(n):
        tag, clk = t[2 * i], t[2 * i + 1]
        name = TAGS.get(tag, f"acc ready stage {tag - 10}" if tag < 20 else f"epilogue done stage {tag - 20}")
        print(f"{clk - t0:9d} (+{clk - prev:7d})  {name}")
        prev = clk

def timed(fn, title):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{title}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
    buf = torch.zeros(500, dtype=torch.int64, device=dev)
    L.rmp_set_row_tc_timeline(buf.data_ptr())
    fn(); torch.cuda.synchronize()
    L.rmp_set_row_tc_timeline(None)
    show(buf, title)

nodes, K = 16384, 3
rows = nodes * K
h = torch.randn(nodes, 128, generator=g).to(dev); e = torch.randn(rows, 128, generator=g).to(dev)
ctr = (torch.arange(rows, dtype=torch.int32) // K).to(dev)
nbr = torch.randint(0, nodes, (rows,), generator=g, dtype=torch.int32).to(dev)
w = torch.full((rows,), 1 / 3, device=dev)
msg = [(lin(128, 384), True), (lin(128, 128), True)]
agg = torch.zeros(nodes, 128, device=dev)
timed(lambda: ops.row_mlp([(h, ctr), (h, nbr), (e, None)], msg, ops.EPI_SEGSUM, agg, rows, seg_id=ctr, alpha=1.0,
                          row_scale=w, tensor_core=True), "message 49152 x (384->128->128) SEGSUM")
h16, e16 = h.to(torch.bfloat16), e.to(torch.bfloat16)
timed(lambda: ops.row_mlp([(h16, ctr), (h16, nbr), (e16, None)], msg, ops.EPI_SEGSUM, agg, rows, seg_id=ctr, alpha=1.0,
                          row_scale=w, tensor_core=True), "message, bf16 side copies gathered with LDG.256")
e2 = e.clone()
timed(lambda: ops.row_mlp([(h16, ctr), (h16, nbr), (e16, None)], msg, ops.EPI_RESID_ADD, e2, rows, resid=e2,
                          tensor_core=True, out_bf16=e16), "edge update, bf16 side copies, fp32 master + bf16 copy out")
ffn = [(lin(512, 128), True), (lin(512, 512), True), (lin(512, 512), True), (lin(128, 512), False)]
out = torch.empty(nodes, 128, device=dev)
timed(lambda: ops.row_mlp([(h, None)], ffn, ops.EPI_STORE, out, nodes, tensor_core=True), "FFN 16384 x (128->512->512->512->128)")
