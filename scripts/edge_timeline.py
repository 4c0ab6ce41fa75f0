"""Tuning aid: clock64 trace of the tcgen05 edge kernel (CTA 0): every epilogue warp and the MMA thread."""
import ctypes as C, os, sys, collections
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
tl = torch.zeros(32768, dtype=torch.int64).cuda()
barr = (C.c_void_p * 4)(None, b[1].data_ptr(), b[2].data_ptr(), None)
L = _lib.lib()
def launch(t):
    _lib.check(L.rmp_edge_mlp_bf16_traced(hE.data_ptr(), W.data_ptr(), 3, C.cast(barr, C.c_void_p), Pc.data_ptr(),
               Pn.data_ptr(), ctr.data_ptr(), nbr.data_ptr(), rp.data_ptr(), Ne, 1 / 30, agg.data_ptr(), None,
               t, torch.cuda.current_stream().cuda_stream))
for _ in range(3):
    launch(tl.data_ptr())
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    launch(None)
e1.record(); torch.cuda.synchronize()
print("untraced launch us", e0.elapsed_time(e1) * 1000 / 20)
t = tl.cpu().tolist()
warps = {}
for ew in range(16):
    n = t[32000 + ew]
    warps[ew] = [(t[ew * 1600 + 2 * i], t[ew * 1600 + 2 * i + 1]) for i in range(n)]
t0 = min(w[0][1] for w in warps.values() if w)
print("total cycles (warp 0)", warps[0][-1][1] - warps[0][0][1])
for ew in (0, 5, 8):
    ev = warps[ew]
    acc = collections.defaultdict(list)
    for (a, ta), (b_, tb) in zip(ev[:-1], ev[1:]):
        acc[(a, b_)].append(tb - ta)
    print(f"--- epilogue warp {ew} (group {ew >> 3}) mean deltas")
    for k in sorted(acc, key=lambda k: -sum(acc[k])):
        v = acc[k]
        print(f"{k[0]:4d} -> {k[1]:4d}: n={len(v):4d} mean={sum(v)/len(v):8.0f} total={sum(v):9d}")
# per (tile, stage): when each warp got the accumulator (2x0) and when it arrived (7x0)
print("--- per group-tile stage windows: min/max over the 8 warps of acc-ready (200+10s) and arrive (700+10s)")
for g in range(2):
    ws = [warps[g * 8 + j] for j in range(8)]
    ntile = sum(1 for tag, _ in ws[0] if tag == 90)
    for k in range(3, min(ntile, 7)):
        for s in range(3):
            rdy = [[ts for tag, ts in w if tag == 200 + 10 * s][k] - t0 for w in ws]
            arr = [[ts for tag, ts in w if tag == 700 + 10 * s][k] - t0 for w in ws]
            print(f"g{g} tile#{k} s{s}: ready {min(rdy)}..{max(rdy)}  arrive {min(arr)}..{max(arr)}  "
                  f"group E = {max(arr) - min(rdy)}  per-warp E = {[a - r for a, r in zip(arr, rdy)]}")
m = t[32016]
print("MMA thread issues (tile+1)*1000+stage, cycle, delta:")
prev = None
for i in range(min(m, 120)):
    tag, ts = t[25600 + 2 * i], t[25601 + 2 * i]
    print(tag, ts - t0, (ts - prev) if prev is not None else 0)
    prev = ts
# staggered schedule: for tiles 6..11 of CTA 0 and every stage -- when the 16 epilogue warps arrived (first / last), when the
# MMA thread issued the next stage's GEMM, and when the warps saw that accumulator (first / last)
if os.environ.get("RMP_EDGE_SCHEDULE", "s")[0] == "s":
    mma = {t[25600 + 2 * i]: t[25601 + 2 * i] - t0 for i in range(m)}
    print("--- hand-off chain per (tile, stage): arrive first/last -> MMA issue of the next stage -> accumulator seen first/last")
    for tile in range(6, 12):
        nth = tile // 2                      # occurrence index within the tile's parity stream
        par = tile & 1
        for s in range(3):
            arr = [[ts for tag, ts in warps[w] if tag == 700 + 10 * s + par][nth] - t0 for w in range(16)]
            wait0 = [[ts for tag, ts in warps[w] if tag == 100 + 10 * s + par][nth] - t0 for w in range(16)]
            line = f"tile {tile} stage {s}: wait starts {min(wait0)}..{max(wait0)}  arrive {min(arr)}..{max(arr)} (spread {max(arr) - min(arr)})"
            if s < 2:
                issue = mma.get((tile + 1) * 1000 + s + 1)
                seen = [[ts for tag, ts in warps[w] if tag == 200 + 10 * (s + 1) + par][nth] - t0 for w in range(16)]
                nwait = [[ts for tag, ts in warps[w] if tag == 100 + 10 * (s + 1) + par][nth] - t0 for w in range(16)]
                line += (f"  | next GEMM issued {issue} (+{issue - max(arr)} after the last arrive)  next wait starts {min(nwait)}..{max(nwait)}"
                         f"  accumulator seen {min(seen)}..{max(seen)} (+{min(seen) - issue} after the issue)")
            print(line)
