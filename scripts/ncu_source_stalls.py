"""Summarise the per-instruction warp-stall samples of an `ncu --set full --import-source on` capture.

    ncu -i prof.ncu-rep --page source --csv --print-source sass > src.csv
    python scripts/ncu_source_stalls.py src.csv [lo_addr_hex hi_addr_hex]

Prints total samples per stall reason (optionally restricted to an address range = one warp role) and the 40
instructions with the most samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
base = int(rows[2][0], 16)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 60
tot = {n: 0 for n in stall_cols}
insts = []
n_exec = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    off = int(r[0], 16) - base
    if not (lo <= off < hi):
        continue
    s = int(r[col["# Samples"]] or 0)
    n_exec += int(r[col["Instructions Executed"]] or 0)
    for n in stall_cols:
        tot[n] += int(r[col[n]] or 0)
    insts.append((s, off, r[1].strip(), {n: int(r[col[n]] or 0) for n in stall_cols if int(r[col[n]] or 0)}))
all_s = sum(tot.values())
print(f"range [{lo:#x}, {hi:#x}): {len(insts)} instructions, {n_exec} warp-instructions executed, {all_s} samples")
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {n:24s} {v:8d}  {100.0 * v / max(1, all_s):5.1f} %")
print("top instructions:")
for s, off, src, d in sorted(insts, key=lambda t: -t[0])[:40]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:3])
    print(f"  {off:#07x} {s:6d}  {src[:60]:60s} {top}")
