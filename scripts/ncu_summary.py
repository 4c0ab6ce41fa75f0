#!/usr/bin/env python
"""Summarise ncu outputs into small text files under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.txt
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r1_kernel.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active" ,
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
    hdr = rows[0]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0]
        tot[name] += float(r[mv].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum launch list, summed per kernel (ns); source {src}\n")
        f.write(f"# cold-cache, serialised launches: compare SHARES, not absolutes.  total {total:.0f} ns over "
                f"{sum(cnt.values())} launches\n")
        f.write(f"{'share':>7} {'total_ns':>14} {'launches':>8} {'avg_ns':>12}  kernel\n")
        for name in sorted(tot, key=tot.get, reverse=True):
            f.write(f"{tot[name] / total:7.3%} {tot[name]:14.0f} {cnt[name]:8d} {tot[name] / cnt[name]:12.0f}  {name}\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none; source {src}\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}\n")
            for m in FULL_METRICS:
                if m in hdr:
                    f.write(f"{m:78s} {r[hdr.index(m)]:>18s} {units[hdr.index(m)]}\n")
    print(open(dst).read())


def table(src, dst):
    """One row per kernel (mean over its launches in the capture): time, DRAM bytes and GB/s against the measured
    HBM peak, tensor / FMA / XU pipe activity.  The per-kernel roofline table of a whole forward pass."""
    import json, os
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {m: hdr.index(m) for m in FULL_METRICS if m in hdr}
    peaks_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0

    def val(r, m):
        v = float(r[col[m]].replace(",", ""))
        u = units[col[m]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        return v * scale
    agg = defaultdict(lambda: defaultdict(float))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        a = agg[name]
        a["n"] += 1
        a["us"] += val(r, "gpu__time_duration.sum")
        a["bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        for k, m in (("tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     ("fma", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                     ("xu", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                     ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active")):
            if m in col:
                a[k] += val(r, m)
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, one forward pass; source {src}\n")
        f.write(f"# per kernel, mean over its launches (cold-cache, serialised).  HBM peak {hbm:.0f} GB/s (MEASURED_PEAKS.json)\n")
        f.write(f"{'launches':>8} {'us':>9} {'DRAM MB':>9} {'GB/s':>8} {'of HBM':>7} {'tensor%':>8} {'fma%':>6} {'xu%':>6} {'issue%':>7}  kernel\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            n = a["n"]
            us, mb = a["us"] / n, a["bytes"] / n / 1e6
            gbs = a["bytes"] / n / (us * 1e-6) / 1e9 if us > 0 else 0.0
            f.write(f"{int(n):8d} {us:9.1f} {mb:9.2f} {gbs:8.0f} {gbs / hbm:7.1%} {a['tensor'] / n:8.1f} {a['fma'] / n:6.1f} "
                    f"{a['xu'] / n:6.1f} {a['issue'] / n:7.1f}  {name}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full, "table": table}[sys.argv[1]](sys.argv[2], sys.argv[3])
