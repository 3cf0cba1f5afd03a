#!/usr/bin/env python3
"""Summarise ncu outputs from gpurun_out/ into tracked text files under profiles/.

    python profiles/summarize.py launches gpurun_out/<launches>.csv > profiles/<name>.csv
    python profiles/summarize.py full gpurun_out/<rep>.ncu-rep > profiles/<name>.txt
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ik].split("(")[0]
        v = float(r[iv].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_us,avg_us,share_of_gpu_time")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k},{a[0]},{a[1] / 1e3:.1f},{a[1] / a[0] / 1e3:.2f},{a[1] / tot:.4f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:100])
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:70s} {r[i]:>16s} {units[i]}")
        d = float(r[hdr.index("dram__bytes_read.sum")]) + float(r[hdr.index("dram__bytes_write.sum")])
        t = float(r[hdr.index("gpu__time_duration.sum")])
        print(f"  {'dram traffic per launch (read+write)':70s} {d:16.3f} {units[hdr.index('dram__bytes_read.sum')]}")
        print(f"  {'=> dram throughput over the launch (Mbyte/us = TB/s)':70s} {d / t:16.2f} TB/s")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
