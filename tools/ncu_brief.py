#!/usr/bin/env python
"""Compact per-launch table out of an `ncu --page raw --csv` export (one row per profiled launch)."""
import csv, sys
KEYS = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%el"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__cycles_active.avg", "smcyc")]
for fn in sys.argv[1:]:
    with open(fn) as f:
        rd = list(csv.reader(l for l in f if not l.startswith("==")))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    print(fn)
    print("  " + " ".join(f"{k[1]:>10s}" for k in KEYS) + "  grid kernel")
    for r in rows:
        out = []
        for k, _ in KEYS:
            if k in hdr:
                i = hdr.index(k); v = r[i]; u = units[i]
                try:
                    v = float(v.replace(",", ""))
                    if u == "ns": v /= 1e3
                    if u == "ms": v *= 1e3
                    if u == "byte": v /= 1e6
                    if u == "Kbyte": v /= 1e3
                    if u == "Gbyte": v *= 1e3
                    out.append(f"{v:10.2f}")
                except ValueError:
                    out.append(f"{v:>10s}")
            else:
                out.append(f"{'-':>10s}")
        print("  " + " ".join(out) + "  " + r[hdr.index("Grid Size")] + " " + r[hdr.index("Kernel Name")][:48])
