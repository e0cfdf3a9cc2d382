#!/usr/bin/env python
"""Join an ncu raw-page CSV (one row per profiled launch, in launch order, ONE eager training step) with the ABI call trace
of the same step (bench.py --dump-trace) -> per-launch table (shape, duration, DRAM bytes, algorithmic work) and the mean
DRAM bytes per launch of every ABI entry point (profiles/traffic.json, read by bench.py for `roofline.traffic`).

    python tools/r2/make_traffic.py trace.json raw_conv.csv [raw_lif.csv ...] --out profiles/r2/traffic_v1 [--traffic profiles/traffic.json]
"""
import argparse, csv, json, sys
sys.path.insert(0, ".")
from bench import EXPECT

ap = argparse.ArgumentParser()
ap.add_argument("trace"); ap.add_argument("raw", nargs="+"); ap.add_argument("--out", required=True); ap.add_argument("--traffic", default=None)
a = ap.parse_args()
trace = json.load(open(a.trace))
rows = []
for fn in a.raw:
    with open(fn) as f:
        rd = list(csv.reader(l for l in f if not l.startswith("==")))
    hdr, units = rd[0], rd[1]
    def col(r, name, scale=1.0):
        if name not in hdr: return None
        i = hdr.index(name); v = float(r[i].replace(",", "")); u = units[i]
        mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        return v * mult * scale
    for r in rd[2:]:
        rows.append(dict(kernel=r[hdr.index("Kernel Name")], us=col(r, "gpu__time_duration.sum"), rd=col(r, "dram__bytes_read.sum"),
                         wr=col(r, "dram__bytes_write.sum"), tensor=col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                         dram=col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")))
# expected kernel sequence per family
fam = {}
for name, work in trace:
    for k, sub in enumerate(EXPECT.get(name, [])):
        fam.setdefault(sub, []).append((name, work if k == 0 else None))
out, per_abi = [], {}
for sub, calls in fam.items():
    got = [r for r in rows if sub in r["kernel"]]
    if not got:
        continue
    if len(got) != len(calls):
        print(f"# {sub}: {len(got)} profiled launches vs {len(calls)} in the trace -- skipped", file=sys.stderr)
        continue
    for (name, work), r in zip(calls, got):
        tag = work[2] if work and len(work) > 2 else ""
        amount = work[1] if work else None
        traffic = (r["rd"] or 0) + (r["wr"] or 0)
        out.append(dict(abi=name, shape=tag, us=round(r["us"], 2), dram_bytes=traffic, work_kind=work[0] if work else None, work=amount,
                        tensor_pct=r["tensor"], dram_pct=r["dram"]))
        d = per_abi.setdefault(name, dict(launches=0, dram_bytes=0.0, us=0.0, work=0.0))
        d["launches"] += 1; d["dram_bytes"] += traffic; d["us"] += r["us"]; d["work"] += amount or 0.0
json.dump(out, open(a.out + "_per_launch.json", "w"), indent=0)
summary = {k: dict(launches=v["launches"], mean_dram_bytes_per_launch=v["dram_bytes"] / v["launches"], ms_under_ncu=v["us"] / 1e3,
                   algorithmic_work=v["work"]) for k, v in per_abi.items()}
json.dump(summary, open(a.out + "_summary.json", "w"), indent=1)
if a.traffic:
    json.dump({k: v["mean_dram_bytes_per_launch"] for k, v in summary.items()}, open(a.traffic, "w"), indent=1)
with open(a.out + "_per_launch.md", "w") as f:
    f.write("| ABI call | shape | us (ncu, cold, serialised) | DRAM MB | algorithmic | tensor % | dram % |\n|---|---|---|---|---|---|---|\n")
    for r in sorted(out, key=lambda r: -r["us"])[:60]:
        alg = "" if not r["work"] else (f"{r['work'] / 1e9:.1f} GFLOP" if r["work_kind"] == "flop" else f"{r['work'] / 1e6:.1f} MB")
        f.write(f"| {r['abi']} | {r['shape']} | {r['us']} | {r['dram_bytes'] / 1e6:.1f} | {alg} | {r['tensor_pct']} | {r['dram_pct']} |\n")
print(json.dumps(summary, indent=1))
