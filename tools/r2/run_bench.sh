#!/bin/bash
# usage: bash tools/r2/run_bench.sh <tag> [bench args...]
tag=$1; shift
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/$tag/build.log; exit 1; }
timeout -s KILL 900 python bench.py "$@" > gpurun_out/$tag/bench.json 2> gpurun_out/$tag/bench.err
echo "bench rc=$?"; tail -n 5 gpurun_out/$tag/bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/$tag/bench.json").read().strip().splitlines()[-1])
except Exception as e:
    print("parse failed", e); raise SystemExit
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "graph", d.get("cuda_graph_active"), "launches", d["gpu_launches"])
print("timing:", d.get("kernel_timing")); print("kernel sum ms", d.get("kernel_ms_sum_per_step"), "other", d.get("other_kernels_ms_per_step"))
print("roofline", d.get("roofline")); print("cfg3", d.get("cfg3") and {k: d["cfg3"][k] for k in ("value","ms_per_step","cuda_graph_active")})
print("gpu_eager", d.get("gpu_eager_baseline")); print("dropin", d.get("dropin_eager")); print("cpu", d.get("cpu_baseline")); print("clocks", d.get("clocks"))
if d.get("lif_microbench"): print("lif min frac fwd/bwd", d["lif_microbench"]["fwd_frac_min"], d["lif_microbench"]["bwd_frac_min"], d["lif_microbench"]["clocks"])
for k, v in sorted((d.get("kernels") or {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print("  %-28s %7.3f ms %5.0f calls %s" % (k, v["ms_per_step"], v["calls_per_step"], {a: round(b, 2) for a, b in v.items() if a in ("tflops", "gbs", "frac_of_hbm", "frac_of_bf16_burst")}))
PY
